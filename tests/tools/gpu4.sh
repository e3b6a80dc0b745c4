python -m pytest tests -m gpu -x -q 2>&1 | tail -15
FITSNE_TRACE=1 python tests/gpu_report.py 1000000 2>&1 | grep -E "reorder|it/s|phases" | tail -14
python bench.py --steps 500 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_1gpu.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_1gpu.json'))
print({k: d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches','regrids','grid')}); print('e2e', d['e2e']); print(d['roofline']); print(d['kernels'])
PY
