python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tests/gpu_report.py 1000000 2>&1 | tail -6
python bench.py --steps 500 --warmup 10 2>&1 | tail -1 > gpurun_out/bench_1gpu.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_1gpu.json'))
print({k: d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches','regrids','grid','clocks')}); print('e2e', d['e2e']); print(d['roofline']); print(d['kernels']); print(d['cpu_baseline'])
PY
