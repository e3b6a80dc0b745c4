python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python tests/gpu_report.py 1000000 2>&1 | tail -8
python bench.py --steps 500 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_r1_b.json; python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_r1_b.json'))
print({k: d[k] for k in ('value','ms_per_step','gpu_launches','regrids','grid','clocks')})
print('e2e', d['e2e']); print('roofline', d['roofline']); print('kernels', d['kernels'])
PY
