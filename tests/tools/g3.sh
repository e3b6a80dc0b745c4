export FITSNE_BENCH_CACHE=/tmp/fitsne_cache
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2g_pytest_gpu.txt
timeout 200 python bench.py --steps 300 --no-cpu-baseline --no-e2e 2>&1 | tail -1 > gpurun_out/r2g_bench.json
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r2g_bench.json')); print('value %.1f' % d['value'], 'ms %.4f' % d['ms_per_step'], d['grid'], {k: v['ms'] for k, v in d['kernels'].items()}, d['kl_last'])
except Exception as e: print('FAILED', open('gpurun_out/r2g_bench.json').read()[-1500:])
PY
