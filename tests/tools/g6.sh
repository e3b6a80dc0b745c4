export FITSNE_BENCH_CACHE=/tmp/fitsne_cache
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python bench.py --steps 300 --no-cpu-baseline --no-e2e 2>&1 | tail -1 > gpurun_out/r2j_bench.json
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r2j_bench.json')); print('value %.1f' % d['value'], 'ms %.4f' % d['ms_per_step'], d['grid'], {k: v['ms'] for k, v in d['kernels'].items()}, d['kl_last'])
except Exception as e: print('FAILED', open('gpurun_out/r2j_bench.json').read()[-1500:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --launch-skip 3300 -c 80 --csv --log-file gpurun_out/r2_launches_warm_late.csv python tests/tools/profile_steps.py 1000000 late 230 > /dev/null 2>&1
