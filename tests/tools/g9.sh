( time timeout 900 python bench.py --steps 200 --warmup 10 ) > gpurun_out/r2k_bench_full.json 2> gpurun_out/r2k_bench_full.err
tail -5 gpurun_out/r2k_bench_full.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2k_bench_full.json').read().strip().splitlines()[-1])
    print('value %.1f' % d['value'], 'ms %.4f' % d['ms_per_step'], d['grid'])
    print('e2e', d['e2e']); print('parity', d['parity']); print('roofline', d['roofline']); print('total', d['roofline_total'])
    print('kernels', {k: v['ms'] for k, v in d['kernels'].items()}); print('cpu', d['cpu_baseline'])
    for k, v in (d['other_configs'] or {}).items(): print(k, v)
except Exception as e: print('FAILED', e, open('gpurun_out/r2k_bench_full.json').read()[-1500:])
PY
