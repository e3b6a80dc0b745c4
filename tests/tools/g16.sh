export FITSNE_BENCH_CACHE=/tmp/fitsne_cache
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 200 --warmup 10 --no-cpu-baseline ) > gpurun_out/r2q_bench_8gpu.json 2> gpurun_out/r2q_bench_8gpu.err
tail -4 gpurun_out/r2q_bench_8gpu.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/r2q_bench_8gpu.json').read().strip().splitlines() if l.startswith('{')][-1])
    print('value %.1f' % d['value'], 'ms %.4f' % d['ms_per_step'], d['grid'])
    print('e2e', d['e2e']['value']); print('parity', d['parity'])
    print('kernels', {k: v['ms'] for k, v in d['kernels'].items()})
    for k, v in (d['other_configs'] or {}).items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('error'))
except Exception as e: print('FAILED', e, open('gpurun_out/r2q_bench_8gpu.json').read()[-1500:])
PY
