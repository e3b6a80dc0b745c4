export FITSNE_BENCH_CACHE=/tmp/fitsne_cache
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/tools/mgpu_check.py 2>&1 | tail -8
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 200 --warmup 10 ) > gpurun_out/r2l_bench_2gpu.json 2> gpurun_out/r2l_bench_2gpu.err
tail -5 gpurun_out/r2l_bench_2gpu.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/r2l_bench_2gpu.json').read().strip().splitlines() if l.startswith('{')][-1])
    print('value %.1f' % d['value'], 'ms %.4f' % d['ms_per_step'], d['grid'])
    print('e2e', d['e2e']); print('parity', d['parity'])
    print('kernels', {k: v['ms'] for k, v in d['kernels'].items()})
    for k, v in (d['other_configs'] or {}).items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('error'))
except Exception as e: print('FAILED', e, open('gpurun_out/r2l_bench_2gpu.json').read()[-1500:])
PY
