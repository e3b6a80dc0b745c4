export FITSNE_BENCH_CACHE=/tmp/fitsne_cache
A="--points 10000000 --steps 100 --warmup 10 --no-extras --no-e2e --no-cpu-baseline --no-parity"
timeout 600 python bench.py --gpus 1 $A 2>&1 | tail -1 > gpurun_out/r2m_10M_1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 $A 2>&1 | grep '^{' | tail -1 > gpurun_out/r2m_10M_2.json
python - <<'PY'
import json
for n in (1,2):
    try:
        d = json.load(open('gpurun_out/r2m_10M_%d.json'%n)); print(n, 'value %.1f' % d['value'], 'ms %.4f' % d['ms_per_step'], d['grid'], {k: v['ms'] for k, v in d['kernels'].items()})
    except Exception as e: print(n, 'FAILED', e, open('gpurun_out/r2m_10M_%d.json'%n).read()[-800:])
PY
