export FITSNE_BENCH_CACHE=/tmp/fitsne_cache
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
A="--steps 100 --warmup 10 --no-extras --no-e2e --no-cpu-baseline --no-parity"
timeout 600 python bench.py --gpus 1 $A 2>&1 | tail -1 > gpurun_out/r2o_1M_1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 $A 2>&1 | grep '^{' | tail -1 > gpurun_out/r2o_1M_2.json
timeout 600 python bench.py --gpus 1 --points 10000000 $A 2>&1 | tail -1 > gpurun_out/r2o_10M_1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --points 10000000 $A 2>&1 | grep '^{' | tail -1 > gpurun_out/r2o_10M_2.json
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2o_*.json')):
    try:
        d = json.load(open(f)); print(f, 'value %.1f' % d['value'], 'ms %.4f' % d['ms_per_step'], d['graph_launches'], {k: v['ms'] for k, v in d['kernels'].items()})
    except Exception as e: print(f, 'FAILED', e, open(f).read()[-300:])
PY
