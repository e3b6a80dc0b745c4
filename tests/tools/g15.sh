export FITSNE_BENCH_CACHE=/tmp/fitsne_cache
FITSNE_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/tools/mgpu_check.py > gpurun_out/mg.log 2>&1
echo rc=$?; grep -v "^\[fitsne" gpurun_out/mg.log | tail -12; grep "peer fabric" gpurun_out/mg.log | head -3
A="--steps 100 --warmup 10 --no-extras --no-e2e --no-cpu-baseline"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 $A 2>&1 | grep '^{' | tail -1 > gpurun_out/r2p_1M_2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --points 10000000 $A 2>&1 | grep '^{' | tail -1 > gpurun_out/r2p_10M_2.json
FITSNE_NO_P2P=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --points 10000000 $A 2>&1 | grep '^{' | tail -1 > gpurun_out/r2p_10M_2_nccl.json
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2p_*.json')):
    try:
        d = json.load(open(f)); print(f, 'value %.1f' % d['value'], 'ms %.4f' % d['ms_per_step'], d['graph_launches'], (d.get('parity') or {}).get('gradient_rel_l2'), {k: v['ms'] for k, v in d['kernels'].items()})
    except Exception as e: print(f, 'FAILED', e, open(f).read()[-300:])
PY
