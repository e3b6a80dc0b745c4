# 2 GPUs: sharded path (Y all-gather on the SpMV stream, stats exchange, plain-launch batches) vs single GPU; 1M bench in both modes
set -x
export FITSNE_BENCH_CACHE=/tmp/fitsne_cache
NG=$(nvidia-smi -L | wc -l)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 tests/tools/mgpu_check.py 2>&1 | grep -v "^W\|^\*\*\*" | tail -8
FITSNE_SHARDED_SYNC=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29518 tests/tools/mgpu_check.py 2>&1 | grep -v "^W\|^\*\*\*" | tail -4
P=1000000
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $NG --points $P --steps 300 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/s5_${P}_${NG}_batched.json
FITSNE_SHARDED_SYNC=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus $NG --points $P --steps 300 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/s5_${P}_${NG}_sync.json
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/s5_*.json')):
    try:
        d = json.load(open(f)); print(f, d['n_gpus'], 'value %.1f it/s' % d['value'], 'ms/step %.3f' % d['ms_per_step'], d['grid'], 'launches', d['gpu_launches'], 'kl', d['kl_last'])
    except Exception as e: print(f, 'FAILED', open(f).read()[-1500:])
PY
