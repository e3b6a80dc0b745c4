# 2 GPUs: full sharded-vs-single check, then A/B of the sharded iteration at 5M points: this build vs the previous commit's library
export FITSNE_BENCH_CACHE=/tmp/fitsne_cache
timeout -k 5 40 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/tools/mgpu_check.py 2>&1 | grep -E "dims=|MGPU_OK|free\(\)" | tee gpurun_out/r1_mgpu_check.txt
P=5000000
timeout -k 5 130 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --points $P --steps 150 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/ab_${P}_2_new.json
FITSNE_LIB=$PWD/fit-sne_b200/lib/old/libfitsne_b200.so timeout -k 5 70 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 2 --points $P --steps 150 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/ab_${P}_2_old.json
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/ab_*.json')):
    try:
        d = json.load(open(f)); print(f, d['n_gpus'], 'value %.1f it/s' % d['value'], 'ms/step %.3f' % d['ms_per_step'], d['grid'], 'launches', d['gpu_launches'], 'kl', d['kl_last'])
    except Exception as e: print(f, 'FAILED', open(f).read()[-800:])
PY
