export FITSNE_BENCH_CACHE=/tmp/fitsne_cache
A="--steps 100 --warmup 10 --no-extras --no-e2e --no-cpu-baseline --no-parity"
run() { tag=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 $A 2>&1 | grep '^{' | tail -1 > gpurun_out/r2s_$tag.json; }
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 tests/tools/mgpu_check.py 2>&1 | grep -v "^\[fitsne" | tail -6
run 1M_dist FITSNE_DIST_CONV=1
run 1M_repl FITSNE_DIST_CONV=0
A="$A --points 10000000"
run 10M_dist FITSNE_DIST_CONV=1
run 10M_repl FITSNE_DIST_CONV=0
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2s_*.json')):
    try:
        d = json.load(open(f)); print(f, 'value %.1f' % d['value'], 'ms %.4f' % d['ms_per_step'], {k: v['ms'] for k, v in d['kernels'].items()})
    except Exception as e: print(f, 'FAILED', e, open(f).read()[-300:])
PY
