export FITSNE_BENCH_CACHE=/tmp/fitsne_cache
A="--points 10000000 --steps 100 --warmup 10 --no-extras --no-e2e --no-cpu-baseline --no-parity"
run() { tag=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 $A 2>&1 | grep '^{' | tail -1 > gpurun_out/r2n_$tag.json; }
run base A=1
run serial FITSNE_SERIAL=1
run agstream FITSNE_AG_STREAM=1
run sync FITSNE_SHARDED_SYNC=1
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2n_*.json')):
    try:
        d = json.load(open(f)); print(f, 'value %.1f' % d['value'], 'ms %.4f' % d['ms_per_step'], {k: v['ms'] for k, v in d['kernels'].items()})
    except Exception as e: print(f, 'FAILED', e, open(f).read()[-300:])
PY
