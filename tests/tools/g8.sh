export FITSNE_BENCH_CACHE=/tmp/fitsne_cache
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
run() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 300 --no-cpu-baseline --no-e2e 2>&1 | tail -1 > gpurun_out/r2i_$tag.json; }
run base A=1
run base2 A=1
run col128 FITSNE_COL_THREADS=128
run col192 FITSNE_COL_THREADS=192
run tex FITSNE_SPMV_TEX=1
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2i_*.json')):
    try:
        d = json.load(open(f)); print(f, 'value %.1f' % d['value'], 'ms %.4f' % d['ms_per_step'], d['grid'], {k: v['ms'] for k, v in d['kernels'].items()})
    except Exception as e: print(f, 'FAILED', open(f).read()[-600:])
PY
