export FITSNE_BENCH_CACHE=/tmp/fitsne_cache
run() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 300 --no-cpu-baseline --no-e2e 2>&1 | tail -1 > gpurun_out/r2e_$tag.json; }
run base A=1
run serial FITSNE_SERIAL=1
run persm16 FITSNE_SPMV_CTAS_PER_SM=16
run persm64 FITSNE_SPMV_CTAS_PER_SM=64
run persmall FITSNE_SPMV_CTAS_PER_SM=100000
run sorted FITSNE_FLAGS=2048
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2e_*.json')):
    try:
        d = json.load(open(f)); print(f, 'value %.1f' % d['value'], 'ms %.4f' % d['ms_per_step'], d['grid'], {k: v['ms'] for k, v in d['kernels'].items()})
    except Exception as e: print(f, 'FAILED', open(f).read()[-600:])
PY
