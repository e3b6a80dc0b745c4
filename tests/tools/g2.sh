export FITSNE_BENCH_CACHE=/tmp/fitsne_cache
for P in 2 3 4 5 6; do
  FITSNE_SPMV_CTAS_PER_SM=$P timeout 200 python bench.py --steps 300 --no-cpu-baseline --no-e2e 2>&1 | tail -1 > gpurun_out/r2b_bench_persm_$P.json
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2b_bench_persm_*.json')):
    try:
        d = json.load(open(f)); print(f, 'value %.1f' % d['value'], 'ms %.4f' % d['ms_per_step'], d['grid'], {k: v['ms'] for k, v in d['kernels'].items()})
    except Exception as e: print(f, 'FAILED', open(f).read()[-600:])
PY
