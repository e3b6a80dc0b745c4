# single GPU: parity after the padding-free FFT input + cheaper last-block epilogues, then the 1M bench
set -x
export FITSNE_BENCH_CACHE=/tmp/fitsne_cache
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python bench.py --steps 500 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_v5_500.json
timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-e2e 2>&1 | tail -1 > gpurun_out/bench_v5_200.json
python - <<'PY'
import json
for f in ('gpurun_out/bench_v5_500.json','gpurun_out/bench_v5_200.json'):
    try:
        d=json.load(open(f)); print(f, 'value %.1f'%d['value'], 'ms %.4f'%d['ms_per_step'], d['grid'], 'e2e', d['e2e'] and d['e2e']['value'], {k:v['ms'] for k,v in d['kernels'].items()})
    except Exception as e: print(f,'FAILED',open(f).read()[-1500:])
PY
