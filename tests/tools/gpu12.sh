for P in 10000 70000; do for ph in early late; do
python bench.py --points $P --phase $ph --steps 1000 --warmup 20 --no-e2e 2>&1 | tail -1 > gpurun_out/small_${P}_${ph}.json
done; done
python bench.py --points 1000000 --dims 1 --df 0.5 --steps 300 --warmup 10 2>&1 | tail -1 > gpurun_out/cfg5_1d.json
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/small_*.json'))+['gpurun_out/cfg5_1d.json']:
    try:
        d = json.load(open(f)); print(f, 'value %.1f it/s' % d['value'], 'ms/step %.4f' % d['ms_per_step'], d['grid'], 'cpu', d['cpu_baseline'] and d['cpu_baseline']['value'], 'e2e', d['e2e'] and d['e2e']['value'], 'kl', d['kl_last'])
    except Exception as e: print(f, 'FAILED', open(f).read()[-600:])
PY
