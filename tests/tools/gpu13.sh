python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -3
NG=$(nvidia-smi -L | wc -l)
P=10000000
python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $NG --points $P --steps 200 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/scale2_${P}_${NG}.json
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/scale2_*.json')):
    try:
        d = json.load(open(f)); print(f, d['n_gpus'], 'value %.1f it/s' % d['value'], 'ms/step %.3f' % d['ms_per_step'], d['grid'], 'launches', d['gpu_launches'])
    except Exception as e: print(f, 'FAILED', open(f).read()[-1500:])
PY
