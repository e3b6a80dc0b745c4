python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -15
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29600 tests/tools/mgpu_check.py 2>&1 | grep -E "dims=|MGPU|rror" | head
python bench.py --steps 500 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_1gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 2 --steps 500 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_2gpu.json
python - <<'PY'
import json
for f in ('gpurun_out/bench_1gpu.json','gpurun_out/bench_2gpu.json'):
    try:
        d = json.load(open(f))
        print(f, {k: d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches','regrids','grid','clocks')})
        print('   e2e', d['e2e'])
    except Exception as e:
        print(f, 'FAILED', e, open(f).read()[-1500:])
PY
