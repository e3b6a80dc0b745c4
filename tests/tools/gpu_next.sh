# First GPU call of the next round (1 GPU, ~3 min): validate the end-of-round-1 state, A/B the opt-in kernels, profile warm.
#   gpurun --timeout 600 -- 'bash tests/tools/gpu_next.sh'
set -x
export FITSNE_BENCH_CACHE=/tmp/fitsne_cache
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/next_pytest_gpu.txt
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/next_smoke.txt
timeout 120 python tests/tools/oneshot.py 2>&1 | tail -70 | tee gpurun_out/next_oneshot.txt          # incl. section D: column-sorted SpMV
timeout 300 python bench.py 2>&1 | tail -1 > gpurun_out/next_bench.json
# packed kernel planes / sorted SpMV in the real benchmark regime (no spectrum-cache hits, real graph, re-ordering on)
for F in 0 1024 2048 3072; do
  FITSNE_FLAGS=$F timeout 200 python bench.py --steps 300 --no-cpu-baseline --no-e2e 2>&1 | tail -1 > gpurun_out/next_bench_flags_$F.json
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/next_bench*.json')):
    try:
        d = json.load(open(f)); print(f, 'value %.1f' % d['value'], 'ms %.4f' % d['ms_per_step'], d['grid'], {k: v['ms'] for k, v in d['kernels'].items()})
    except Exception as e: print(f, 'FAILED', open(f).read()[-600:])
PY
# warm per-kernel times (no cache flush between replays) + the usual cold launch list
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 110 --csv --log-file gpurun_out/next_launches_warm.csv python tests/tools/profile_steps.py 1000000 late 4 > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 110 --csv --log-file gpurun_out/next_launches_cold.csv python tests/tools/profile_steps.py 1000000 late 4 > /dev/null 2>&1
ls -la gpurun_out | tail -12
# then, on 8 GPUs (charged 8x: keep it to the 10M scaling line, generation ~2 min):
#   gpurun --gpus 8 --timeout 900 -- 'FITSNE_BENCH_CACHE=/tmp/fc python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --points 10000000 --steps 200 --warmup 10 --no-cpu-baseline | tail -1'
#   (A/B knobs: FITSNE_SHARDED_SYNC=1 = one host round trip per iteration; FITSNE_AG_STREAM=1 = all-gather on its own stream)
