set -x
export FITSNE_BENCH_CACHE=/tmp/fitsne_cache
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2a_pytest_gpu.txt
timeout 150 python tests/tools/oneshot.py 2>&1 | tail -80 > gpurun_out/r2a_oneshot.txt
for F in 0 2048; do
  FITSNE_FLAGS=$F timeout 200 python bench.py --steps 300 --no-cpu-baseline --no-e2e 2>&1 | tail -1 > gpurun_out/r2a_bench_flags_$F.json
done
tail -3 gpurun_out/r2a_oneshot.txt
