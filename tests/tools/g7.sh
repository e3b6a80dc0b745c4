export FITSNE_BENCH_CACHE=/tmp/fitsne_cache
python -c "import bench; bench.workload(1000000,'late')" > /dev/null 2>&1
timeout 800 ncu --set full --clock-control none --import-source on --launch-skip 1650 -c 16 -o gpurun_out/r2_full_iter python tests/tools/profile_steps.py 1000000 late 120 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
