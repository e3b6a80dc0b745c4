export FITSNE_BENCH_CACHE=/tmp/fitsne_cache
python -c "import bench; bench.workload(1000000,'late')" > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 120 --csv --log-file gpurun_out/r2_launches_warm.csv python tests/tools/profile_steps.py 1000000 late 6 > /dev/null 2>&1
