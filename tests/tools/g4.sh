export FITSNE_BENCH_CACHE=/tmp/fitsne_cache
python -c "import bench; bench.workload(1000000,'late')" > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --launch-skip 3300 -c 80 --csv --log-file gpurun_out/r2_launches_warm_late.csv python tests/tools/profile_steps.py 1000000 late 230 > /dev/null 2>&1
tail -2 gpurun_out/r2_launches_warm_late.csv | cut -c1-200
