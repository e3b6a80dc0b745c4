# round-1 evidence for HEAD: GPU tests, smoke, default bench, ncu launch list + one full capture of the top kernels
set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r1_pytest_gpu.txt; cat gpurun_out/r1_pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r1_smoke.txt
timeout 400 python bench.py 2>&1 | tail -1 > gpurun_out/bench_r1_v4.json; cut -c1-1500 gpurun_out/bench_r1_v4.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_late_v4.csv python tests/tools/profile_steps.py 1000000 late 4 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_attract|k_fft_pass|k_hadamard" -s 8 -c 7 -o gpurun_out/prof_r1_v4 python tests/tools/profile_steps.py 1000000 late 4 > /dev/null 2>&1
ls -la gpurun_out | tail -6
