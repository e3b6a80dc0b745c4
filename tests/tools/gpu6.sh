python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tests/gpu_report.py 1000000 2>&1 | tail -20
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_late_v2.csv python tests/tools/profile_steps.py 1000000 late 4 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_fft_cols|k_fft_rows|k_attract|k_hadamard" -s 12 -c 8 -o gpurun_out/prof_fft_attract python tests/tools/profile_steps.py 1000000 late 4 > /dev/null 2>&1
ls -la gpurun_out/
