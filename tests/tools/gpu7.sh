ncu --set full --clock-control none --import-source on -k regex:"k_fft_pass" -s 8 -c 4 -o gpurun_out/prof_fft2 python tests/tools/profile_steps.py 1000000 late 4 > /dev/null 2>&1
ls -la gpurun_out/
