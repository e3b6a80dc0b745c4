for ph in late early; do ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_${ph}_v3.csv python tests/tools/profile_steps.py 1000000 $ph 4 > /dev/null 2>&1; done
ncu --set full --clock-control none --import-source on -k regex:"k_attract|k_spread_chunks|k_gather|k_radix_scatter" -s 10 -c 5 -o gpurun_out/prof_r1_top python tests/tools/profile_steps.py 1000000 late 4 > /dev/null 2>&1
ls -la gpurun_out | tail -5
