# final single-GPU evidence for this state: GPU tests, smoke, default bench (roofline + cpu_baseline + e2e), reference arm, ncu launch list
export FITSNE_BENCH_CACHE=/tmp/fitsne_cache
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r1_pytest_gpu.txt; cat gpurun_out/r1_pytest_gpu.txt
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r1_smoke.txt
timeout 300 python bench.py 2>&1 | tail -1 > gpurun_out/bench_r1_v5.json; cut -c1-400 gpurun_out/bench_r1_v5.json
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 110 --csv --log-file gpurun_out/launches_late_v5.csv python tests/tools/profile_steps.py 1000000 late 4 > /dev/null 2>&1
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1_v5.json')); print('value %.1f e2e %.1f cpu %s'%(d['value'], d['e2e']['value'], d['cpu_baseline'])); print(d['kernels'])
PY
