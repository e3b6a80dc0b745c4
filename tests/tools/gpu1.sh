set -x
python __graft_entry__.py smoke 2>&1 | tail -3
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python tests/gpu_report.py 1000000 2>&1 | tail -12
python bench.py --steps 200 --warmup 10 2>&1 | tail -3 > gpurun_out/bench_first.log; cat gpurun_out/bench_first.log
python bench.py --impl reference --steps 20 --warmup 1 2>&1 | tail -2
