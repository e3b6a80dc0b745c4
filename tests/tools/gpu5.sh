python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python tests/gpu_report.py 1000000 2>&1 | tail -22
