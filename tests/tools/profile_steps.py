"""Not a test: a short fixed workload for ncu (python tests/tools/profile_steps.py [points] [phase] [steps])."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fit-sne_b200"))
import bench  # noqa
import fitsne_b200 as fb  # noqa
points = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
phase = sys.argv[2] if len(sys.argv) > 2 else "late"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
row, col, val, Y0, sched = bench.workload(points, phase)
with fb.FitSNE(row, col, val, Y0) as t:
    t.run(fetch_Y=False, max_iter=steps, **sched)
    print(t.stats())
