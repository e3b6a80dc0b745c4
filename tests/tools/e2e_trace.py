import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fit-sne_b200"))
import bench, fitsne_b200 as fb
row, col, val, Y0, sched = bench.workload(1000000, "late")
fb.run_host(row, col, val, Y0, max_iter=20, **sched)   # warm the process (context, module load)
t0 = time.perf_counter()
Y, c = fb.run_host(row, col, val, Y0, max_iter=500, **sched)
print("e2e 500 steps: %.1f ms -> %.1f it/s" % ((time.perf_counter() - t0) * 1e3, 500 / (time.perf_counter() - t0)))
