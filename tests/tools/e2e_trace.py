"""Not a test: where the end-to-end call's time goes (FITSNE_TRACE=1 python tests/tools/e2e_trace.py [points] [steps])."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fit-sne_b200"))
import numpy as np
import torch
import bench, fitsne_b200 as fb
points = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
row, col, val, Y0, sched = bench.workload(points, "late")
pinned = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (row, col, val, Y0)]
prow, pcol, pval, pY = [p.numpy() for p in pinned]
torch.cuda.init(); torch.zeros(1, device="cuda"); torch.cuda.synchronize()
for rep in range(3):
    print("---- call %d" % rep, file=sys.stderr, flush=True)
    t0 = time.perf_counter()
    Y, costs = fb.run_host(prow, pcol, pval, pY.copy(), max_iter=steps, **sched)
    dt = time.perf_counter() - t0
    print("call %d: %.1f ms -> %.0f it/s" % (rep, dt * 1e3, steps / dt), file=sys.stderr, flush=True)
