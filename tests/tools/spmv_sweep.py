import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fit-sne_b200"))
import bench, fitsne_b200 as fb
row, col, val, Y0, sched = bench.workload(1000000, "late")
for flags, label in ((0, "graph"), (fb.FLAG_TIMERS, "timers")):
    with fb.FitSNE(row, col, val, Y0, flags=flags) as t:
        for _ in range(5): t.step(exaggeration=1.0, momentum=0.8, learning_rate=1e6/12, max_step_norm=5.0)
        t.synchronize(); t.reset_stats(); t0 = time.time()
        for _ in range(100): t.step(exaggeration=1.0, momentum=0.8, learning_rate=1e6/12, max_step_norm=5.0)
        t.synchronize(); dt = time.time() - t0
        st = t.stats()
        print("per_sm=%s smem=%s %s: %.3f ms/it  attract %.4f" % (os.environ.get("FITSNE_SPMV_CTAS_PER_SM"), os.environ.get("FITSNE_SPMV_SMEM_KB"), label, dt*10, st["phase_ms"]["attract_update"]/100), flush=True)
