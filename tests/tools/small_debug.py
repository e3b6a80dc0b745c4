import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fit-sne_b200"))
import bench, fitsne_b200 as fb
which = sys.argv[1]
row, col, val, Y0, sched = bench.workload(10000, "late")
if which == "a":
    t = fb.FitSNE(row, col, val, Y0); t.run(fetch_Y=False, max_iter=100, **sched); t.close()
elif which == "b":
    t = fb.FitSNE(row, col, val, Y0, flags=fb.FLAG_TIMERS)
    for _ in range(10): t.step(exaggeration=1.0, momentum=0.8, learning_rate=800.0, max_step_norm=5.0)
    print(t.stats()["phase_ms"]); t.close()
elif which == "c":
    Y, c = fb.run_host(row, col, val, Y0, max_iter=100, **sched)
elif which == "d":
    t = fb.FitSNE(row, col, val, Y0); t.run(fetch_Y=False, max_iter=100, **sched); t.prewarm(115, 265); t.close()
elif which == "e":
    t = fb.FitSNE(row, col, val, Y0, flags=fb.FLAG_NO_REORDER); t.run(fetch_Y=False, max_iter=100, **sched); t.close()
print("done", which)
