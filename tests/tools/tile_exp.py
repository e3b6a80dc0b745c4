import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fit-sne_b200"))
import bench, fitsne_b200 as fb
row, col, val, Y0, sched = bench.workload(1000000, "late")
with fb.FitSNE(row, col, val, Y0, flags=fb.FLAG_TIMERS | fb.FLAG_FORCE_TILES) as t:
    for _ in range(3): t.step(exaggeration=1.0, momentum=0.8, learning_rate=1e6/12, max_step_norm=5.0)
    t.reset_stats()
    for _ in range(20): t.step(exaggeration=1.0, momentum=0.8, learning_rate=1e6/12, max_step_norm=5.0)
    print("ACC", os.environ.get("FITSNE_TILE_ACC"), "attract ms", t.stats()["phase_ms"]["attract_update"]/20)
