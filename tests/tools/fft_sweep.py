import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fit-sne_b200"))
import bench, fitsne_b200 as fb
row, col, val, Y0, sched = bench.workload(1000000, "late")
with fb.FitSNE(row, col, val, Y0, flags=fb.FLAG_TIMERS) as t:
    for _ in range(5): t.step(exaggeration=1.0, momentum=0.8, learning_rate=1e6/12, max_step_norm=5.0)
    t.synchronize(); t.reset_stats()
    for _ in range(50): t.step(exaggeration=1.0, momentum=0.8, learning_rate=1e6/12, max_step_norm=5.0)
    st = t.stats()
    print("threads=%s lc=%s lr=%s: fft %.4f ms (M=%d)" % (os.environ.get("FITSNE_FFT_THREADS"), os.environ.get("FITSNE_FFT_LINES_COLS"), os.environ.get("FITSNE_FFT_LINES_ROWS"), st["phase_ms"]["fft"]/50, st["fft_side"]), flush=True)
