"""Not a test: spread / sort kernel times against the spread chunk length for small slices (python tests/tools/chunk_sweep.py)."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fit-sne_b200"))
if len(sys.argv) > 2:
    import bench, fitsne_b200 as fb
    points = int(sys.argv[1])
    row, col, val, Y0, sched = bench.workload(points, "late")
    kw = dict(exaggeration=sched["early_exag_coeff"], momentum=sched["momentum"], learning_rate=sched["learning_rate"], max_step_norm=5.0)
    with fb.FitSNE(row, col, val, Y0, flags=fb.FLAG_TIMERS) as t:
        for _ in range(40): t.step(**kw)
        t.reset_stats()
        for _ in range(60): t.step(**kw)
        kt = t.kernel_times(); st = t.stats()
    print("N=%d chunk=%s M=%d: " % (points, os.environ.get("FITSNE_CHUNK", "auto"), st["fft_side"]) +
          "  ".join("%s %.1f" % (k.replace("k_", ""), 1e3 * v[0] / max(v[1], 1)) for k, v in kt.items() if "spread" in k or "sweep" in k or "bin" in k or "gather" in k), flush=True)
else:
    for n in (125000, 250000, 500000, 1000000):
        for ch in ("1", "2", "4", "8"):
            env = dict(os.environ, FITSNE_KTIMES="1", FITSNE_CHUNK=ch)
            subprocess.run([sys.executable, __file__, str(n), "x"], env=env, timeout=600)
