"""Not a test: prints a parity / timing report on a GPU box.  python tests/gpu_report.py [N]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "fit-sne_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import fitsne_b200 as fb  # noqa: E402
import bench_util  # noqa: E402


def rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def parity():
    g = np.load(os.path.join(ROOT, "tests/golden/gradients_n3000.npz"))
    gg = np.load(os.path.join(ROOT, "tests/golden/graph_n3000.npz"))
    row, col, val = gg["row"], gg["col"], gg["val"].astype(np.float64)
    for name in sorted(set(k.split("__")[0] for k in g.files)):
        dims, df, nterms, ipi, min_int, Z, kl = g[name + "__meta"]
        Y = g[name + "__Y"].astype(np.float64)
        try:
            with fb.FitSNE(row, col, val, Y, nterms=int(nterms), intervals_per_integer=ipi, min_num_intervals=int(min_int), df=df) as t:
                dC, z = t.gradient(1.0)
                frep = t.debug("frep", np.float32).reshape(len(Y), -1)
                k = t.kl(1.0)
                st = t.stats()
            print("%-14s dC %.2e  rep %.2e  Z %.2e  KL %.2e  B=%d M=%d" % (name, rel(dC, g[name + "__dC"]), rel(-frep, g[name + "__dC_rep"]),
                                                                        abs(z - Z) / Z, abs(k - kl) / abs(kl), st["n_boxes"], st["fft_side"]))
        except Exception as e:  # noqa
            print("%-14s FAILED %s" % (name, e))


def timing(N, K=15):
    t0 = time.time()
    row, col, val, labels = bench_util.knn_like_graph(N, K)
    print("graph N=%d E=%d built in %.1fs" % (N, len(col), time.time() - t0))
    for phase, Y0, alpha in (("early", bench_util.early_embedding(N, 2), 12.0), ("late", bench_util.clustered_embedding(labels, 2, 170.0), 1.0)):
        for flags, label in ((0, "graph"), (fb.FLAG_TIMERS, "timers")):
            with fb.FitSNE(row, col, val, Y0, flags=flags) as t:
                for _ in range(5):
                    t.step(exaggeration=alpha, momentum=0.5, learning_rate=N / 12.0, max_step_norm=5.0)
                t.synchronize()
                t.reset_stats()
                n = 50
                t1 = time.time()
                for _ in range(n):
                    t.step(exaggeration=alpha, momentum=0.5, learning_rate=N / 12.0, max_step_norm=5.0)
                t.synchronize()
                dt = time.time() - t1
                st = t.stats()
                print("%s/%s: %.1f it/s (%.3f ms/it) B=%d G=%d M=%d launches/it=%.1f regrids=%d" % (
                    phase, label, n / dt, dt / n * 1e3, st["n_boxes"], st["grid_side"], st["fft_side"], st["kernel_launches"] / n, st["regrids"]))
                if flags:
                    print("   phases ms/it:", {k: round(v / n, 4) for k, v in st["phase_ms"].items() if v})


if __name__ == "__main__":
    print(fb.load_library().fitsne_version().decode())
    parity()
    timing(int(sys.argv[1]) if len(sys.argv) > 1 else 1000000)
