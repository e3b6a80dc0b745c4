import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fit-sne_b200"))
import bench_util, fitsne_b200 as fb
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200003
row, col, val, labels = bench_util.knn_like_graph(N, 8, seed=3)
for dims, df, span in ((2, 1.0, 60.0), (1, 0.5, 120.0)):
    Y0 = bench_util.clustered_embedding(labels, dims, span, seed=5)
    sched = dict(max_iter=60, stop_lying_iter=20, mom_switch_iter=20, learning_rate=500.0, early_exag_coeff=4.0)
    with fb.FitSNE(row, col, val, Y0, df=df) as s:
        dC1, Z1 = s.gradient(4.0)
        kl1 = s.kl(4.0)
        Y1, costs1 = s.run(**sched)
        print("case", dims, df, "ok", costs1[costs1 != 0], s.stats()["regrids"], flush=True)
print("done")
