"""How often does the kernel-spectrum cache hit in a realistic full run (early exaggeration then relaxation)?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fit-sne_b200"))
import bench_util, fitsne_b200 as fb
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
row, col, val, labels = bench_util.knn_like_graph(N, 15)
Y0 = bench_util.early_embedding(N, 2)
with fb.FitSNE(row, col, val, Y0) as t:
    done = 0
    for seg, kw in ((250, dict(early_exag_coeff=12.0, stop_lying_iter=10**9, mom_switch_iter=10**9, momentum=0.5)),
                    (250, dict(early_exag_coeff=1.0, stop_lying_iter=-1, mom_switch_iter=-1, momentum=0.8)),
                    (250, dict(early_exag_coeff=1.0, stop_lying_iter=-1, mom_switch_iter=-1, momentum=0.8))):
        t.reset_stats()
        t0 = time.time()
        Y, costs = t.run(fetch_Y=False, max_iter=seg, learning_rate=N / 12.0, max_step_norm=5.0, **kw)
        st = t.stats()
        done += seg
        print("iters %d-%d: %.1f it/s device, cache hits %d/%d, regrids %d, B=%d M=%d KL %.4f" % (done - seg, done, seg / (t.last_run_ms() * 1e-3),
              st["spectrum_cache_hits"], seg, st["regrids"], st["n_boxes"], st["fft_side"], costs[costs != 0][-1]), flush=True)
