"""One short GPU shot (no torch import; finishes in seconds): A/B of alternative kernels against the shipped path.
  FLAG_SPREAD_PER_NODE -- first spread formulation (thread per chunk and node)   FLAG_FFT_WIDE -- radix-16/9 FFT plans
  FLAG_KPACK -- four kernel planes in one complex transform                      FLAG_FUSED_COLSUM -- column sums inside k_update
(Round 1 ran this with the then-opt-in kernels as variants: profiles/r1_oneshot_ab.json; the per-chunk spread and the
separate column-sum pass became the defaults as a result, so their flags now select the OLD variants.)
For each: gradient parity with the shipped kernels (and with the oracle on the small cases) + per-phase device times.
Writes gpurun_out/oneshot.json incrementally (the call may be cut short)."""
import json, os, sys, time
T0 = time.time()
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fit-sne_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import fitsne_b200 as fb
OUT = {"log": []}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
def save():
    with open(os.path.join(ROOT, "gpurun_out", "oneshot.json"), "w") as f:
        json.dump(OUT, f, indent=1)
def say(msg):
    line = "[%5.1fs] %s" % (time.time() - T0, msg)
    print(line, flush=True); OUT["log"].append(line); save()

def ring_graph(N):
    i = np.arange(N, dtype=np.int64)
    row = (2 * np.arange(N + 1)).astype(np.uint32)
    col = np.empty(2 * N, np.uint32); col[0::2] = (i - 1) % N; col[1::2] = (i + 1) % N
    return row, col, np.full(2 * N, 1.0 / (2 * N))
def blobs(N, dims, span, seed, spread=0.03):
    rng = np.random.default_rng(seed)
    lab = rng.integers(0, 10, N)
    cen = rng.uniform(-0.5, 0.5, (10, dims)) * span
    Y = cen[lab] + rng.standard_normal((N, dims)) * span * spread
    return (Y - Y.mean(0)).astype(np.float32).astype(np.float64)
def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))

ALL = fb.FLAG_FFT_WIDE | fb.FLAG_KPACK
VARIANTS = (("spread_per_node", fb.FLAG_SPREAD_PER_NODE), ("fftwide", fb.FLAG_FFT_WIDE), ("kpack", fb.FLAG_KPACK), ("both", ALL))
BASEF = fb.FLAG_NO_REORDER

def grad(row, col, val, Y, flags, **kw):
    with fb.FitSNE(row, col, val, Y, flags=BASEF | flags, **kw) as t:
        dC, Z = t.gradient(1.0)
        st = t.stats()
    return dC, Z, st

def timed(row, col, val, Y, flags, steps=20, ktimes=False, **kw):
    os.environ["FITSNE_KTIMES"] = "1" if ktimes else "0"          # read at context creation: warm per-kernel times
    with fb.FitSNE(row, col, val, Y, flags=BASEF | fb.FLAG_TIMERS | flags, **kw) as t:
        for _ in range(3):
            t.step(exaggeration=1.0, momentum=0.8, learning_rate=1000.0, max_step_norm=5.0)
        t.reset_stats()
        for _ in range(steps):
            t.step(exaggeration=1.0, momentum=0.8, learning_rate=1000.0, max_step_norm=5.0)
        st = t.stats()
        if ktimes:
            kt = t.kernel_times()
            OUT["kernel_us_warm"] = {k: round(1e3 * ms / max(n, 1), 2) for k, (ms, n) in kt.items()}
            say("warm per-kernel us: " + "  ".join("%s %.1f" % kv for kv in sorted(OUT["kernel_us_warm"].items(), key=lambda kv: -kv[1])))
    return {k: round(v / steps, 5) for k, v in st["phase_ms"].items() if v > 0}, st["fft_side"]

try:
    # ---- A: the benchmark size (spread / FFT phases are what matters; the graph is a trivial ring)
    N = 1000000
    row, col, val = ring_graph(N)
    Y = blobs(N, 2, 170.0, 1)
    say("data ready")
    base_dC, base_Z, st = grad(row, col, val, Y, 0)
    say("1M base gradient: Z %.6e, grid %s/%s" % (base_Z, st["n_boxes"], st["fft_side"]))
    OUT["A"] = {}
    for name, fl in VARIANTS:
        dC, Z, _ = grad(row, col, val, Y, fl)
        OUT["A"][name] = {"dC_rel": rel(dC, base_dC), "Z_rel": abs(Z - base_Z) / base_Z, "bitwise": bool(np.array_equal(dC, base_dC))}
        say("1M %-8s vs shipped: dC rel %.2e  Z rel %.2e  bitwise %s" % (name, OUT["A"][name]["dC_rel"], OUT["A"][name]["Z_rel"], OUT["A"][name]["bitwise"]))
    OUT["A_ms"] = {}
    for name, fl in (("shipped", 0), ("spread_per_node", fb.FLAG_SPREAD_PER_NODE), ("fftwide", fb.FLAG_FFT_WIDE), ("kpack", fb.FLAG_KPACK), ("all", ALL),
                     ("fused_colsum", fb.FLAG_FUSED_COLSUM)):
        ms, M = timed(row, col, val, Y, fl)
        OUT["A_ms"][name] = ms
        if name == "shipped":
            timed(row, col, val, Y, fl, ktimes=True)
        say("1M %-8s per-step ms (M=%d): %s" % (name, M, ms))
    # ---- B: small cases against the oracle too
    from pyoracle import Oracle
    O = Oracle()
    n = 20000
    row, col, val = ring_graph(n)
    OUT["B"] = {}
    for tag, dims, df, p, span in (("2d_p3_late", 2, 1.0, 3, 75.0), ("2d_p3_early", 2, 1.0, 3, 3e-4), ("2d_p4", 2, 1.0, 4, 60.0),
                                   ("2d_p2", 2, 1.0, 2, 80.0), ("1d_df05", 1, 0.5, 3, 150.0), ("1d_p5", 1, 1.0, 5, 90.0), ("2d_df05", 2, 0.5, 3, 70.0)):
        Yc = blobs(n, dims, span, 7)
        ref, Zr = O.gradient(Yc, row, col, val, nterms=p, df=df)
        res = {}
        for name, fl in (("shipped", 0),) + VARIANTS:
            dC, Z, st = grad(row, col, val, Yc, fl, nterms=p, df=df)
            res[name] = {"vs_oracle": rel(dC, ref), "Z_rel": abs(Z - Zr) / Zr, "M": st["fft_side"]}
        OUT["B"][tag] = res
        say("%-12s M=%-5d dC vs oracle: %s" % (tag, res["both"]["M"], "  ".join("%s %.1e" % (k, v["vs_oracle"]) for k, v in res.items())))
        say("%-12s         Z  vs oracle: %s" % (tag, "  ".join("%s %.1e" % (k, v["Z_rel"]) for k, v in res.items())))
    # ---- C: FFT lengths across the n_boxes ladder (fp32 vs fp32: the two builds round differently, ~1e-6 expected)
    OUT["C"] = {}
    for span in (40, 53, 58, 63, 68, 73, 78, 83, 88, 94, 98, 105, 115, 125, 135, 145, 160, 190, 230, 270, 330, 420, 600):
        Yc = blobs(n, 2, float(span) / 1.05, 11, spread=0.05)
        Yc *= span / (Yc.max() - Yc.min())
        a, Za, st = grad(row, col, val, Yc, 0)
        b, Zb, _ = grad(row, col, val, Yc, ALL)
        OUT["C"][str(st["fft_side"])] = {"n_boxes": st["n_boxes"], "both_vs_shipped": rel(b, a), "Z_rel": abs(Zb - Za) / Za}
        say("span %-4d B=%-4d M=%-5d both vs shipped: dC rel %.2e  Z rel %.1e" % (span, st["n_boxes"], st["fft_side"], rel(b, a), abs(Zb - Za) / Za))
    # ---- D: column-sorted SpMV (needs a re-ordered context and a graph with real edges)
    import bench_util
    nD = 200000
    rowD, colD, valD, labD = bench_util.knn_like_graph(nD, 15, seed=0)
    YD = bench_util.clustered_embedding(labD, 2, 120.0)
    OUT["D"] = {}
    res = {}
    for name, fl in (("csr", 0), ("sorted", fb.FLAG_SORTED_SPMV)):
        with fb.FitSNE(rowD, colD, valD, YD, flags=fl) as t:          # re-ordering ON (first gradient re-orders)
            dC, Z = t.gradient(1.0)
        res[name] = dC
        with fb.FitSNE(rowD, colD, valD, YD, flags=fl | fb.FLAG_TIMERS) as t:
            for _ in range(3):
                t.step(exaggeration=1.0, momentum=0.8, learning_rate=1000.0, max_step_norm=5.0)
            t.reset_stats()
            for _ in range(20):
                t.step(exaggeration=1.0, momentum=0.8, learning_rate=1000.0, max_step_norm=5.0)
            OUT["D"][name + "_attract_update_ms"] = t.stats()["phase_ms"]["attract_update"] / 20
    OUT["D"]["sorted_vs_csr"] = rel(res["sorted"], res["csr"])
    say("200k kNN-like graph: sorted vs CSR gradient rel %.2e; attract+update ms: csr %.4f sorted %.4f" % (
        OUT["D"]["sorted_vs_csr"], OUT["D"]["csr_attract_update_ms"], OUT["D"]["sorted_attract_update_ms"]))
    say("ONESHOT_DONE")
except Exception as e:                                       # keep whatever was measured
    say("FAILED: %r" % (e,))
    raise
