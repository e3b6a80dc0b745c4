#!/usr/bin/env python
"""bench.py -- t-SNE gradient-loop iterations/sec (BASELINE.json's metric) on 1..8 B200s, with roofline, parity and CPU baseline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--points 1000000] [--phase late|early]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full iteration of TSNE::run's loop (gradient + gains/momentum update + zero-mean; the KL
evaluation runs every 50th step inside the timed region, as in the reference).  The headline (`value`) is
BASELINE.json's config 3: synthetic N=1M points, 2-D, learning_rate=N/12, fixed kNN-style graph injected like
load_affinities=1 (E ~ 30 N), late phase.  `value` times device-resident state with CUDA events on the library's
stream (fitsne_run); `e2e` times create + run + download through the C ABI with host buffers and a host clock
(single GPU: the median of three whole calls after one untimed call; all three are listed in `e2e.calls_ms`).
The same line also carries: `roofline` (the phase that takes longest, live CUDA-event times), `roofline_total`,
`kernels` (every phase), `parity` (our gradient vs the oracle on this very workload, outside the timed region),
`cpu_baseline` (the unmodified reference on the host cores) and `other_configs` (BASELINE configs 1, 2, 4, 5 and
the early phase of config 3, short runs).  `--impl reference` times the UNMODIFIED reference binary
(oracle/_ref/fast_tsne_ref) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "fit-sne_b200"))
os.environ.setdefault("FITSNE_BENCH_CACHE", os.path.join(tempfile.gettempdir(), "fitsne_bench_cache"))
import bench_util  # noqa: E402

METRIC = "tsne_iterations_per_sec"
UNIT = "it/s"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def workload(points, phase, K_nn=15, dims=2, span=170.0):
    """config 3 / 4: (row, col, val, Y0, schedule) -- the kNN-like graph + an early or late looking embedding."""
    row, col, val, labels = bench_util.knn_like_graph(points, K_nn, seed=0)
    return (row, col, val) + embedding_and_schedule(points, labels, phase, dims, span)


def embedding_and_schedule(points, labels, phase, dims, span, late_exag=None):
    if phase == "early":
        Y0 = bench_util.early_embedding(points, dims)
        sched = dict(early_exag_coeff=12.0, stop_lying_iter=10 ** 9, mom_switch_iter=10 ** 9, momentum=0.5, final_momentum=0.8)
    else:
        Y0 = bench_util.clustered_embedding(labels, dims, span)
        # late phase: exaggeration off (coefficient 1 from the start) or the late-exaggeration coefficient, final momentum
        sched = dict(early_exag_coeff=late_exag or 1.0, stop_lying_iter=10 ** 9 if late_exag else -1, mom_switch_iter=-1, momentum=0.8,
                     final_momentum=0.8)
    sched.update(learning_rate=points / 12.0, max_step_norm=5.0, start_late_exag_iter=-1, late_exag_coeff=-1.0)
    return Y0, sched


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smmax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smmax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smmax) if smmax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------- reference arm --
def run_reference(points, phase, steps, threads, keep_dir=None, dims=2, df=1.0):
    """Time the unmodified reference binary's own loop (its 'N iterations in X seconds' lines, tsne.cpp:574)."""
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "fast_tsne_ref")
    if not os.path.exists(ref_bin):
        return None, "oracle/_ref/fast_tsne_ref missing (built by __graft_entry__.build() where /root/reference exists)"
    row, col, val, Y0, sched = workload(points, phase, dims=dims)
    with tempfile.TemporaryDirectory(dir=keep_dir) as td:
        bench_util.write_reference_inputs(td, row, col, val, Y0, max_iter=steps, no_dims=dims, df=df, learning_rate=sched["learning_rate"],
                                          stop_lying_iter=sched["stop_lying_iter"] if sched["stop_lying_iter"] < 10 ** 8 else steps + 1,
                                          mom_switch_iter=sched["mom_switch_iter"] if sched["mom_switch_iter"] < 10 ** 8 else steps + 1,
                                          early_exag=sched["early_exag_coeff"], momentum=sched["momentum"],
                                          final_momentum=sched["final_momentum"], max_step_norm=sched["max_step_norm"])
        env = dict(os.environ, MKL_NUM_THREADS="1", OMP_NUM_THREADS="1")   # reference FFTs are single-threaded
        t0 = time.perf_counter()
        out = subprocess.run([ref_bin, "1.2.1", "data.dat", "result.dat", str(threads)], cwd=td, env=env,
                             capture_output=True, text=True)
        wall = time.perf_counter() - t0
        if out.returncode != 0:
            return None, "reference binary failed rc=%d: %s" % (out.returncode, out.stdout[-300:] + out.stderr[-300:])
        secs = 0.0
        for ln in out.stdout.splitlines():
            if ln.startswith("Iteration ") and "iterations in" in ln:
                secs += float(ln.split("iterations in")[1].split("seconds")[0])
        if secs <= 0:
            return None, "could not parse the reference's timing lines"
    return {"loop_seconds": secs, "wall_seconds": wall, "iterations": steps, "it_per_s": steps / secs, "n_edges": int(len(col))}, None


# ------------------------------------------------------------------------------------------ our arm --
class Launcher:
    """torch.distributed plumbing: rank / world, barrier, max over ranks, ncclUniqueId distribution."""

    def __init__(self):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
        torch.cuda.set_device(self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist

    def barrier(self):
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max(self, x):
        if not self.dist:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_array(self, a):
        if not self.dist:
            return a
        t = self.torch.from_numpy(np.ascontiguousarray(a)).cuda()
        self.dist.all_reduce(t)
        return t.cpu().numpy()

    def nccl_id(self, fb):
        if self.world == 1:
            return None
        import ctypes
        idt = self.torch.zeros(128, dtype=self.torch.uint8, device="cuda")
        if self.rank == 0:
            buf = (ctypes.c_char * 128)()
            rc = fb.load_library().fitsne_nccl_unique_id(buf)
            assert rc == 0, "fitsne_nccl_unique_id failed"
            idt.copy_(self.torch.frombuffer(bytearray(buf.raw), dtype=self.torch.uint8))
        self.dist.broadcast(idt, 0)
        return bytes(idt.cpu().numpy().tobytes())

    def graph(self, maker):
        """rank 0 builds (and caches) the graph first, then everybody loads it: one generation per box, not one per rank"""
        if self.dist:
            if self.rank == 0:
                out = maker()
            self.dist.barrier()
            if self.rank != 0:
                out = maker()
            return out
        return maker()


def timed_run(L, fb, row, col, val, Y0, sched, steps, warmup, dims=2, df=1.0, keep=False):
    """W untimed + K timed steps on device-resident state; returns (it/s, ms total, stats, costs[, context])."""
    t = fb.FitSNE(row, col, val, Y0, df=df, device=L.local_rank, rank=L.rank, world=L.world, nccl_id=L.nccl_id(fb))
    t.run(fetch_Y=False, max_iter=max(warmup, 3), **sched)          # graph capture, clocks, first re-ordering
    b0 = t.stats()["n_boxes"]
    t.prewarm(max(25, b0 - 60), b0 + 90)     # twiddle tables for the FFT lengths the run can drift through (milliseconds)
    L.barrier()
    t.reset_stats()
    L.barrier()
    _, costs = t.run(fetch_Y=False, max_iter=steps, **sched)   # CUDA events around the loop, on the library's stream
    L.barrier()
    ms = L.max(t.last_run_ms())
    st = t.stats()
    if keep:
        return steps / (ms * 1e-3), ms, st, costs, t
    t.close()
    return steps / (ms * 1e-3), ms, st, costs, None


def phase_times(L, fb, row, col, val, Y, uY, gains, sched, nsteps, dims, df):
    """Per-phase device times, live, with CUDA events on the launching stream (timers mode = plain launches, serialised)."""
    tt = fb.FitSNE(row, col, val, Y, df=df, device=L.local_rank, flags=fb.FLAG_TIMERS, rank=L.rank, world=L.world, nccl_id=L.nccl_id(fb))
    tt.set_optimizer_state(uY, gains)
    kw = dict(exaggeration=sched["early_exag_coeff"], momentum=sched["momentum"], learning_rate=sched["learning_rate"], max_step_norm=5.0)
    for _ in range(3):
        tt.step(**kw)
    b0 = tt.stats()["n_boxes"]
    tt.prewarm(max(25, b0 - 30), b0 + 30)
    tt.reset_stats()
    for _ in range(nsteps):
        tt.step(**kw)
    sst = tt.stats()
    if os.environ.get("FITSNE_KTIMES"):      # diagnostics: every rank's own warm per-kernel times (us per launch)
        kt = tt.kernel_times()
        print("[ktimes rank %d] " % L.rank + "  ".join("%s %.1f" % (k, 1e3 * v[0] / max(v[1], 1)) for k, v in kt.items()), file=sys.stderr, flush=True)
    tt.close()
    return sst


def short_config(L, fb, name, desc, maker, phase, dims, df, span, steps, late_exag=None):
    """One of the other BASELINE configs: a short device-resident run, same timing rules."""
    t0 = time.perf_counter()
    try:
        row, col, val, labels = L.graph(maker)
        N = len(row) - 1
        Y0, sched = embedding_and_schedule(N, labels, phase, dims, span, late_exag)
        gen_s = time.perf_counter() - t0
        value, ms, st, costs, _ = timed_run(L, fb, row, col, val, Y0, sched, steps, 5, dims=dims, df=df)
        kls = [float(c) for c in costs if c != 0]
        return {"workload": desc, "points": N, "n_edges": int(len(col)), "phase": phase, "dims": dims, "df": df, "steps": steps,
                "value": round(value, 1), "unit": UNIT, "ms_per_step": round(ms / steps, 5),
                "grid": {"n_boxes": st["n_boxes"], "grid_side": st["grid_side"], "fft_side": st["fft_side"]},
                "kl_last": kls[-1] if kls else None, "gpu_launches": int(st["kernel_launches"]), "setup_seconds": round(gen_s, 1)}
    except Exception as e:            # a failing extra must not take the headline down with it
        return {"workload": desc, "error": repr(e)[:300]}


def from_raw_data(fb, device, N=70000, D=50, perplexity=30.0, iters=750):
    """The whole job a wrapper call does, from the data matrix: exact kNN (K = 3 * perplexity), perplexity calibration and
    symmetrisation on the device, then the reference's default schedule (750 iterations, early exaggeration 12 for 250) -- the
    MNIST-sized case of BASELINE config 2 on a synthetic 10-component Gaussian mixture in 50 dimensions."""
    desc = "config 2 end to end from X: N=%d, D=%d, perplexity %g -> K=%d, %d iterations (preprocessing included)" % (N, D, perplexity, int(3 * perplexity), iters)
    try:
        rng = np.random.default_rng(0)
        centres = rng.standard_normal((10, D)) * 4.0
        X = centres[rng.integers(0, 10, N)] + rng.standard_normal((N, D))
        Y0 = rng.standard_normal((N, 2)) * 1e-4
        t0 = time.perf_counter()
        Xc = X - X.mean(0)
        Xc /= np.abs(Xc).max()                                   # the reference's prologue (tsne.cpp:153-161)
        nbr, dist = fb.knn(Xc, int(3 * perplexity), device)
        t1 = time.perf_counter()
        row, col, val = fb.similarities(nbr, dist, perplexity, device=device)
        t2 = time.perf_counter()
        Y, costs = fb.run_host(row, col, val, Y0, max_iter=iters, stop_lying_iter=250, mom_switch_iter=250, momentum=0.5, final_momentum=0.8,
                               learning_rate=max(200.0, N / 12.0), early_exag_coeff=12.0, device=device)
        t3 = time.perf_counter()
        kls = [float(c) for c in costs if c != 0]
        return {"workload": desc, "points": N, "n_edges": int(len(col)), "seconds": {"knn": round(t1 - t0, 3), "similarities": round(t2 - t1, 3),
                "iterations (create + run + download)": round(t3 - t2, 3), "total": round(t3 - t0, 3)},
                "value": round(iters / (t3 - t0), 1), "unit": UNIT + " (iterations / total seconds, preprocessing included)",
                "kl_last": kls[-1] if kls else None, "finite": bool(np.isfinite(Y).all())}
    except Exception as e:            # a failing extra must not take the headline down with it
        return {"workload": desc, "error": repr(e)[:300]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--points", type=int, default=1000000)
    ap.add_argument("--phase", default="late", choices=["late", "early"])
    ap.add_argument("--dims", type=int, default=2, choices=[1, 2], help="embedding dimension (BASELINE config 5 uses 1)")
    ap.add_argument("--df", type=float, default=1.0, help="t-kernel degrees of freedom (config 5 uses 0.5)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the other BASELINE configs (1, 2, 4, 5, early phase)")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()

    # watchdog: a wedged collective must not hold the box -- give up loudly after 20 minutes (the default run takes ~3-4).
    def _give_up():
        sys.stderr.write("bench.py: watchdog expired after 1200 s (rank %s); aborting\n" % os.environ.get("RANK", "0"))
        sys.stderr.flush()
        os._exit(3)
    wd = threading.Timer(1200.0, _give_up)
    wd.daemon = True
    wd.start()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    threads = os.cpu_count() or 1
    config = {"workload": "BASELINE config 3: synthetic N=%d, 2-D, lr=N/12, fixed kNN-style graph (K=15 same-cluster neighbours, "
                          "symmetrised, ~30 nnz/row) injected as load_affinities=1; %s phase" % (args.points, args.phase),
              "points": args.points, "phase": args.phase, "dims": args.dims, "df": args.df, "nterms": 3, "intervals_per_integer": 1, "min_num_intervals": 50,
              "l2": "inputs larger than L2: the CSR P (8 B/edge, ~240 MB at 1M points) is streamed from HBM every step",
              "sharding": ("single GPU" if world <= 1 else "points/rows sharded across %d ranks; exchanges (partial grids, convolution transposes from 4 ranks up, "
                           "statistics, Y slices) go over peer memory on NVLink inside the kernels / by the copy engines, NCCL only at set-up" % world)}

    if args.impl == "reference":
        if rank != 0:
            return 0
        steps = min(args.steps, 60 if args.points >= 500000 else 1000)
        res, err = run_reference(args.points, args.phase, steps, threads, dims=args.dims, df=args.df)
        if res is None:
            print(json.dumps({"impl": "reference", "unavailable": err}))
            return 0
        v = res["it_per_s"]
        sample = "%d iterations of the full workload by oracle/_ref/fast_tsne_ref (unmodified reference + MKL-DFTI FFTW shim), %d threads, its own loop timer" % (steps, threads)
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                          "steps_timed": steps, "warmup": args.warmup, "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
                          "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    import fitsne_b200 as fb
    fb.load_library()
    L = Launcher()

    row, col, val, Y0, sched = L.graph(lambda: workload(args.points, args.phase, dims=args.dims))
    N, E, d = args.points, int(len(col)), args.dims

    # ---- parity, outside the timed region: our gradient vs the oracle on this very workload (rank 0 runs the oracle)
    parity = None
    if not args.no_parity:
        with fb.FitSNE(row, col, val, Y0, df=args.df, device=L.local_rank, rank=L.rank, world=L.world, nccl_id=L.nccl_id(fb)) as tp:
            dC, Z = tp.gradient(sched["early_exag_coeff"])
            kl = tp.kl(sched["early_exag_coeff"])
        if world > 1:                          # sharded: every rank computed its own rows; the rest of the buffer is scratch
            b_, e_ = fb.shard_range(N, L.rank, L.world)
            dC[:b_] = 0; dC[e_:] = 0
        dC = L.sum_array(dC)
        if rank == 0:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            try:
                from pyoracle import Oracle
                O = Oracle()
                a = sched["early_exag_coeff"]
                ref, Zr = O.gradient(Y0, row, col, a * val, df=args.df)
                klr = O.kl(Y0, row, col, a * val, Zr, df=args.df)
                parity = {"checker": "oracle/fitsne_oracle.c (fp64 restatement, pinned to the compiled reference's golden vectors)",
                          "gradient_rel_l2": float(np.linalg.norm(dC - ref) / np.linalg.norm(ref)), "tolerance": 1e-4,
                          "sum_Q_rel": abs(Z - Zr) / Zr, "kl_rel": abs(kl - klr) / abs(klr), "kl": kl,
                          "ok": bool(np.linalg.norm(dC - ref) / np.linalg.norm(ref) < 1e-4 and abs(Z - Zr) / Zr < 1e-5)}
            except Exception as e:
                parity = {"error": repr(e)[:200]}

    # ---- headline: W warm-up + K timed steps, device-resident
    L.barrier()
    sampler = ClockSampler(L.local_rank)
    if rank == 0:
        sampler.start()
    value, ms, st, costs, t = timed_run(L, fb, row, col, val, Y0, sched, args.steps, args.warmup, dims=d, df=args.df, keep=True)
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-phase durations + roofline (every rank runs the timers context: it is sharded like the run)
    Ynow = t.get_Y()
    uY, gains = t.get_optimizer_state()
    t.close()
    nsteps_t = 30
    sst = phase_times(L, fb, row, col, val, Ynow, uY, gains, sched, nsteps_t, d, args.df)
    kern, roofline, roofline_total = {}, None, None
    if rank == 0:
        peak, peak_src = load_peaks()
        G, M = sst["grid_side"], sst["fft_side"]
        Nl, El = N / world, E / world                      # per-rank points / edges (replicated: the convolution)
        # algorithmic bytes per launch group, per rank
        algo = {"attract_update": 8 * El + 30 * d * Nl, "spread": (4 + 4 * d) * Nl + 16 * G ** d, "gather": (12 + 4 * d) * Nl + 16 * G ** d,
                "sort": 16 * Nl + 2 * 16 * Nl, "fft": 4 * 28 * M ** d + 12 * M ** d, "center": 12 * d * Nl}
        names = {"attract_update": "k_attract + k_update (CSR SpMV over 8-byte edge words, then exaggeration/gains/momentum/clip/Y update + column sums)",
                 "fft": "convolution: k_kspec_rows/cols (kernel spectra) + k_conv_rows_fwd + k_conv_cols (TMA tiles, in-place column FFTs, Hadamard, sum_Q) + k_conv_rows_inv",
                 "sort": "k_bin + 2 x k_radix_sweep (box keys, stable LSD radix sort with look-back)",
                 "spread": "k_spread_chunks + k_spread_combine (Lagrange spread, stitched in shared memory)",
                 "gather": "k_gather", "center": "k_center_bounds (zero-mean + bounds)"}
        for k, b in algo.items():
            dur = sst["phase_ms"][k] / nsteps_t
            if k == "fft":
                dur += sst["phase_ms"]["kernel_spectrum"] / nsteps_t
            if dur > 0:
                kern[k] = {"ms": round(dur, 5), "algorithmic_bytes": int(b), "gbs": round(b / dur / 1e6, 1), "frac": round(b / dur / 1e6 / peak, 4)}
        if world > 1 and sst["phase_ms"].get("collectives", 0) > 0:
            kern["collectives"] = {"ms": round(sst["phase_ms"]["collectives"] / nsteps_t, 5), "what": "NCCL all-reduce of the spread grid (the Y all-gather overlaps the sort)"}
        dom = max((k for k in kern if k in algo), key=lambda k: kern[k]["ms"])
        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this very workload
        # (profiles/r2_ncu_full_iteration.txt: k_attract 252.07 + 9.35 MB, k_update 40.02 + 0.02 MB; the 126 MB L2 absorbs most
        # writes of the 8 MB per-point arrays).  Only quoted for the configuration the capture was taken on.
        traffic, traffic_note = None, "not measured for this configuration; ncu dram__bytes of the 1M-point single-GPU workload are in profiles/"
        if dom == "attract_update" and world == 1 and N == 1000000 and d == 2 and args.phase == "late" and args.df == 1.0:
            traffic = int(round((252.067584 + 9.353728 + 40.021760 + 0.019712) * 1e6))
            traffic_note = ("ncu --set full capture of the same workload, committed as profiles/r2_ncu_full_iteration.txt (k_attract + k_update, "
                            "DRAM read + write per launch); not re-measured in this run")
        roofline = {"kernel": names[dom], "phase": dom, "dominant_by": "time (live CUDA events, serialised phases)", "bound": "hbm",
                    "achieved": kern[dom]["gbs"], "peak": peak, "unit": "GB/s", "frac": kern[dom]["frac"], "traffic": traffic,
                    "traffic_note": traffic_note,
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": kern[dom]["algorithmic_bytes"], "avg_launch_ms": kern[dom]["ms"]}
        total_bytes = 8 * El + 140 * Nl + 32 * G ** d + 124 * M ** d          # whole iteration
        roofline_total = {"algorithmic_bytes_per_step": int(total_bytes), "ms_per_step": round(ms / args.steps, 5),
                          "achieved": round(total_bytes / (ms / args.steps) / 1e6, 1), "peak": peak, "unit": "GB/s",
                          "frac": round(total_bytes / (ms / args.steps) / 1e6 / peak, 4)}

    # ---- end to end through the C ABI with host buffers: upload P + Y0, run K iterations, download Y + costs
    e2e = None
    if not args.no_e2e:
        torch = L.torch
        b, e = fb.shard_range(N, L.rank, L.world)
        lc, lv = (col, val) if world == 1 else (np.ascontiguousarray(col[row[b]:row[e]]), np.ascontiguousarray(val[row[b]:row[e]]))
        pinned = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (row, lc, lv, Y0)]   # keep the owners alive
        prow, pcol, pval, pY = [p_.numpy() for p_ in pinned]
        nid = L.nccl_id(fb)
        calls_ms = []
        if world == 1:
            # the whole call, repeated: one untimed call, then three timed ones; the MEDIAN is reported (a call is ~0.1 s of
            # host-side work -- allocation, a 364 MB upload over PCIe -- and single calls on a shared box scatter by 2x)
            for rep in range(4):
                L.barrier()
                t0 = time.perf_counter()
                Yout, costs2 = fb.run_host(prow, pcol, pval, pY, max_iter=args.steps, device=L.local_rank, df=args.df, **sched)
                if rep > 0:
                    calls_ms.append((time.perf_counter() - t0) * 1e3)
            dt = sorted(calls_ms)[len(calls_ms) // 2] * 1e-3
        else:
            L.barrier()
            t0 = time.perf_counter()
            with fb.FitSNE(prow, pcol, pval, pY, df=args.df, device=L.local_rank, rank=L.rank, world=L.world, nccl_id=nid) as te:
                t_created = time.perf_counter()
                Yout, costs2 = te.run(max_iter=args.steps, **sched)
                t_ran = time.perf_counter()
            L.barrier()
            dt = L.max(time.perf_counter() - t0)
        h2d = prow.nbytes + pcol.nbytes + pval.nbytes + pY.nbytes
        d2h = Yout.nbytes + costs2.nbytes
        e2e = {"value": args.steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps,
               "call": ("fitsne_run_host" if world == 1 else "fitsne_create_sharded + fitsne_run + download, per rank") +
                       " (create + %d iterations + download), host wall clock, max over ranks" % args.steps}
        if calls_ms:
            e2e["calls_ms"] = [round(x, 1) for x in calls_ms]
            e2e["aggregate"] = "median of 3 timed calls after 1 untimed call"
        if world > 1:     # where a sharded call's time goes: the one-off set-up (NCCL communicator, peer-memory handles, CSR upload) vs the iterations
            e2e["breakdown_s"] = {"create_sharded (NCCL init + IPC fabric + upload)": round(L.max(t_created - t0), 3),
                                  "run + download": round(L.max(t_ran - t_created), 3)}
        del prow, pcol, pval, pY, pinned       # release the pinned buffers while the CUDA context is still alive

    # ---- the other BASELINE configs, short runs (every rank takes part: the contexts are sharded like the headline)
    extras = None
    if not args.no_extras:
        extras = {}
        if args.phase == "late":
            v2, ms2, st2, _, _ = timed_run(L, fb, row, col, val, *embedding_and_schedule(N, None, "early", d, 0), 100, 5, dims=d, df=args.df)
            extras["config3_early_phase"] = {"value": round(v2, 1), "unit": UNIT, "ms_per_step": round(ms2 / 100, 5), "steps": 100,
                                             "grid": {"n_boxes": st2["n_boxes"], "fft_side": st2["fft_side"]}}
        del row, col, val
        extras["config1_10k"] = short_config(L, fb, "c1", "config 1: N=10k, perplexity-30-like graph (~137 nnz/row), 2-D, late phase",
                                             lambda: bench_util.knn_like_graph(10000, 69, seed=0), "late", 2, 1.0, 60.0, 200)
        extras["config2_70k_late_exag"] = short_config(L, fb, "c2", "config 2: N=70k, ~137 nnz/row, 2-D, late exaggeration 4",
                                                       lambda: bench_util.knn_like_graph(70000, 69, seed=0), "late", 2, 1.0, 110.0, 200, late_exag=4.0)
        extras["config5_1d_df05_K300"] = short_config(L, fb, "c5", "config 5: N=1M, 1-D, df=0.5, 300 neighbours per point (perplexity list [10,100] -> K=300)",
                                                      lambda: bench_util.ring_cluster_graph(1000000, 300, seed=0), "late", 1, 0.5, 900.0, 50)
        extras["config4_10M"] = short_config(L, fb, "c4", "config 4: N=10M, same generator as config 3 (K=15, ~30 nnz/row), 2-D, late phase",
                                             lambda: bench_util.knn_like_graph(10000000, 15, seed=0), "late", 2, 1.0, 170.0, 50)

        if world == 1:
            extras["config2_70k_from_raw_data"] = from_raw_data(fb, L.local_rank)

    cpu_baseline = None
    if rank == 0 and not args.no_cpu_baseline:
        steps_ref = 20 if args.points >= 500000 else 200
        res, err = run_reference(args.points, args.phase, steps_ref, threads, dims=args.dims, df=args.df)
        if res is not None:
            cpu_baseline = {"value": res["it_per_s"], "unit": UNIT, "cores": threads, "kind": "reference",
                            "sample": "%d iterations of the same workload by oracle/_ref/fast_tsne_ref (unmodified reference, FFTW->MKL shim), %d threads" % (steps_ref, threads)}
        else:
            cpu_baseline = {"value": None, "unit": UNIT, "cores": threads, "kind": "reference", "sample": "unavailable: " + err}

    if rank == 0:
        kls = [float(c) for c in costs if c != 0]
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config, "impl": "ours", "clocks": clocks, "e2e": e2e,
                "gpu_launches": int(st["kernel_launches"]), "graph_launches": int(st["graph_launches"]), "regrids": int(st["regrids"]),
                "grid": {"n_boxes": st["n_boxes"], "grid_side": st["grid_side"], "fft_side": st["fft_side"]},
                "n_edges": E, "kl_last": kls[-1] if kls else None, "parity": parity, "roofline": roofline, "roofline_total": roofline_total,
                "kernels": kern, "cpu_baseline": cpu_baseline, "other_configs": extras}
        print(json.dumps(line))
    if L.dist:
        L.dist.barrier()
        L.dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
