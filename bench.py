#!/usr/bin/env python
"""bench.py -- t-SNE gradient-loop iterations/sec (BASELINE.json's metric) on 1..8 B200s, with roofline and CPU baseline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--points 1000000] [--phase late|early]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full iteration of TSNE::run's loop (gradient + gains/momentum update + zero-mean; the KL
evaluation runs every 50th step inside the timed region, as in the reference) on BASELINE.json's config 3:
synthetic N=1M points, 2-D, learning_rate=N/12, fixed kNN-style graph injected like load_affinities=1
(E ~ 30 N).  `value` times device-resident state with CUDA events on the library's stream (fitsne_run);
`e2e` times fitsne_run_host (host CSR P + host Y in, host Y + costs out) with a host clock.
`--impl reference` times the UNMODIFIED reference binary (oracle/_ref/fast_tsne_ref) on the host cores.
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
import bench_util  # noqa: E402

METRIC = "tsne_iterations_per_sec"
UNIT = "it/s"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def workload(points, phase, K_nn=15, dims=2):
    row, col, val, labels = bench_util.knn_like_graph(points, K_nn, seed=0)
    if phase == "early":
        Y0 = bench_util.early_embedding(points, dims)
        sched = dict(early_exag_coeff=12.0, stop_lying_iter=10 ** 9, mom_switch_iter=10 ** 9, momentum=0.5, final_momentum=0.8)
    else:
        Y0 = bench_util.clustered_embedding(labels, dims, 170.0)
        # late phase: exaggeration off (coefficient 1 from the start), final momentum
        sched = dict(early_exag_coeff=1.0, stop_lying_iter=-1, mom_switch_iter=-1, momentum=0.8, final_momentum=0.8)
    sched.update(learning_rate=points / 12.0, max_step_norm=5.0, start_late_exag_iter=-1, late_exag_coeff=-1.0)
    return row, col, val, Y0, sched


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
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
    args = ap.parse_args()
    # watchdog: a wedged collective must not hold the box -- give up loudly after 15 minutes (the default run takes ~1-2).
    # A thread, not SIGALRM: the main thread may be blocked inside the library (ctypes releases the GIL, Python-level
    # signal handlers would only run once the call returns).
    def _give_up():
        sys.stderr.write("bench.py: watchdog expired after 900 s (rank %s); aborting\n" % os.environ.get("RANK", "0"))
        sys.stderr.flush()
        os._exit(3)
    wd = threading.Timer(900.0, _give_up)
    wd.daemon = True
    wd.start()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    threads = os.cpu_count() or 1
    config = {"workload": "BASELINE config 3: synthetic N=%d, 2-D, lr=N/12, fixed kNN-style graph (K=15 same-cluster neighbours, "
                          "symmetrised, ~30 nnz/row) injected as load_affinities=1; %s phase" % (args.points, args.phase),
              "points": args.points, "phase": args.phase, "dims": args.dims, "df": args.df, "nterms": 3, "intervals_per_integer": 1, "min_num_intervals": 50,
              "l2": "inputs larger than L2: the CSR P (~8 B/edge, ~240 MB at 1M points) is streamed from HBM every step",
              "sharding": "points/rows sharded across %d rank(s); NCCL grid all-reduce + Y all-gather" % max(world, 1)}

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

    import torch
    import fitsne_b200 as fb
    fb.load_library()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    nccl_id = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            import ctypes
            buf = (ctypes.c_char * 128)()
            rc = fb.load_library().fitsne_nccl_unique_id(buf)
            assert rc == 0, "fitsne_nccl_unique_id failed"
            idt.copy_(torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        nccl_id = bytes(idt.cpu().numpy().tobytes())

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    row, col, val, Y0, sched = workload(args.points, args.phase, dims=args.dims)
    N, E = args.points, int(len(col))
    t = fb.FitSNE(row, col, val, Y0, df=args.df, device=local_rank, rank=rank, world=world, nccl_id=nccl_id)
    # warm-up: W untimed steps (graph capture, clocks)
    t.run(fetch_Y=False, max_iter=max(args.warmup, 3), **sched)
    b0 = t.stats()["n_boxes"]
    t.prewarm(max(25, b0 - 60), b0 + 90)     # twiddle tables for the FFT lengths the run can drift through (milliseconds)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t.reset_stats()
    barrier()
    _, costs = t.run(fetch_Y=False, max_iter=args.steps, **sched)   # CUDA events around the loop, on the library's stream
    barrier()
    ms = t.last_run_ms()
    st = t.stats()
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        import torch.distributed as dist
        mt = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(mt, op=dist.ReduceOp.MAX)
        ms = float(mt.item())
    value = args.steps / (ms * 1e-3)

    # per-kernel durations, live, with CUDA events on the launching stream (timers mode = plain launches)
    kern = {}
    roofline = None
    if rank == 0:
        peak, peak_src = load_peaks()
        nsteps_t = 30
        tt = fb.FitSNE(row, col, val, t.get_Y(), df=args.df, device=local_rank, flags=fb.FLAG_TIMERS) if world == 1 else None
        if tt is not None:
            tt.prewarm(max(25, st["n_boxes"] - 30), st["n_boxes"] + 30)
        if tt is not None:
            uY, gains = t.get_optimizer_state()
            tt.set_optimizer_state(uY, gains)
            alpha = sched["early_exag_coeff"]
            for _ in range(3):
                tt.step(exaggeration=alpha, momentum=sched["momentum"], learning_rate=sched["learning_rate"], max_step_norm=5.0)
            tt.reset_stats()
            for _ in range(nsteps_t):
                tt.step(exaggeration=alpha, momentum=sched["momentum"], learning_rate=sched["learning_rate"], max_step_norm=5.0)
            sst = tt.stats()
            G, M = sst["grid_side"], sst["fft_side"]
            # algorithmic bytes per launch
            d = args.dims
            algo = {"attract_update": 8 * E + 30 * d * N, "spread": (4 + 4 * d) * N + 16 * G ** d, "gather": (12 + 4 * d) * N + 16 * G ** d,
                    "sort": 16 * N + 2 * 16 * N, "fft": 4 * 28 * M ** d + 12 * M ** d, "center": 12 * d * N}
            for k, b in algo.items():
                dur = sst["phase_ms"][k] / nsteps_t
                if k == "fft":
                    dur += sst["phase_ms"]["kernel_spectrum"] / nsteps_t
                if dur > 0:
                    kern[k] = {"ms": round(dur, 5), "algorithmic_bytes": int(b), "gbs": round(b / dur / 1e6, 1), "frac": round(b / dur / 1e6 / peak, 4)}
            tt.close()
            dom = "attract_update"
            # traffic: dram__bytes_read+write of k_attract from the committed ncu --set full capture of this exact workload
            # (profiles/r1_ncu_full_top_kernels.txt: 252.4 MB + 9.9 MB per launch); null for any other workload
            traffic = 262.3e6 if (N == 1000000 and d == 2 and args.phase == "late") else None
            roofline = {"kernel": "k_attract + k_update (CSR SpMV on its own stream, then exaggeration/gains/momentum/clip/Y update)", "bound": "hbm",
                        "achieved": kern[dom]["gbs"], "peak": peak, "unit": "GB/s", "frac": kern[dom]["frac"],
                        "traffic": traffic, "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": kern[dom]["algorithmic_bytes"], "avg_launch_ms": kern[dom]["ms"]}

    # end to end through the C ABI with host buffers: upload P + Y0, run K iterations, download Y + costs
    e2e = None
    if rank == 0 and world == 1 and not args.no_e2e:
        pinned = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (row, col, val, Y0)]   # keep the owners alive
        prow, pcol, pval, pY = [p_.numpy() for p_ in pinned]
        t.close()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        Yout, costs2 = fb.run_host(prow, pcol, pval, pY, max_iter=args.steps, device=local_rank, df=args.df, **sched)
        dt = time.perf_counter() - t0
        h2d = prow.nbytes + pcol.nbytes + pval.nbytes + pY.nbytes
        d2h = Yout.nbytes + costs2.nbytes
        e2e = {"value": args.steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps,
               "call": "fitsne_run_host (create + %d iterations + download), host wall clock" % args.steps}
        del prow, pcol, pval, pY, pinned       # release the pinned buffers while the CUDA context is still alive

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
                "n_edges": E, "kl_last": kls[-1] if kls else None, "roofline": roofline, "kernels": kern, "cpu_baseline": cpu_baseline}
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        t.close()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
