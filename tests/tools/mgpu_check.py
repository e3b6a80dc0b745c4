"""Run under torchrun (one rank per GPU): sharded context vs the ORACLE (gradient rel-L2 <= 1e-4, sum_Q <= 1e-5: the
north-star tolerances) and vs a single-GPU context on identical inputs (gradient, KL, a 60-step run).
Prints 'MGPU_OK' from rank 0 when every check passes."""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "fit-sne_b200"))
import bench_util  # noqa: E402
import fitsne_b200 as fb  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    def fresh_id():
        """every sharded context is its own NCCL communicator and needs its own ncclUniqueId"""
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf = (ctypes.c_char * 128)()
            assert fb.load_library().fitsne_nccl_unique_id(buf) == 0
            idt.copy_(torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        return idt.cpu().numpy().tobytes()

    def mark(msg):
        if os.environ.get("MGPU_TRACE"):
            print("[rank %d] %s" % (rank, msg), file=sys.stderr, flush=True)

    N = int(os.environ.get("MGPU_N", "200003"))   # not divisible by the world size on purpose
    row, col, val, labels = bench_util.knn_like_graph(N, 8, seed=3)
    ok = True
    cases = ((2, 1.0, 60.0), (1, 0.5, 120.0))
    if os.environ.get("MGPU_CASES"):
        cases = tuple(cases[int(i)] for i in os.environ["MGPU_CASES"].split(","))
    for dims, df, span in cases:
        Y0 = bench_util.clustered_embedding(labels, dims, span, seed=5)
        mark("case dims=%d: create" % dims)
        t = fb.FitSNE(row, col, val, Y0, df=df, device=local, rank=rank, world=world, nccl_id=fresh_id())
        mark("gradient")
        dC, Z = t.gradient(4.0)
        mark("kl")
        kl = t.kl(4.0)
        b, e = t.row_begin, t.row_end
        sched = dict(max_iter=60, stop_lying_iter=20, mom_switch_iter=20, learning_rate=500.0, early_exag_coeff=4.0)
        mark("run")
        Y, costs = t.run(**sched)
        mark("close")
        t.close()
        mark("closed; compare: |Y| %.12g costs %s" % (np.linalg.norm(Y), costs[costs != 0]))
        if os.environ.get("MGPU_SKIP_SINGLE"):
            continue
        # every rank must hold the same Y after the all-gather
        ty = torch.from_numpy(Y).cuda()
        ref = ty.clone()
        dist.broadcast(ref, 0)
        same = bool(torch.equal(ty, ref))
        # gather the sharded gradient rows on rank 0
        parts = [None] * world
        dist.all_gather_object(parts, (b, e, dC[b:e]))
        if rank == 0:
            full = np.zeros_like(dC)
            for bb, ee, part in parts:
                full[bb:ee] = part
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            from pyoracle import Oracle
            ref_o, Z_o = Oracle().gradient(Y0, row, col, 4.0 * val, df=df)
            r_o = np.linalg.norm(full - ref_o) / np.linalg.norm(ref_o)
            print("dims=%d df=%g: sharded gradient vs ORACLE rel-L2 %.2e, sum_Q rel %.2e" % (dims, df, r_o, abs(Z - Z_o) / Z_o), flush=True)
            ok = ok and r_o < 1e-4 and abs(Z - Z_o) / Z_o < 1e-5
            mark("single-GPU reference")
            with fb.FitSNE(row, col, val, Y0, df=df, device=local) as s:
                dC1, Z1 = s.gradient(4.0)
                kl1 = s.kl(4.0)
                Y1, costs1 = s.run(**sched)
            r_g = np.linalg.norm(full - dC1) / np.linalg.norm(dC1)
            r_y = np.linalg.norm(Y - Y1) / np.linalg.norm(Y1)
            nz = costs1 != 0
            r_c = np.max(np.abs(costs[nz] - costs1[nz]) / np.abs(costs1[nz]))
            mark("costs sharded %s single %s" % (costs[costs != 0], costs1[nz]))
            print("dims=%d df=%g: grad rel-L2 %.2e, Z rel %.2e, KL rel %.2e, 60-step Y rel-L2 %.2e, costs rel %.2e, ranks agree %s"
                  % (dims, df, r_g, abs(Z - Z1) / Z1, abs(kl - kl1) / abs(kl1), r_y, r_c, same), flush=True)
            ok = ok and r_g < 1e-5 and abs(Z - Z1) / Z1 < 1e-6 and abs(kl - kl1) / abs(kl1) < 1e-6 and r_c < 1e-2 and r_y < 5e-2
        flag = torch.tensor([1 if same else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = ok and bool(flag.item())
    if rank == 0 and ok:
        print("MGPU_OK", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
