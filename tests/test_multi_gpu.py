"""Sharded (multi-GPU) path: NCCL grid all-reduce + Y all-gather.  The GPU test needs >= 2 devices (gpurun --gpus 2);
the gloo test exercises the launcher-side plumbing (shard ranges, CSR slicing, id broadcast) on the CPU."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.skipif(_n_gpus() < 2, reason="needs at least 2 GPUs")
def test_sharded_matches_single_gpu():
    n = min(_n_gpus(), 4)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
                          "--master-addr", "127.0.0.1", "--master-port", "29517",
                          os.path.join(ROOT, "tests", "tools", "mgpu_check.py")], capture_output=True, text=True, timeout=300)
    assert "MGPU_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "fit-sne_b200"))
    import bench_util
    import fitsne_b200 as fb
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the 128-byte id travels as a uint8 tensor from rank 0 (what bench.py does over NCCL)
    idt = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        idt.copy_(torch.arange(128, dtype=torch.uint8))
    dist.broadcast(idt, 0)
    N = 1001
    row, col, val, _ = bench_util.knn_like_graph(N, 4, seed=1)
    b, e = fb.shard_range(N, rank, world)
    # what FitSNE.__init__ hands to fitsne_create_sharded for this rank
    col_l, val_l = col[row[b]:row[e]], val[row[b]:row[e]]
    edges = torch.tensor([len(col_l)], dtype=torch.int64)
    dist.all_reduce(edges)
    rows = torch.tensor([e - b], dtype=torch.int64)
    dist.all_reduce(rows)
    wsum = torch.tensor([float(val_l.sum())], dtype=torch.float64)
    dist.all_reduce(wsum)
    q.put((rank, bytes(idt.numpy().tobytes()), b, e, int(edges.item()), int(rows.item()), float(wsum.item()), len(col)))
    dist.destroy_process_group()


def test_shard_plumbing_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29531
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
    (r0, id0, b0, e0, edges, rows, wsum, E), (r1, id1, b1, e1, *_) = res
    assert id0 == id1 == bytes(range(128))
    assert (b0, e0, b1, e1) == (0, 501, 501, 1001)          # contiguous ceil(N/world) slices (fitsne_create_sharded)
    assert rows == 1001 and edges == E and abs(wsum - 1.0) < 1e-9


def test_shard_range_covers_everything():
    import fitsne_b200 as fb
    for N in (7, 1000, 1000003):
        for world in (1, 2, 3, 4, 8):
            spans = [fb.shard_range(N, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == N
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            per = spans[0][1] - spans[0][0]
            assert all(e - b <= per for b, e in spans)
