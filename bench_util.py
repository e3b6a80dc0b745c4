"""Synthetic workloads of BASELINE.json's configs (shared by bench.py, tests and the report scripts).

config 3: N points in 10 Gaussian clusters, fixed kNN-style graph (K random same-cluster neighbours per row,
symmetrised (A+A^T)/2, normalised to sum 1) -> E ~ 2*K*N, injected exactly like the reference's
load_affinities=1 files.  Everything is seeded with numpy default_rng.
"""
import numpy as np


def knn_like_graph(N, K, seed=0, n_clusters=10):
    """CSR (row u32, col u32, val f64-of-f32) of a symmetric same-cluster random-neighbour graph, sum(val)=1."""
    rng = np.random.default_rng(seed)
    labels = rng.integers(0, n_clusters, N).astype(np.int32)
    order = np.argsort(labels, kind="stable")
    starts = np.searchsorted(labels[order], np.arange(n_clusters + 1))
    # neighbour j of row i: random member of i's cluster
    pos_in_cluster = rng.random((N, K))
    sizes = (starts[1:] - starts[:-1])[labels]
    nb = order[starts[labels][:, None] + np.minimum((pos_in_cluster * sizes[:, None]).astype(np.int64), sizes[:, None] - 1)]
    rows = np.repeat(np.arange(N, dtype=np.int64), K)
    cols = nb.reshape(-1).astype(np.int64)
    w = (rng.random(N * K) + 0.1)
    keep = rows != cols
    rows, cols, w = rows[keep], cols[keep], w[keep]
    # symmetrise: concatenate both directions, sort by (row, col), merge duplicates
    r2 = np.concatenate([rows, cols])
    c2 = np.concatenate([cols, rows])
    w2 = np.concatenate([w, w])
    key = r2 * N + c2
    o = np.argsort(key, kind="stable")
    key, w2 = key[o], w2[o]
    first = np.ones(len(key), bool)
    first[1:] = key[1:] != key[:-1]
    idx = np.nonzero(first)[0]
    wsum = np.add.reduceat(w2, idx)
    ukey = key[idx]
    rr = (ukey // N).astype(np.int64)
    cc = (ukey % N).astype(np.uint32)
    val = (wsum / wsum.sum()).astype(np.float32).astype(np.float64)
    row = np.zeros(N + 1, np.uint32)
    row[1:] = np.cumsum(np.bincount(rr, minlength=N)).astype(np.uint32)
    return row, cc, val, labels


def clustered_embedding(labels, dims, span, seed=1, n_clusters=10, spread=0.03):
    """A late-phase looking embedding: cluster centres in a box of side `span`, Gaussian blobs around them."""
    rng = np.random.default_rng(seed)
    centres = rng.uniform(-0.5, 0.5, (n_clusters, dims)) * span
    Y = centres[labels] + rng.standard_normal((len(labels), dims)) * span * spread
    Y -= Y.mean(0)
    return Y.astype(np.float32).astype(np.float64)


def early_embedding(N, dims, seed=2):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((N, dims)) * 1e-4).astype(np.float32).astype(np.float64)


def write_reference_inputs(dirname, row, col, val, Y0, max_iter, no_dims, learning_rate, stop_lying_iter,
                           mom_switch_iter, early_exag=12.0, df=1.0, nterms=3, ipi=1.0, min_int=50,
                           max_step_norm=5.0, start_late_exag_iter=-1, late_exag_coeff=-1.0, momentum=0.5,
                           final_momentum=0.8, no_momentum_during_exag=0):
    """data.dat + P_row/P_col/P_val.dat for `fast_tsne <ver> data.dat result.dat <nthreads>` with
    load_affinities=1 (byte layout: /root/reference/fast_tsne.py:259-297 == src/tsne.cpp:1915-1985)."""
    import os
    import struct
    os.makedirs(dirname, exist_ok=True)
    N = len(Y0)
    np.ascontiguousarray(row, np.uint32).tofile(os.path.join(dirname, "P_row.dat"))
    np.ascontiguousarray(col, np.uint32).tofile(os.path.join(dirname, "P_col.dat"))
    np.ascontiguousarray(val, np.float64).tofile(os.path.join(dirname, "P_val.dat"))
    with open(os.path.join(dirname, "data.dat"), "wb") as f:
        D = 1
        f.write(struct.pack("=i", N)); f.write(struct.pack("=i", D))
        f.write(struct.pack("=d", 0.5))            # theta
        f.write(struct.pack("=d", -1.0))           # perplexity < 0: manual sigma/K branch (never reached with load)
        f.write(struct.pack("=i", no_dims)); f.write(struct.pack("=i", max_iter))
        f.write(struct.pack("=i", stop_lying_iter)); f.write(struct.pack("=i", mom_switch_iter))
        f.write(struct.pack("=d", momentum)); f.write(struct.pack("=d", final_momentum))
        f.write(struct.pack("=d", learning_rate)); f.write(struct.pack("=d", max_step_norm))
        f.write(struct.pack("=i", 1)); f.write(struct.pack("=d", 1.0))     # K, sigma
        f.write(struct.pack("=i", 2)); f.write(struct.pack("=i", 1))       # nbody_algo=FFT, knn_algo
        f.write(struct.pack("=d", early_exag)); f.write(struct.pack("=i", no_momentum_during_exag))
        f.write(struct.pack("=i", 1)); f.write(struct.pack("=i", 1))       # n_trees, search_k
        f.write(struct.pack("=i", start_late_exag_iter)); f.write(struct.pack("=d", late_exag_coeff))
        f.write(struct.pack("=i", nterms)); f.write(struct.pack("=d", ipi)); f.write(struct.pack("=i", min_int))
        f.write(np.zeros(N * D, np.float64).tobytes())
        f.write(struct.pack("=i", 42)); f.write(struct.pack("=d", df)); f.write(struct.pack("=i", 1))
        f.write(np.ascontiguousarray(Y0, np.float64).tobytes())


def read_result(path):
    import struct
    with open(path, "rb") as f:
        n, d = struct.unpack("=ii", f.read(8))
        Y = np.frombuffer(f.read(8 * n * d), np.float64).reshape(n, d).copy()
        (m,) = struct.unpack("=i", f.read(4))
        costs = np.frombuffer(f.read(8 * m), np.float64).copy()
    return Y, costs
