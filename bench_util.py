"""Synthetic workloads of BASELINE.json's configs (shared by bench.py, tests and the report scripts).

config 3: N points in 10 Gaussian clusters, fixed kNN-style graph (K random same-cluster neighbours per row,
symmetrised (A+A^T)/2, normalised to sum 1) -> E ~ 2*K*N, injected exactly like the reference's
load_affinities=1 files.  Everything is seeded with numpy default_rng.
"""
import numpy as np


def knn_like_graph(N, K, seed=0, n_clusters=10):
    """CSR (row u32, col u32, val f64-of-f32) of a symmetric same-cluster random-neighbour graph, sum(val)=1.
    With FITSNE_BENCH_CACHE=<dir> the (deterministic) result is kept as an .npz there: torchrun ranks and repeated
    bench runs on one box then generate a 10M-point graph once instead of once per process."""
    import os
    cache = os.environ.get("FITSNE_BENCH_CACHE")
    path = os.path.join(cache, "knn_like_%d_%d_%d_%d.npz" % (N, K, seed, n_clusters)) if cache else None
    if path and os.path.exists(path):
        try:
            z = np.load(path)
            return z["row"], z["col"], z["val"], z["labels"]
        except Exception:
            pass
    out = _knn_like_graph(N, K, seed, n_clusters)
    if path:
        try:
            os.makedirs(cache, exist_ok=True)
            tmp = "%s.%d.tmp.npz" % (path, os.getpid())
            np.savez(tmp, row=out[0], col=out[1], val=out[2], labels=out[3])
            os.replace(tmp, path)
        except Exception:
            pass
    return out


def _knn_like_graph(N, K, seed=0, n_clusters=10):
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    labels = rng.integers(0, n_clusters, N).astype(np.int32)
    order = np.argsort(labels, kind="stable")
    starts = np.searchsorted(labels[order], np.arange(n_clusters + 1))
    # neighbour j of row i: random member of i's cluster
    pos_in_cluster = rng.random((N, K))
    sizes = (starts[1:] - starts[:-1])[labels]
    nb = order[starts[labels][:, None] + np.minimum((pos_in_cluster * sizes[:, None]).astype(np.int64), sizes[:, None] - 1)]
    rows = np.repeat(np.arange(N, dtype=np.int32), K)
    cols = nb.reshape(-1).astype(np.int32)
    w = (rng.random(N * K) + 0.1)
    keep = rows != cols
    # symmetrise: A + A^T with duplicates merged (scipy's O(nnz) counting-sort routines: 3x faster than sorting keys)
    A = sp.csr_matrix((w[keep], (rows[keep], cols[keep])), shape=(N, N))
    S = (A + A.T).tocsr()
    S.sort_indices()
    val = (S.data / S.data.sum()).astype(np.float32).astype(np.float64)
    return S.indptr.astype(np.uint32), S.indices.astype(np.uint32), val, labels


def ring_cluster_graph(N, K, seed=0, n_clusters=10):
    """Symmetric K-regular graph for LARGE K (config 5: perplexity list [10, 100] -> 300 neighbours per point), built without
    any sort: every cluster's members are put in a random cyclic order and each is linked to its K/2 successors and K/2
    predecessors.  Random in index space (like a kNN graph before any re-ordering), symmetric by construction, weight a
    symmetric hash of the pair, normalised to sum 1.  Returns (row u32, col u32, val f64-of-f32, labels)."""
    rng = np.random.default_rng(seed)
    labels = rng.integers(0, n_clusters, N).astype(np.int32)
    order = np.argsort(labels, kind="stable")
    starts = np.searchsorted(labels[order], np.arange(n_clusters + 1))
    half = K // 2
    offs = np.concatenate([np.arange(-half, 0), np.arange(1, half + 1)]).astype(np.int64)
    col = np.empty((N, 2 * half), np.uint32)
    val = np.empty((N, 2 * half), np.float32)
    for c in range(n_clusters):
        members = order[starts[c]:starts[c + 1]]
        members = members[rng.permutation(len(members))]              # random cyclic order
        n_c = len(members)
        for lo in range(0, n_c, 1 << 16):                               # bounded temporaries
            k = np.arange(lo, min(n_c, lo + (1 << 16)), dtype=np.int64)[:, None]
            kk = (k + offs[None, :]) % n_c
            a, b = np.minimum(k, kk), np.maximum(k, kk)
            h = (a * 2654435761 + b * 40503 + c * 97) & 0xFFFFFFFF   # symmetric in the pair
            col[members[k[:, 0]]] = members[kk]
            val[members[k[:, 0]]] = 0.1 + h.astype(np.float32) / np.float32(4294967296.0)
    val = val.reshape(-1).astype(np.float64)
    val = (val / val.sum()).astype(np.float32).astype(np.float64)
    row = (np.arange(N + 1, dtype=np.uint64) * (2 * half)).astype(np.uint32)
    return row, col.reshape(-1), val, labels


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


def write_data_dat(path, X, theta=0.5, perplexity=30.0, perplexity_list=None, no_dims=2, max_iter=750,
                   stop_lying_iter=250, mom_switch_iter=250, momentum=0.5, final_momentum=0.8, learning_rate=200.0,
                   max_step_norm=5.0, K=-1, sigma=-1.0, nbody_algo=2, knn_algo=1, early_exag_coeff=12.0,
                   no_momentum_during_exag=0, n_trees=50, search_k=-1, start_late_exag_iter=-1, late_exag_coeff=-1.0,
                   nterms=3, intervals_per_integer=1.0, min_num_intervals=50, seed=-1, df=1.0, load_affinities=0,
                   initialization=None):
    """Our own writer of the data.dat protocol (field order and packing of /root/reference/fast_tsne.py:259-297,
    read by src/tsne.cpp:1915-1985).  tests/test_protocol.py checks it byte for byte against the reference wrapper."""
    import struct
    X = np.ascontiguousarray(X, dtype=np.float64)
    n, d = X.shape
    with open(path, "wb") as f:
        f.write(struct.pack("=i", n)); f.write(struct.pack("=i", d))
        f.write(struct.pack("=d", theta)); f.write(struct.pack("=d", perplexity))
        if perplexity == 0:
            f.write(struct.pack("=i", len(perplexity_list)))
            for pp in perplexity_list:
                f.write(struct.pack("=d", pp))
        f.write(struct.pack("=i", no_dims)); f.write(struct.pack("=i", max_iter))
        f.write(struct.pack("=i", stop_lying_iter)); f.write(struct.pack("=i", mom_switch_iter))
        f.write(struct.pack("=d", momentum)); f.write(struct.pack("=d", final_momentum))
        f.write(struct.pack("=d", learning_rate)); f.write(struct.pack("=d", max_step_norm))
        f.write(struct.pack("=i", K)); f.write(struct.pack("=d", sigma))
        f.write(struct.pack("=i", nbody_algo)); f.write(struct.pack("=i", knn_algo))
        f.write(struct.pack("=d", early_exag_coeff)); f.write(struct.pack("=i", no_momentum_during_exag))
        f.write(struct.pack("=i", n_trees)); f.write(struct.pack("=i", search_k))
        f.write(struct.pack("=i", start_late_exag_iter)); f.write(struct.pack("=d", late_exag_coeff))
        f.write(struct.pack("=i", nterms)); f.write(struct.pack("=d", intervals_per_integer))
        f.write(struct.pack("=i", min_num_intervals))
        f.write(X.tobytes())
        f.write(struct.pack("=i", seed)); f.write(struct.pack("=d", df)); f.write(struct.pack("=i", load_affinities))
        if initialization is not None:
            f.write(np.ascontiguousarray(initialization, dtype=np.float64).tobytes())


def write_reference_inputs(dirname, row, col, val, Y0, max_iter, no_dims, learning_rate, stop_lying_iter,
                           mom_switch_iter, early_exag=12.0, df=1.0, nterms=3, ipi=1.0, min_int=50,
                           max_step_norm=5.0, start_late_exag_iter=-1, late_exag_coeff=-1.0, momentum=0.5,
                           final_momentum=0.8, no_momentum_during_exag=0):
    """data.dat + P_row/P_col/P_val.dat for `fast_tsne 1.2.1 data.dat result.dat <nthreads>` with load_affinities=1
    (the reference's own injection hook, src/tsne.cpp:236-281): X is an N x 1 dummy, kNN is bypassed."""
    import os
    os.makedirs(dirname, exist_ok=True)
    N = len(Y0)
    np.ascontiguousarray(row, np.uint32).tofile(os.path.join(dirname, "P_row.dat"))
    np.ascontiguousarray(col, np.uint32).tofile(os.path.join(dirname, "P_col.dat"))
    np.ascontiguousarray(val, np.float64).tofile(os.path.join(dirname, "P_val.dat"))
    write_data_dat(os.path.join(dirname, "data.dat"), np.zeros((N, 1)), theta=0.5, perplexity=-1.0, no_dims=no_dims,
                   max_iter=max_iter, stop_lying_iter=stop_lying_iter, mom_switch_iter=mom_switch_iter, momentum=momentum,
                   final_momentum=final_momentum, learning_rate=learning_rate, max_step_norm=max_step_norm, K=1, sigma=1.0,
                   nbody_algo=2, knn_algo=1, early_exag_coeff=early_exag, no_momentum_during_exag=no_momentum_during_exag,
                   n_trees=1, search_k=1, start_late_exag_iter=start_late_exag_iter, late_exag_coeff=late_exag_coeff,
                   nterms=nterms, intervals_per_integer=ipi, min_num_intervals=min_int, seed=42, df=df, load_affinities=1,
                   initialization=Y0)


def read_result(path):
    import struct
    with open(path, "rb") as f:
        n, d = struct.unpack("=ii", f.read(8))
        Y = np.frombuffer(f.read(8 * n * d), np.float64).reshape(n, d).copy()
        (m,) = struct.unpack("=i", f.read(4))
        costs = np.frombuffer(f.read(8 * m), np.float64).copy()
    return Y, costs
