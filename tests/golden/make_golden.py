"""Generate tests/golden/*.npz from the UNMODIFIED compiled reference (oracle/_ref/libfitsne_ref.so).

Run in the build container only (needs /root/reference to have been compiled by `make -C oracle ref`):
    MKL_NUM_THREADS=1 python tests/golden/make_golden.py
The reference ships no golden vectors of its own, so these are the pinned
known answers for the hot path: per-iteration gradient dC, sum_Q (Z), the KL value, and short
optimiser trajectories, all from the reference's own object code on seeded inputs.  Inputs are stored
as float32 (exactly representable, so the fp32 device and the fp64 reference see identical numbers).
"""
import os
import sys
import tempfile

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
from pyoracle import Reference  # noqa: E402


def knn_like_graph(rng, N, K, n_clusters=10):
    """Symmetric kNN-style CSR P: K random same-cluster neighbours per row, (A+A^T), sum to 1.
    Values are rounded to float32 so the device's fp32 copy is exact."""
    labels = rng.integers(0, n_clusters, N)
    order = np.argsort(labels, kind="stable")
    starts = np.searchsorted(labels[order], np.arange(n_clusters + 1))
    rows = np.repeat(np.arange(N), K)
    cols = np.empty(N * K, np.int64)
    for c in range(n_clusters):
        members = order[starts[c]:starts[c + 1]]
        idx = np.nonzero(labels[rows] == c)[0]
        cols[idx] = members[rng.integers(0, len(members), len(idx))]
    A = sp.csr_matrix((rng.random(N * K) + 0.1, (rows, cols)), shape=(N, N))
    A.setdiag(0)
    A.eliminate_zeros()
    A = (A + A.T).tocsr()
    A.sum_duplicates()
    A.sort_indices()
    val = (A.data / A.data.sum()).astype(np.float32)
    return A.indptr.astype(np.uint32), A.indices.astype(np.uint32), val, labels


def clustered_embedding(rng, labels, dims, span, n_clusters=10):
    centres = rng.uniform(-0.5, 0.5, (n_clusters, dims)) * span
    Y = centres[labels] + rng.standard_normal((len(labels), dims)) * span * 0.03
    return Y.astype(np.float32)


def main():
    R = Reference()
    rng = np.random.default_rng(20260101)
    N, K = 3000, 6
    row, col, val, labels = knn_like_graph(rng, N, K)
    np.savez_compressed(os.path.join(HERE, "graph_n3000.npz"), row=row, col=col, val=val, labels=labels.astype(np.uint8))
    val64 = val.astype(np.float64)

    cases = [
        # name, dims, df, nterms, span, ipi, min_int
        ("g2d_early", 2, 1.0, 3, 6e-4, 1.0, 50),
        ("g2d_mid", 2, 1.0, 3, 40.0, 1.0, 50),
        ("g2d_late", 2, 1.0, 3, 170.0, 1.0, 50),
        ("g2d_wide", 2, 1.0, 3, 230.0, 1.0, 50),     # n_boxes >= 200: raw value, no rounding list
        ("g2d_p5", 2, 1.0, 5, 60.0, 1.0, 50),
        ("g2d_p2_ipi2", 2, 1.0, 2, 90.0, 2.0, 30),
        ("g2d_df05", 2, 0.5, 3, 120.0, 1.0, 50),
        ("g2d_df100", 2, 100.0, 3, 50.0, 1.0, 50),
        ("g1d_early", 1, 1.0, 3, 6e-4, 1.0, 50),
        ("g1d_late", 1, 1.0, 3, 170.0, 1.0, 50),
        ("g1d_df05", 1, 0.5, 3, 300.0, 1.0, 50),
        ("g1d_df100_p4", 1, 100.0, 4, 80.0, 1.0, 50),
    ]
    out = {}
    for name, dims, df, nterms, span, ipi, min_int in cases:
        if span < 1e-2:
            Y = (rng.standard_normal((N, dims)) * 1e-4).astype(np.float32)
        else:
            Y = clustered_embedding(rng, labels, dims, span)
        Y64 = Y.astype(np.float64)
        dC, Z = R.gradient(Y64, row, col, val64, nterms=nterms, ipi=ipi, min_int=min_int, df=df)
        empty = np.zeros(N + 1, np.uint32)
        dC_rep, Z2 = R.gradient(Y64, empty, col[:1], val64[:1], nterms=nterms, ipi=ipi, min_int=min_int, df=df)
        assert Z == Z2
        kl = R.kl(Y64, row, col, val64, Z, df=df)
        out[name + "__Y"] = Y
        out[name + "__dC"] = dC
        out[name + "__dC_rep"] = dC_rep          # = -F_rep/Z (empty P)
        out[name + "__meta"] = np.array([dims, df, nterms, ipi, min_int, Z, kl], np.float64)
        print(name, "Z=%.6g kl=%.6g |dC|=%.4g" % (Z, kl, np.linalg.norm(dC)))
    np.savez_compressed(os.path.join(HERE, "gradients_n3000.npz"), **out)

    # short optimiser trajectories through the reference's TSNE::run (P via load_affinities=1)
    runs = [
        # name, dims, df, kwargs
        ("run2d_default", 2, 1.0, dict(max_iter=100, stop_lying_iter=40, mom_switch_iter=40, learning_rate=250.0,
                                       early_exag_coeff=12.0, max_step_norm=5.0)),
        ("run2d_nomom_late", 2, 1.0, dict(max_iter=60, stop_lying_iter=20, mom_switch_iter=30, learning_rate=100.0,
                                          early_exag_coeff=4.0, no_momentum_during_exag=True,
                                          start_late_exag_iter=40, late_exag_coeff=2.0, max_step_norm=-1.0)),
        ("run1d_df05", 1, 0.5, dict(max_iter=60, stop_lying_iter=25, mom_switch_iter=25, learning_rate=200.0,
                                    early_exag_coeff=12.0, max_step_norm=5.0)),
        ("run2d_df2_autoexag", 2, 2.0, dict(max_iter=50, stop_lying_iter=20, mom_switch_iter=20, learning_rate=500.0,
                                            early_exag_coeff=0.0, max_step_norm=5.0)),
    ]
    out = {}
    for name, dims, df, kw in runs:
        Y0 = (rng.standard_normal((N, dims)) * 1e-4).astype(np.float32)
        with tempfile.TemporaryDirectory() as td:
            Y, costs = R.run(Y0.astype(np.float64), row, col, val64, scratch_dir=td, df=df, nthreads=1, **kw)
        out[name + "__Y0"] = Y0
        out[name + "__Y"] = Y
        out[name + "__costs"] = costs
        keys = sorted(kw)
        out[name + "__kwkeys"] = np.array(keys)
        out[name + "__kwvals"] = np.array([float(kw[k]) for k in keys])
        out[name + "__meta"] = np.array([dims, df], np.float64)
        print(name, "costs", costs[costs != 0])
    np.savez_compressed(os.path.join(HERE, "runs_n3000.npz"), **out)


if __name__ == "__main__":
    main()
