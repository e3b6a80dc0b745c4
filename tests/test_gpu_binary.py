"""bin/fast_tsne (our host shell + CUDA loop) driven through the reference's file protocol on a GPU box."""
import os
import subprocess

import numpy as np
import pytest

import bench_util
from conftest import ROOT

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "bin", "fast_tsne")


def run_bin(cwd, threads=4):
    out = subprocess.run([BIN, "1.2.1", "data.dat", "result.dat", str(threads)], cwd=cwd, capture_output=True, text=True)
    return out


def test_binary_with_injected_affinities_matches_reference_golden(tmp_path, golden_graph, golden_runs):
    row, col, val, _ = golden_graph
    g = golden_runs
    name = "run2d_default"
    kw = {k: v for k, v in zip(g[name + "__kwkeys"], g[name + "__kwvals"])}
    bench_util.write_reference_inputs(str(tmp_path), row, col, val.astype(np.float64), g[name + "__Y0"].astype(np.float64),
                                      max_iter=int(kw["max_iter"]), no_dims=2, learning_rate=kw["learning_rate"],
                                      stop_lying_iter=int(kw["stop_lying_iter"]), mom_switch_iter=int(kw["mom_switch_iter"]),
                                      early_exag=kw["early_exag_coeff"], max_step_norm=kw["max_step_norm"])
    out = run_bin(tmp_path)
    assert out.returncode == 0, out.stdout[-800:] + out.stderr[-400:]
    assert "Iteration 50 (50 iterations in" in out.stdout and "Wrote the 3000 x 2 data matrix successfully." in out.stdout
    Y, costs = bench_util.read_result(str(tmp_path / "result.dat"))
    ref = g[name + "__costs"]
    assert np.array_equal(costs != 0, ref != 0)
    nz = ref != 0
    assert np.all(np.abs(costs[nz] - ref[nz]) / ref[nz] < 1e-2)      # north_star: final KL within 1 %
    assert Y.shape == (3000, 2) and abs(Y.mean()) < 1e-4


def test_binary_end_to_end_from_raw_data(tmp_path):
    """Full path: X -> (host) kNN + perplexity + symmetrise -> (B200) loop; clusters must come apart and KL must drop."""
    rng = np.random.default_rng(0)
    N, D, C = 2000, 20, 5
    labels = rng.integers(0, C, N)
    X = rng.standard_normal((C, D))[labels] * 6 + rng.standard_normal((N, D))
    init = rng.standard_normal((N, 2)) * 1e-4
    bench_util.write_data_dat(str(tmp_path / "data.dat"), X, perplexity=20.0, max_iter=300, stop_lying_iter=100, mom_switch_iter=100,
                              learning_rate=200.0, knn_algo=2, seed=3, initialization=init, load_affinities=2)
    out = run_bin(tmp_path)
    assert out.returncode == 0, out.stdout[-800:] + out.stderr[-400:]
    Y, costs = bench_util.read_result(str(tmp_path / "result.dat"))
    kl = costs[costs != 0]
    assert len(kl) == 6 and kl[-1] < kl[0] and kl[-1] < 2.5
    cent = np.stack([Y[labels == c].mean(0) for c in range(C)])
    within = np.mean([np.linalg.norm(Y[labels == c] - cent[c], axis=1).mean() for c in range(C)])
    between = np.mean([np.linalg.norm(cent[a] - cent[b]) for a in range(C) for b in range(a + 1, C)])
    assert between > 3 * within
    # load_affinities=2 wrote the P files in the reference's layout (tsne.cpp:334-358)
    r = np.fromfile(tmp_path / "P_row.dat", np.uint32)
    assert len(r) == N + 1 and np.fromfile(tmp_path / "P_col.dat", np.uint32).size == r[-1]
    assert abs(np.fromfile(tmp_path / "P_val.dat", np.float64).sum() - 1) < 1e-9


def test_binary_rejects_bad_version_and_unsupported_modes(tmp_path):
    out = subprocess.run([BIN, "1.0.0"], cwd=tmp_path, capture_output=True, text=True)
    assert out.returncode != 0 and "wrong version number" in out.stdout
    X = np.random.default_rng(1).standard_normal((100, 4))
    bench_util.write_data_dat(str(tmp_path / "data.dat"), X, perplexity=10.0, nbody_algo=1, max_iter=10)
    out = run_bin(tmp_path)
    assert out.returncode == 2 and "FFT-interpolation path only" in out.stdout
    bench_util.write_data_dat(str(tmp_path / "data.dat"), X, perplexity=40.0, max_iter=10)
    out = run_bin(tmp_path)
    assert out.returncode == 1 and "Perplexity too large" in out.stdout       # tsne.cpp:128-131


def test_in_process_fast_tsne_matches_binary(tmp_path):
    """fitsne_b200.fast_tsne() (same signature as the reference wrapper, no files) == bin/fast_tsne on the same input."""
    import fitsne_b200 as fb
    rng = np.random.default_rng(5)
    N, D, C = 1500, 12, 4
    labels = rng.integers(0, C, N)
    X = rng.standard_normal((C, D))[labels] * 5 + rng.standard_normal((N, D))
    init = rng.standard_normal((N, 2)) * 1e-4
    kw = dict(perplexity=15, max_iter=200, stop_early_exag_iter=80, mom_switch_iter=80, learning_rate=150.0, late_exag_coeff=1.5,
              knn_algo="vp-tree", df=0.9, initialization=init)
    Y, loss = fb.fast_tsne(X, return_loss=True, **kw)
    bench_util.write_data_dat(str(tmp_path / "data.dat"), X, perplexity=15.0, max_iter=200, stop_lying_iter=80, mom_switch_iter=80,
                              learning_rate=150.0, start_late_exag_iter=80, late_exag_coeff=1.5, knn_algo=2, df=0.9, seed=-1,
                              initialization=init, search_k=15 * 3 * 50)
    out = run_bin(tmp_path)
    assert out.returncode == 0, out.stdout[-600:]
    Yb, costs = bench_util.read_result(str(tmp_path / "result.dat"))
    nz = costs != 0
    assert np.array_equal(np.isfinite(loss), nz)
    assert np.allclose(loss[nz], costs[nz], rtol=1e-9)
    assert np.allclose(Y, Yb, rtol=0, atol=1e-9 * np.abs(Yb).max())
    # PCA / random initialisations and the 1-D path at least run and separate the clusters
    Y1 = fb.fast_tsne(X, perplexity=15, max_iter=250, map_dims=1, seed=3)
    assert Y1.shape == (N, 1) and np.isfinite(Y1).all()


def test_affinity_files_are_streamed_to_the_device(tmp_path):
    """P_row/P_col/P_val.dat as a first-class input -- from an explicit directory (C ABI / Python mirror), from
    $FITSNE_AFFINITIES_DIR, and through the binary's load_affinities=1 (current directory, like the reference): all three
    give the run that the same P gives when handed over as host arrays."""
    import fitsne_b200 as fb
    N = 20000
    row, col, val, labels = bench_util.knn_like_graph(N, 12, seed=4)
    Y0 = bench_util.clustered_embedding(labels, 2, 40.0, seed=2)
    pdir = tmp_path / "affinities"
    pdir.mkdir()
    row.tofile(pdir / "P_row.dat"); col.tofile(pdir / "P_col.dat"); val.tofile(pdir / "P_val.dat")
    kw = dict(max_iter=100, stop_lying_iter=30, mom_switch_iter=30, learning_rate=800.0, early_exag_coeff=6.0)
    Yh, ch = fb.run_host(row, col, val, Y0, **kw)
    Yf, cf = fb.run_files(str(pdir), N, Y0, **kw)
    assert np.array_equal(Yh, Yf) and np.array_equal(ch, cf)
    os.environ["FITSNE_AFFINITIES_DIR"] = str(pdir)
    try:
        Ye, ce = fb.run_files(None, N, Y0, **kw)
    finally:
        del os.environ["FITSNE_AFFINITIES_DIR"]
    assert np.array_equal(Yh, Ye) and np.array_equal(ch, ce)
    with pytest.raises(fb.FitsneError):
        fb.run_files(str(tmp_path / "nowhere"), N, Y0, **kw)
    (pdir / "P_val.dat").write_bytes(val[:100].tobytes())                 # truncated: loud failure, not garbage
    with pytest.raises(fb.FitsneError):
        fb.run_files(str(pdir), N, Y0, **kw)
    # the binary: load_affinities=1 reads ./P_*.dat like the reference (tsne.cpp:236-281)
    row.tofile(tmp_path / "P_row.dat"); col.tofile(tmp_path / "P_col.dat"); val.tofile(tmp_path / "P_val.dat")
    bench_util.write_data_dat(str(tmp_path / "data.dat"), np.zeros((N, 1)), perplexity=-1.0, K=1, sigma=1.0, max_iter=100, stop_lying_iter=30,
                              mom_switch_iter=30, learning_rate=800.0, early_exag_coeff=6.0, load_affinities=1, initialization=Y0, seed=1)
    out = run_bin(tmp_path)
    assert out.returncode == 0, out.stdout[-600:]
    Yb, cb = bench_util.read_result(str(tmp_path / "result.dat"))
    assert np.array_equal(Yb, Yh) and np.array_equal(cb, ch)
