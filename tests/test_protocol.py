"""The bin/fast_tsne file protocol (data.dat / result.dat / P_*.dat) and the host preprocessing, on the CPU.

The GPU box has neither /root/reference nor its wrappers, so the chain is closed in two halves:
here (build container) our writer is compared byte for byte with what the UNMODIFIED fast_tsne.py writes and our
result file is parsed by the reference wrapper's own reader; on the GPU box tests/test_gpu_binary.py drives
bin/fast_tsne with our writer."""
import ctypes
import importlib.util
import os
import struct
import subprocess
import sys

import numpy as np
import pytest

import bench_util
from conftest import ROOT

HOSTLIB = os.path.join(ROOT, "fit-sne_b200", "lib", "libfitsne_host.so")
REF_WRAPPER = "/root/reference/fast_tsne.py"
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "fast_tsne_ref")


@pytest.fixture(scope="module")
def hostlib():
    if not os.path.exists(HOSTLIB):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "fit-sne_b200"), "hostlib"])
    return ctypes.CDLL(HOSTLIB)


def parse(hostlib, path):
    ints = (ctypes.c_int * 20)()
    dbls = (ctypes.c_double * 16)()
    X, Y, pl = ctypes.POINTER(ctypes.c_double)(), ctypes.POINTER(ctypes.c_double)(), ctypes.POINTER(ctypes.c_double)()
    rc = hostlib.fitsne_host_parse(path.encode(), ints, dbls, ctypes.byref(X), ctypes.byref(Y), ctypes.byref(pl))
    assert rc == 0
    names_i = ["n", "d", "no_dims", "max_iter", "stop_lying_iter", "mom_switch_iter", "K", "nbody_algo", "knn_algo",
               "no_momentum_during_exag", "n_trees", "search_k", "start_late_exag_iter", "nterms", "min_num_intervals",
               "rand_seed", "load_affinities", "perplexity_list_length", "skip_random_init"]
    names_d = ["theta", "perplexity", "momentum", "final_momentum", "learning_rate", "max_step_norm", "sigma",
               "early_exag_coeff", "late_exag_coeff", "intervals_per_integer", "df"]
    out = {k: ints[i] for i, k in enumerate(names_i)}
    out.update({k: dbls[i] for i, k in enumerate(names_d)})
    n, d, nd = out["n"], out["d"], out["no_dims"]
    out["X"] = np.ctypeslib.as_array(X, shape=(n * d,)).copy().reshape(n, d)
    out["Y"] = np.ctypeslib.as_array(Y, shape=(n * nd,)).copy().reshape(n, nd)
    out["plist"] = np.ctypeslib.as_array(pl, shape=(out["perplexity_list_length"],)).copy() if out["perplexity_list_length"] else None
    return out


def test_data_dat_round_trip(hostlib, tmp_path):
    rng = np.random.default_rng(0)
    X = rng.standard_normal((37, 5))
    init = rng.standard_normal((37, 2))
    p = str(tmp_path / "data.dat")
    bench_util.write_data_dat(p, X, theta=0.5, perplexity=12.5, no_dims=2, max_iter=123, stop_lying_iter=45, mom_switch_iter=46,
                              momentum=0.4, final_momentum=0.9, learning_rate=321.0, max_step_norm=4.5, K=7, sigma=2.5,
                              nbody_algo=2, knn_algo=2, early_exag_coeff=11.0, no_momentum_during_exag=1, n_trees=13,
                              search_k=77, start_late_exag_iter=99, late_exag_coeff=3.5, nterms=4, intervals_per_integer=2.0,
                              min_num_intervals=33, seed=5, df=0.75, load_affinities=2, initialization=init)
    r = parse(hostlib, p)
    assert (r["n"], r["d"], r["no_dims"], r["max_iter"], r["stop_lying_iter"], r["mom_switch_iter"]) == (37, 5, 2, 123, 45, 46)
    assert (r["K"], r["nbody_algo"], r["knn_algo"], r["no_momentum_during_exag"], r["n_trees"], r["search_k"]) == (7, 2, 2, 1, 13, 77)
    assert (r["start_late_exag_iter"], r["nterms"], r["min_num_intervals"], r["rand_seed"], r["load_affinities"]) == (99, 4, 33, 5, 2)
    assert (r["theta"], r["perplexity"], r["momentum"], r["final_momentum"], r["learning_rate"], r["max_step_norm"]) == (0.5, 12.5, 0.4, 0.9, 321.0, 4.5)
    assert (r["sigma"], r["early_exag_coeff"], r["late_exag_coeff"], r["intervals_per_integer"], r["df"]) == (2.5, 11.0, 3.5, 2.0, 0.75)
    assert r["skip_random_init"] == 1 and np.array_equal(r["X"], X) and np.array_equal(r["Y"], init)


def test_data_dat_perplexity_list_and_missing_tail(hostlib, tmp_path):
    X = np.arange(12, dtype=np.float64).reshape(6, 2)
    p = str(tmp_path / "d.dat")
    bench_util.write_data_dat(p, X, perplexity=0, perplexity_list=[3.0, 10.0, 30.0], no_dims=1, initialization=None)
    r = parse(hostlib, p)
    assert r["perplexity"] == 0 and list(r["plist"]) == [3.0, 10.0, 30.0] and r["no_dims"] == 1
    assert r["skip_random_init"] == 0            # no initialisation appended (tsne.cpp:1976-1985)
    # old-style file that ends right after X: seed / df / load_affinities keep their defaults (0, 1.0, 0)
    raw = open(p, "rb").read()
    tail = 4 + 8 + 4
    open(p, "wb").write(raw[:-tail])
    r = parse(hostlib, p)
    assert (r["rand_seed"], r["df"], r["load_affinities"], r["skip_random_init"]) == (0, 1.0, 0, 0)
    # a partial initialisation is ignored (tsne.cpp:1977-1982)
    open(p, "wb").write(raw + struct.pack("=d", 1.0))
    assert parse(hostlib, p)["skip_random_init"] == 0


def test_result_dat_layout(hostlib, tmp_path):
    rng = np.random.default_rng(1)
    Y = rng.standard_normal((11, 2))
    costs = np.zeros(100)
    costs[49] = 3.5
    costs[99] = 2.25
    p = str(tmp_path / "result.dat")
    hostlib.fitsne_host_write_result(p.encode(), Y.ctypes.data_as(ctypes.c_void_p), costs.ctypes.data_as(ctypes.c_void_p), 11, 2, 100)
    raw = open(p, "rb").read()
    assert len(raw) == 4 + 4 + 8 * 22 + 4 + 8 * 100              # tsne.cpp:2031-2035
    assert struct.unpack("=ii", raw[:8]) == (11, 2)
    Y2, c2 = bench_util.read_result(p)
    assert np.array_equal(Y2, Y) and np.array_equal(c2, costs)


@pytest.mark.skipif(not os.path.exists(REF_WRAPPER), reason="reference wrapper only exists in the build container")
def test_writer_matches_unmodified_reference_wrapper(hostlib, tmp_path, monkeypatch):
    """Run the reference's fast_tsne() with subprocess.call stubbed out: capture the data file it writes, compare
    with our writer byte for byte, and let ITS reader parse a result file written by OUR save_data."""
    spec = importlib.util.spec_from_file_location("ref_fast_tsne", REF_WRAPPER)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rng = np.random.default_rng(2)
    X = rng.standard_normal((60, 8))
    init = rng.standard_normal((60, 2)) * 1e-4
    captured = {}

    def fake_call(argv):
        captured["argv"] = argv
        captured["data"] = open(argv[2], "rb").read()
        n, nd, max_iter = 60, 2, 100
        Y = np.arange(n * nd, dtype=np.float64).reshape(n, nd)
        costs = np.zeros(max_iter); costs[49] = 4.0; costs[99] = 3.0
        hostlib.fitsne_host_write_result(argv[3].encode(), Y.ctypes.data_as(ctypes.c_void_p),
                                         costs.ctypes.data_as(ctypes.c_void_p), n, nd, max_iter)
        return 0

    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(ref.subprocess, "call", fake_call)
    Yr, loss = ref.fast_tsne(X, perplexity=15, max_iter=100, stop_early_exag_iter=30, late_exag_coeff=2.0, start_late_exag_iter=30,
                             learning_rate=77.0, initialization=init, seed=9, df=0.8, nthreads=3, return_loss=True)
    assert captured["argv"][1] == "1.2.1" and captured["argv"][4] == "3"
    ours = str(tmp_path / "ours.dat")
    bench_util.write_data_dat(ours, X, theta=0.5, perplexity=15, no_dims=2, max_iter=100, stop_lying_iter=30, mom_switch_iter=250,
                              momentum=0.5, final_momentum=0.8, learning_rate=77.0, max_step_norm=5, K=-1, sigma=-1, nbody_algo=2,
                              knn_algo=1, early_exag_coeff=12, no_momentum_during_exag=0, n_trees=50, search_k=15 * 3 * 50,
                              start_late_exag_iter=30, late_exag_coeff=2.0, nterms=3, intervals_per_integer=1, min_num_intervals=50,
                              seed=9, df=0.8, load_affinities=0, initialization=init)
    assert open(ours, "rb").read() == captured["data"]
    assert np.array_equal(Yr, np.arange(120, dtype=np.float64).reshape(60, 2))
    assert loss[49] == 4.0 and loss[99] == 3.0 and np.isnan(loss[0])


def _similarities(hostlib, X, perplexity, K, sigma=-1.0, plist=None, threads=4):
    X = np.ascontiguousarray(X, np.float64)
    row, col, val = ctypes.POINTER(ctypes.c_uint)(), ctypes.POINTER(ctypes.c_uint)(), ctypes.POINTER(ctypes.c_double)()
    pl = np.ascontiguousarray(plist if plist is not None else [0.0], np.float64)
    rc = hostlib.fitsne_host_similarities(X.ctypes.data_as(ctypes.c_void_p), X.shape[0], X.shape[1], ctypes.c_double(perplexity), K,
                                          ctypes.c_double(sigma), len(pl) if plist is not None else 0, pl.ctypes.data_as(ctypes.c_void_p),
                                          threads, ctypes.byref(row), ctypes.byref(col), ctypes.byref(val))
    assert rc == 0
    N = X.shape[0]
    r = np.ctypeslib.as_array(row, shape=(N + 1,)).copy()
    c = np.ctypeslib.as_array(col, shape=(r[-1],)).copy()
    v = np.ctypeslib.as_array(val, shape=(r[-1],)).copy()
    for ptr in (row, col, val):
        hostlib.fitsne_host_free(ptr)
    return r, c, v


def test_input_similarities_properties(hostlib):
    import scipy.sparse as sp
    rng = np.random.default_rng(3)
    X = rng.standard_normal((300, 6))
    r, c, v = _similarities(hostlib, X, 10.0, 30)
    A = sp.csr_matrix((v, c, r), shape=(300, 300))
    assert abs(A.sum() - 1.0) < 1e-12 and abs(A - A.T).max() < 1e-15 and A.diagonal().max() == 0
    assert np.all(np.diff(r) >= 30)
    # row perplexity of the conditional distribution is the target: check through the entropy of a fresh calibration
    D = np.sqrt(((X[:, None, :] - X[None, :, :]) ** 2).sum(-1))
    nn = np.argsort(D[0])[1:31]
    assert set(nn) <= set(c[r[0]:r[1]])


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="compiled reference only exists where it was built")
def test_input_similarities_match_reference_vptree(hostlib, tmp_path):
    """Exact-kNN preprocessing vs the reference binary with knn_algo=2 (VP-tree, exact), P saved via load_affinities=2."""
    import scipy.sparse as sp
    rng = np.random.default_rng(4)
    N, D = 400, 7
    X = rng.standard_normal((N, D))
    for kw in (dict(perplexity=12.0), dict(perplexity=0, perplexity_list=[5.0, 20.0]), dict(perplexity=-1.0, K=9, sigma=0.7)):
        bench_util.write_data_dat(str(tmp_path / "data.dat"), X, max_iter=1, knn_algo=2, load_affinities=2, seed=1, **kw)
        out = subprocess.run([REF_BIN, "1.2.1", "data.dat", "result.dat", "2"], cwd=tmp_path, capture_output=True, text=True,
                             env=dict(os.environ, MKL_NUM_THREADS="1"))
        assert out.returncode == 0, out.stdout[-500:]
        rr = np.fromfile(tmp_path / "P_row.dat", np.uint32)
        rc_ = np.fromfile(tmp_path / "P_col.dat", np.uint32)
        rv = np.fromfile(tmp_path / "P_val.dat", np.float64)
        Aref = sp.csr_matrix((rv, rc_, rr), shape=(N, N))
        # same preprocessing of X as TSNE::run: zero-mean, then max-abs normalisation when a perplexity is used
        Xp = X - X.mean(0)
        if kw.get("perplexity", 1) >= 0:
            Xp = Xp / np.abs(Xp).max()
        perp = kw["perplexity"]
        plist = kw.get("perplexity_list")
        K = kw.get("K", int(3 * (perp if perp > 0 else max(plist or [0]))))
        r, c, v = _similarities(hostlib, Xp, perp, K, sigma=kw.get("sigma", -1.0), plist=plist)
        A = sp.csr_matrix((v, c, r), shape=(N, N))
        assert A.nnz == Aref.nnz
        assert abs(A - Aref).max() < 1e-12
