"""Pin the oracle (oracle/fitsne_oracle.c) against golden vectors produced by the unmodified compiled
reference (tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

GRAD_CASES = ["g2d_early", "g2d_mid", "g2d_late", "g2d_wide", "g2d_p5", "g2d_p2_ipi2", "g2d_df05", "g2d_df100",
              "g1d_early", "g1d_late", "g1d_df05", "g1d_df100_p4"]
RUN_CASES = ["run2d_default", "run2d_nomom_late", "run1d_df05", "run2d_df2_autoexag"]


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(b)


@pytest.mark.parametrize("name", GRAD_CASES)
def test_oracle_gradient_matches_reference_golden(oracle, golden_graph, golden_gradients, name):
    row, col, val, _ = golden_graph
    g = golden_gradients
    dims, df, nterms, ipi, min_int, Z, kl = g[name + "__meta"]
    Y = g[name + "__Y"].astype(np.float64)
    dC, z = oracle.gradient(Y, row, col, val.astype(np.float64), nterms=int(nterms), ipi=ipi, min_int=int(min_int), df=df)
    assert rel(dC, g[name + "__dC"]) < 1e-9          # fp64 restatement vs fp64 reference
    assert abs(z - Z) / Z < 1e-10
    empty = np.zeros(len(Y) + 1, np.uint32)
    dC_rep, z2 = oracle.gradient(Y, empty, col[:1], val[:1].astype(np.float64), nterms=int(nterms), ipi=ipi,
                                 min_int=int(min_int), df=df)
    assert z2 == z
    assert rel(dC_rep, g[name + "__dC_rep"]) < 1e-9
    assert abs(oracle.kl(Y, row, col, val.astype(np.float64), z, df=df) - kl) / abs(kl) < 1e-10


@pytest.mark.parametrize("name", RUN_CASES)
def test_oracle_run_matches_reference_golden(oracle, golden_graph, golden_runs, name):
    row, col, val, _ = golden_graph
    g = golden_runs
    dims, df = g[name + "__meta"]
    kw = {k: v for k, v in zip(g[name + "__kwkeys"], g[name + "__kwvals"])}
    for k in ("max_iter", "stop_lying_iter", "mom_switch_iter", "start_late_exag_iter"):
        if k in kw:
            kw[k] = int(kw[k])
    if "no_momentum_during_exag" in kw:
        kw["no_momentum_during_exag"] = bool(kw["no_momentum_during_exag"])
    Y, costs = oracle.run(g[name + "__Y0"].astype(np.float64), row, col, val.astype(np.float64), df=df, **kw)
    ref_costs = g[name + "__costs"]
    assert np.array_equal(costs != 0, ref_costs != 0)         # costs[] written only every 50th / last iteration
    nz = ref_costs != 0
    assert np.allclose(costs[nz], ref_costs[nz], rtol=1e-7)
    assert rel(Y, g[name + "__Y"]) < 1e-6


def test_grid_sizing_quirks(oracle):
    # 2-D: else-if min/max scan (tsne.cpp:1045-1048): the first x only ever sets max
    Y = np.array([[-5.0, 1.0], [2.0, 3.0], [0.5, -1.0]])
    mn, mx, B = oracle.grid(Y)
    assert (mn, mx) == (-1.0, 3.0)        # -5 is never considered for min
    assert B == 50
    # rounding list: span 97.x -> 100 ; >= 200 stays raw
    Y = np.array([[0.0, 0.0], [97.5, 1.0], [3.0, 2.0]])
    assert oracle.grid(Y)[2] == 100
    Y = np.array([[0.0, 0.0], [213.7, 1.0], [3.0, 2.0]])
    assert oracle.grid(Y)[2] == 213
    # 1-D: plain min/max, no list
    Y = np.array([[-5.0], [92.2], [3.0]])
    assert oracle.grid(Y) == (-5.0, 92.2, 97)


@pytest.mark.skipif(not os.path.exists("/root/reference/src/tsne.cpp") or
                    not os.path.exists(os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "_ref", "libfitsne_ref.so")),
                    reason="compiled reference only exists in the build container")
def test_oracle_matches_live_reference_random_case(oracle, golden_graph):
    from pyoracle import Reference
    R = Reference()
    row, col, val, _ = golden_graph
    rng = np.random.default_rng(7)
    for dims, df, scale in ((2, 1.0, 12.0), (1, 0.7, 60.0)):
        Y = (rng.standard_normal((3000, dims)) * scale).astype(np.float32).astype(np.float64)
        a, za = oracle.gradient(Y, row, col, val.astype(np.float64), df=df)
        b, zb = R.gradient(Y, row, col, val.astype(np.float64), df=df)
        assert rel(a, b) < 1e-9 and abs(za - zb) / zb < 1e-10
