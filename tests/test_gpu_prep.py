"""Device preprocessing (SURVEY section 8 f1 + f4): fitsne_knn and fitsne_similarities against the CPU statement of the
same arithmetic (host/tsne_host.cpp, itself checked against the reference's VP-tree path in tests/test_protocol.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fb():
    import fitsne_b200
    fitsne_b200.load_library()
    return fitsne_b200


def blobs(N, D, seed):
    rng = np.random.default_rng(seed)
    centres = rng.standard_normal((10, D)) * 10
    X = centres[rng.integers(0, 10, N)] + rng.standard_normal((N, D))
    X -= X.mean(0)
    return X / np.abs(X).max()


@pytest.mark.parametrize("N,D,K", [(3000, 50, 90), (1500, 7, 30), (700, 128, 300), (65, 3, 64)])
def test_knn_is_exact(fb, N, D, K):
    X = blobs(N, D, N)
    nbr, dist = fb.knn(X, K)
    # brute force in fp64
    d2 = ((X[:, None, :] - X[None, :, :]) ** 2).sum(-1) if N <= 3000 else None
    np.fill_diagonal(d2, np.inf)
    order = np.lexsort((np.broadcast_to(np.arange(N), (N, N)), d2), axis=1)[:, :K]      # by distance, ties by index
    ref_d = np.sqrt(np.take_along_axis(d2, order, 1))
    assert np.array_equal(nbr, order.astype(np.uint32))
    assert np.allclose(dist, ref_d, rtol=1e-13, atol=1e-15)
    assert np.all(np.diff(dist, axis=1) >= 0) and not np.any(nbr == np.arange(N)[:, None])


def test_knn_with_duplicate_points(fb):
    X = blobs(400, 5, 1)
    X[100:140] = X[7]                       # 40 copies of one point: zero distances, ties broken by index
    nbr, dist = fb.knn(X, 20)
    assert np.all(dist[100:140, :20] == 0) and np.all(dist[7, :20] == 0)
    assert np.array_equal(nbr[7], np.arange(100, 120, dtype=np.uint32))


@pytest.mark.parametrize("kw", [dict(perplexity=30.0), dict(perplexity=5.0), dict(perplexity_list=[10.0, 40.0]),
                                dict(perplexity=-1.0, K=25, sigma=0.3)])
def test_similarities_match_host_statement(fb, kw):
    X = blobs(2500, 20, 5) * 3.0
    r0, c0, v0 = fb.input_similarities(X, nthreads=8, **kw)                 # CPU statement (libfitsne_host.so)
    r1, c1, v1 = fb.input_similarities_device(X, **kw)
    assert np.array_equal(r0, r1) and np.array_equal(c0, c1)
    assert np.allclose(v1, v0, rtol=1e-9, atol=0)
    assert abs(v1.sum() - 1) < 1e-12
    # symmetric
    import scipy.sparse as sp
    A = sp.csr_matrix((v1, c1.astype(np.int64), r1.astype(np.int64)), shape=(len(r1) - 1,) * 2)
    assert abs(A - A.T).max() < 1e-18


def test_fast_tsne_end_to_end_uses_device_preprocessing(fb):
    """The in-process mirror of the reference wrapper, raw data in, embedding out: kNN + similarities + loop on the device;
    the clusters of the input must come out separated."""
    rng = np.random.default_rng(0)
    lab = rng.integers(0, 5, 4000)
    X = (rng.standard_normal((5, 30)) * 8)[lab] + rng.standard_normal((4000, 30))
    Y, loss = fb.fast_tsne(X, perplexity=30, max_iter=300, seed=1, return_loss=True)
    assert Y.shape == (4000, 2) and np.isfinite(Y).all()
    cen = np.stack([Y[lab == k].mean(0) for k in range(5)])
    spread = np.mean([Y[lab == k].std(0).mean() for k in range(5)])
    dmin = min(np.linalg.norm(cen[a] - cen[b]) for a in range(5) for b in range(a))
    assert dmin > 3 * spread
    assert np.nanmin(loss) < np.nanmax(loss)
