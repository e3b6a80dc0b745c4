"""Sharded zero-mean + bounds: the per-rank record scheme of k_shard_stats / k_center_shard (modelled in
tests/device_model.py) gives exactly the bounds of the reference's scan (tsne.cpp:1045-1048) on the centred embedding.
CPU only; the GPU side of the same property is tests/tools/mgpu_check.py (sharded vs single-GPU runs)."""
import numpy as np
import pytest

from device_model import SHARD_HEAD, centre32, literal_bounds_scan, shard_bounds, shard_records


def _check(Y, world):
    N, d = Y.shape
    recs = shard_records(Y, world)
    mean, bmn, bmx = shard_bounds(recs, N, d)
    Yc = np.stack([centre32(Y[:, k], mean[k]) for k in range(d)], 1)
    if d == 2:
        mn_ref, mx_ref = literal_bounds_scan(Yc.reshape(-1))
    else:
        mn_ref, mx_ref = Yc.min(), Yc.max()
    assert np.float32(bmx) == np.float32(mx_ref)
    assert np.float32(bmn) == np.float32(mn_ref)


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("d", [1, 2])
def test_random_embeddings(world, d):
    rng = np.random.default_rng(world * 10 + d)
    for N in (24, 1000, 4099):          # every rank must own at least one point (fitsne_create_sharded rejects empty shards)
        Y = (rng.standard_normal((N, d)) * rng.uniform(0.01, 50) + rng.uniform(-3, 3)).astype(np.float32)
        _check(Y, world)


@pytest.mark.parametrize("k", list(range(1, 2 * SHARD_HEAD + 1)))
def test_ascending_prefix_of_every_resolvable_length(k):
    """first k interleaved values strictly ascending AND below everything else: the scan must skip exactly those"""
    rng = np.random.default_rng(k)
    N = 500
    Y = rng.uniform(0.0, 10.0, (N, 2)).astype(np.float32)
    flat = Y.reshape(-1)
    flat[:k] = np.linspace(-100.0, -90.0, k, dtype=np.float32) + 0.0      # global minima, in ascending order
    if k < flat.size:
        flat[k] = np.float32(-95.0) if k > 1 else np.float32(-200.0)      # breaks the ascent (or undercuts a 1-prefix)
    Y = flat.reshape(N, 2)
    # the x/y column means differ, so the ascent has to hold after centring: check with the literal scan either way
    for world in (2, 4):
        _check(Y, world)


def test_ties_and_tiny_first_shard():
    Y = np.array([[1, 1], [1, 2], [0, 3], [5, -1], [2, 2], [7, 0]], np.float32)       # equal values end the strict ascent
    for world in (2, 3, 6):
        _check(Y, world)
    rng = np.random.default_rng(0)
    _check(rng.standard_normal((16, 2)).astype(np.float32), 8)                         # rank 0 owns 2 points only


def test_prefix_longer_than_the_head_is_cut_at_the_head():
    """documented limit: a strictly ascending prefix longer than 2*SHARD_HEAD values is treated as ending there"""
    N = 64
    flat = np.tile(np.array([23.0, 21.0]), N)                                           # every other point: (23, 21)
    flat[:24] = np.arange(24) - 100.0                                                   # 24 ascending values, the smallest of all
    flat[24] = -80.0                                                                    # ends the ascent: the reference's min
    Y = flat.reshape(N, 2)
    Y[-1] = -Y[:-1].sum(0)                                                              # integer column sums = 0: centring is the identity
    assert Y[-1].min() > -70
    Y = Y.astype(np.float32)
    recs = shard_records(Y, 2)
    mean, bmn, bmx = shard_bounds(recs, N, 2)
    assert np.all(mean == 0)
    mn_ref, mx_ref = literal_bounds_scan(Y.reshape(-1))
    assert mn_ref == -80.0                                                              # indices 0..23 never reach `min`
    assert np.float32(bmn) == np.float32(-84.0) and np.float32(bmx) == np.float32(mx_ref)   # head of 16 values: index 16 is seen
