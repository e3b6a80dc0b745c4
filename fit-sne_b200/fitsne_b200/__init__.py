"""fitsne_b200 -- ctypes binding of libfitsne_b200.so (include/fitsne_b200.h).

Python-side mirror of the C ABI that replaces the per-iteration loop of the reference's TSNE::run
(/root/reference/src/tsne.cpp:437-577).  This module holds no numerics: every call goes to the CUDA library,
and importing/using it without the built library or without a CUDA device raises -- there is no CPU fallback.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libfitsne_b200.so")

STEP_MOMENTUM_CLIP, STEP_MOMENTUM, STEP_PLAIN_GD = 0, 1, 2
FLAG_NO_GRAPH, FLAG_TIMERS, FLAG_NO_REORDER, FLAG_FORCE_TILES, FLAG_NO_TILES, FLAG_NO_SPECULATION = 1, 2, 4, 8, 16, 32
PHASES = ["bounds", "sort", "spread", "kernel_spectrum", "fft", "gather", "attract_update", "center", "kl",
          "collectives", "allgather"]

EXPORTS = [
    "fitsne_create", "fitsne_create_sharded", "fitsne_nccl_unique_id", "fitsne_destroy", "fitsne_last_error",
    "fitsne_set_Y", "fitsne_get_Y", "fitsne_set_optimizer_state", "fitsne_get_optimizer_state", "fitsne_gradient",
    "fitsne_step", "fitsne_kl", "fitsne_run", "fitsne_run_host", "fitsne_synchronize", "fitsne_get_stats",
    "fitsne_reset_stats", "fitsne_last_run_ms", "fitsne_debug_copy", "fitsne_version", "fitsne_prewarm",
    "fitsne_knn", "fitsne_similarities", "fitsne_free", "fitsne_prep_last_error",
    "fitsne_create_from_files", "fitsne_create_from_files_sharded", "fitsne_run_files",
]


class Config(ctypes.Structure):
    _fields_ = [("nterms", ctypes.c_int), ("intervals_per_integer", ctypes.c_double),
                ("min_num_intervals", ctypes.c_int), ("df", ctypes.c_double), ("device", ctypes.c_int),
                ("flags", ctypes.c_int)]


class StepParams(ctypes.Structure):
    _fields_ = [("exaggeration", ctypes.c_double), ("momentum", ctypes.c_double), ("learning_rate", ctypes.c_double),
                ("max_step_norm", ctypes.c_double), ("mode", ctypes.c_int)]


class Schedule(ctypes.Structure):
    _fields_ = [("max_iter", ctypes.c_int), ("stop_lying_iter", ctypes.c_int), ("mom_switch_iter", ctypes.c_int),
                ("start_late_exag_iter", ctypes.c_int), ("momentum", ctypes.c_double),
                ("final_momentum", ctypes.c_double), ("learning_rate", ctypes.c_double),
                ("early_exag_coeff", ctypes.c_double), ("late_exag_coeff", ctypes.c_double),
                ("max_step_norm", ctypes.c_double), ("no_momentum_during_exag", ctypes.c_int),
                ("verbose", ctypes.c_int)]


class Stats(ctypes.Structure):
    _fields_ = [("iterations", ctypes.c_uint64), ("kernel_launches", ctypes.c_uint64),
                ("graph_launches", ctypes.c_uint64), ("regrids", ctypes.c_uint64), ("n_boxes", ctypes.c_int),
                ("grid_side", ctypes.c_int), ("fft_side", ctypes.c_int), ("min_coord", ctypes.c_double),
                ("max_coord", ctypes.c_double), ("phase_ms", ctypes.c_double * 16), ("reorders", ctypes.c_uint64)]


class FitsneError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libfitsne_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def load_library(path=None):
    """dlopen the CUDA library; raises OSError if it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(path or os.environ.get("FITSNE_LIB") or LIB_PATH)      # FITSNE_LIB: A/B another build (diagnostics)
        lib.fitsne_last_error.restype = ctypes.c_char_p
        lib.fitsne_last_error.argtypes = [ctypes.c_void_p]
        lib.fitsne_version.restype = ctypes.c_char_p
        _lib = lib
    return _lib


def _dp(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def make_schedule(max_iter=750, stop_lying_iter=250, mom_switch_iter=250, start_late_exag_iter=-1, momentum=0.5,
                  final_momentum=0.8, learning_rate=200.0, early_exag_coeff=12.0, late_exag_coeff=-1.0,
                  max_step_norm=5.0, no_momentum_during_exag=False, verbose=False):
    """Defaults are those of the reference wrapper (fast_tsne.py:19-51)."""
    return Schedule(int(max_iter), int(stop_lying_iter), int(mom_switch_iter), int(start_late_exag_iter),
                    float(momentum), float(final_momentum), float(learning_rate), float(early_exag_coeff),
                    float(late_exag_coeff), float(max_step_norm), int(bool(no_momentum_during_exag)),
                    int(bool(verbose)))


def shard_range(N, rank, world):
    per = (N + world - 1) // world
    return rank * per, min(N, (rank + 1) * per)


class FitSNE:
    """Device-resident t-SNE optimiser state for a fixed CSR P (row_P, col_P, val_P as in tsne.cpp:168-170)."""

    def __init__(self, row_P, col_P, val_P, Y0, nterms=3, intervals_per_integer=1.0, min_num_intervals=50, df=1.0,
                 device=-1, flags=0, rank=0, world=1, nccl_id=None):
        self._h = ctypes.c_void_p()
        self._lib = load_library()
        Y0 = np.ascontiguousarray(Y0, dtype=np.float64)
        if Y0.ndim == 1:
            Y0 = Y0[:, None]
        self.N, self.no_dims = Y0.shape
        row = np.ascontiguousarray(row_P, dtype=np.uint32)
        if row.shape[0] != self.N + 1:
            raise ValueError("row_P must have N+1 entries")
        b, e = shard_range(self.N, rank, world)
        col = np.ascontiguousarray(col_P, dtype=np.uint32)
        val = np.ascontiguousarray(val_P, dtype=np.float64)
        if world > 1 and col.shape[0] == int(row[-1]):   # full arrays given: slice this rank's edges
            col = np.ascontiguousarray(col[row[b]:row[e]])
            val = np.ascontiguousarray(val[row[b]:row[e]])
        flags = int(flags) | int(os.environ.get("FITSNE_FLAGS", "0"))      # diagnostics: OR extra FLAG_* bits into every context
        cfg = Config(int(nterms), float(intervals_per_integer), int(min_num_intervals), float(df), int(device),
                     int(flags))
        idbuf = None
        if world > 1:
            idbuf = (ctypes.c_char * 128).from_buffer_copy(bytes(nccl_id))
        rc = self._lib.fitsne_create_sharded(ctypes.byref(cfg), self.N, self.no_dims, _dp(row),
                                             _dp(col) if col.size else None, _dp(val) if val.size else None,
                                             _dp(Y0), int(rank), int(world), int(b), int(e), idbuf,
                                             ctypes.byref(self._h))
        if rc != 0:
            raise FitsneError(rc, self._lib.fitsne_last_error(None).decode())
        self.row_begin, self.row_end = b, e

    # -- plumbing
    def _ck(self, rc):
        if rc != 0:
            raise FitsneError(rc, self._lib.fitsne_last_error(self._h).decode())

    def close(self):
        if self._h:
            self._lib.fitsne_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- state
    def set_Y(self, Y):
        Y = np.ascontiguousarray(Y, dtype=np.float64)
        assert Y.size == self.N * self.no_dims
        self._ck(self._lib.fitsne_set_Y(self._h, _dp(Y)))

    def get_Y(self):
        Y = np.empty((self.N, self.no_dims))
        self._ck(self._lib.fitsne_get_Y(self._h, _dp(Y)))
        return Y

    def set_optimizer_state(self, uY=None, gains=None):
        uY = None if uY is None else np.ascontiguousarray(uY, dtype=np.float64)
        gains = None if gains is None else np.ascontiguousarray(gains, dtype=np.float64)
        self._ck(self._lib.fitsne_set_optimizer_state(self._h, _dp(uY), _dp(gains)))

    def get_optimizer_state(self):
        uY = np.empty((self.N, self.no_dims))
        gains = np.empty((self.N, self.no_dims))
        self._ck(self._lib.fitsne_get_optimizer_state(self._h, _dp(uY), _dp(gains)))
        return uY, gains

    # -- hot path
    def gradient(self, exaggeration=1.0):
        """(dC, sum_Q) for the current Y -- computeFftGradient* of the reference."""
        dC = np.empty((self.N, self.no_dims))
        z = ctypes.c_double(0)
        self._ck(self._lib.fitsne_gradient(self._h, ctypes.c_double(exaggeration), _dp(dC), ctypes.byref(z)))
        return dC, z.value

    def step(self, exaggeration=1.0, momentum=0.5, learning_rate=200.0, max_step_norm=5.0, mode=STEP_MOMENTUM_CLIP):
        sp = StepParams(float(exaggeration), float(momentum), float(learning_rate), float(max_step_norm), int(mode))
        self._ck(self._lib.fitsne_step(self._h, ctypes.byref(sp)))

    def kl(self, exaggeration=1.0):
        c = ctypes.c_double(0)
        self._ck(self._lib.fitsne_kl(self._h, ctypes.c_double(exaggeration), ctypes.byref(c)))
        return c.value

    def run(self, schedule=None, fetch_Y=True, **kw):
        """The loop of TSNE::run; returns (Y or None, costs)."""
        s = schedule or make_schedule(**kw)
        costs = np.zeros(s.max_iter)
        Y = np.empty((self.N, self.no_dims)) if fetch_Y else None
        self._ck(self._lib.fitsne_run(self._h, ctypes.byref(s), _dp(costs), _dp(Y)))
        return Y, costs

    def prewarm(self, n_boxes_lo, n_boxes_hi):
        """Prepare twiddle tables / buffers for grids of n_boxes_lo..n_boxes_hi boxes per dimension ahead of the loop."""
        self._ck(self._lib.fitsne_prewarm(self._h, int(n_boxes_lo), int(n_boxes_hi)))

    def synchronize(self):
        self._ck(self._lib.fitsne_synchronize(self._h))

    def last_run_ms(self):
        ms = ctypes.c_double(0)
        self._ck(self._lib.fitsne_last_run_ms(self._h, ctypes.byref(ms)))
        return ms.value

    def stats(self):
        st = Stats()
        self._ck(self._lib.fitsne_get_stats(self._h, ctypes.byref(st)))
        d = {k: getattr(st, k) for k, _ in Stats._fields_ if k != "phase_ms"}
        d["phase_ms"] = {PHASES[i]: st.phase_ms[i] for i in range(len(PHASES))}
        return d

    def reset_stats(self):
        self._ck(self._lib.fitsne_reset_stats(self._h))

    def debug(self, what, dtype):
        need = ctypes.c_size_t(0)
        self._ck(self._lib.fitsne_debug_copy(self._h, what.encode(), None, 0, ctypes.byref(need)))
        out = np.empty(need.value // np.dtype(dtype).itemsize, dtype=dtype)
        self._ck(self._lib.fitsne_debug_copy(self._h, what.encode(), _dp(out), ctypes.c_size_t(out.nbytes),
                                             ctypes.byref(need)))
        return out


def run_files(directory, N, Y0, schedule=None, nterms=3, intervals_per_integer=1.0, min_num_intervals=50, df=1.0, device=-1,
              flags=0, **kw):
    """fitsne_run_files: P from <directory>/P_row.dat, P_col.dat, P_val.dat (the reference's load_affinities files), streamed
    to the device; host Y in, host Y + costs out."""
    lib = load_library()
    s = schedule or make_schedule(**kw)
    Y = np.array(Y0, dtype=np.float64, order="C")
    if Y.ndim == 1:
        Y = Y[:, None]
    costs = np.zeros(s.max_iter)
    cfg = Config(int(nterms), float(intervals_per_integer), int(min_num_intervals), float(df), int(device), int(flags))
    rc = lib.fitsne_run_files(ctypes.byref(cfg), ctypes.byref(s), (directory or "").encode(), int(N), Y.shape[1], _dp(Y), _dp(costs))
    if rc != 0:
        raise FitsneError(rc, lib.fitsne_last_error(None).decode())
    return Y, costs


def _kernel_times(self):
    """{kernel: (total ms, launches)} of the timers-mode steps so far (needs FITSNE_KTIMES=1 in the environment when the
    context is created, and FLAG_TIMERS): warm per-kernel device times, one CUDA event after every kernel."""
    txt = self.debug("ktimes", np.uint8).tobytes().decode()
    out = {}
    for line in txt.splitlines():
        name, ms, n = line.split("\t")
        out[name] = (float(ms), int(n))
    return out


FitSNE.kernel_times = _kernel_times


def run_host(row_P, col_P, val_P, Y0, schedule=None, nterms=3, intervals_per_integer=1.0, min_num_intervals=50,
             df=1.0, device=-1, flags=0, **kw):
    """fitsne_run_host: host CSR P + host Y in, host Y + costs out (the call a patched tsne.cpp makes)."""
    lib = load_library()
    s = schedule or make_schedule(**kw)
    Y = np.array(Y0, dtype=np.float64, order="C")
    if Y.ndim == 1:
        Y = Y[:, None]
    N, d = Y.shape
    row = np.ascontiguousarray(row_P, dtype=np.uint32)
    col = np.ascontiguousarray(col_P, dtype=np.uint32)
    val = np.ascontiguousarray(val_P, dtype=np.float64)
    costs = np.zeros(s.max_iter)
    cfg = Config(int(nterms), float(intervals_per_integer), int(min_num_intervals), float(df), int(device), int(flags))
    rc = lib.fitsne_run_host(ctypes.byref(cfg), ctypes.byref(s), N, d, _dp(row), _dp(col), _dp(val), _dp(Y), _dp(costs))
    if rc != 0:
        raise FitsneError(rc, lib.fitsne_last_error(None).decode())
    return Y, costs


# ------------------------------------------------------------------------------------------------------------------
# In-process mirror of the reference wrapper's fast_tsne() (fast_tsne.py:19-51): same arguments and defaults, but no
# data.dat / result.dat round trip and no subprocess -- host preprocessing by libfitsne_host.so (exact kNN, perplexity
# calibration, symmetrisation: the reference's own semantics, see host/tsne_host.cpp), gradient loop by libfitsne_b200.so.
HOSTLIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libfitsne_host.so")
_hostlib = None


def _load_hostlib():
    global _hostlib
    if _hostlib is None:
        _hostlib = ctypes.CDLL(HOSTLIB_PATH)
    return _hostlib


def knn(X, K, device=-1):
    """Exact Euclidean kNN on the device (fitsne_knn): (nbr u32[N,K], dist f64[N,K]), ascending, self excluded."""
    lib = load_library()
    X = np.ascontiguousarray(X, dtype=np.float64)
    N, D = X.shape
    nbr = np.empty((N, K), np.uint32)
    dist = np.empty((N, K), np.float64)
    rc = lib.fitsne_knn(_dp(X), N, D, int(K), int(device), _dp(nbr), _dp(dist))
    if rc != 0:
        lib.fitsne_prep_last_error.restype = ctypes.c_char_p
        raise FitsneError(rc, lib.fitsne_prep_last_error().decode())
    return nbr, dist


def similarities(nbr, dist, perplexity=30.0, sigma=-1.0, perplexity_list=None, device=-1):
    """Perplexity search + symmetrisation on the device (fitsne_similarities): CSR (row u32, col u32, val f64)."""
    lib = load_library()
    nbr = np.ascontiguousarray(nbr, dtype=np.uint32)
    dist = np.ascontiguousarray(dist, dtype=np.float64)
    N, K = nbr.shape
    if perplexity_list is not None:
        perplexity = 0.0
    pl = np.ascontiguousarray(perplexity_list if perplexity_list is not None else [0.0], np.float64)
    row, col, val = ctypes.POINTER(ctypes.c_uint)(), ctypes.POINTER(ctypes.c_uint)(), ctypes.POINTER(ctypes.c_double)()
    rc = lib.fitsne_similarities(_dp(nbr), _dp(dist), N, K, ctypes.c_double(perplexity), ctypes.c_double(sigma),
                                 len(pl) if perplexity_list is not None else 0, _dp(pl), int(device),
                                 ctypes.byref(row), ctypes.byref(col), ctypes.byref(val))
    if rc != 0:
        lib.fitsne_prep_last_error.restype = ctypes.c_char_p
        raise FitsneError(rc, lib.fitsne_prep_last_error().decode())
    r = np.ctypeslib.as_array(row, shape=(N + 1,)).copy()
    c = np.ctypeslib.as_array(col, shape=(int(r[-1]),)).copy()
    v = np.ctypeslib.as_array(val, shape=(int(r[-1]),)).copy()
    lib.fitsne_free.argtypes = [ctypes.c_void_p]
    for ptr in (row, col, val):
        lib.fitsne_free(ctypes.cast(ptr, ctypes.c_void_p))
    return r, c, v


def input_similarities_device(X, perplexity=30.0, K=-1, sigma=-1.0, perplexity_list=None, device=-1):
    """The reference's prologue (centre, max-abs normalise when a perplexity is used, K = 3 * perplexity; tsne.cpp:153-161,
    :287-304) followed by the device kNN + similarities: the CSR TSNE::run builds before the loop."""
    X = np.array(X, dtype=np.float64, order="C")
    X -= X.mean(0)
    if perplexity_list is not None:
        perplexity = 0.0
    if perplexity >= 0:
        X /= np.abs(X).max()
        K_use = int(3 * (perplexity if perplexity > 0 else max(perplexity_list)))
        sigma_use = -1.0
    else:
        K_use, sigma_use = int(K), float(sigma)
    nbr, dist = knn(X, K_use, device)
    return similarities(nbr, dist, perplexity, sigma_use, perplexity_list, device)


def input_similarities(X, perplexity=30.0, K=-1, sigma=-1.0, perplexity_list=None, nthreads=0):
    """CPU statement of the same (libfitsne_host.so, exact multi-threaded kNN): used by the protocol tests and as the
    checker of the device path."""
    lib = _load_hostlib()
    X = np.array(X, dtype=np.float64, order="C")
    X -= X.mean(0)
    if perplexity_list is not None:
        perplexity = 0.0
    if perplexity >= 0:
        X /= np.abs(X).max()
        K_use = int(3 * (perplexity if perplexity > 0 else max(perplexity_list)))
        sigma_use = -1.0
    else:
        K_use, sigma_use = int(K), float(sigma)
    row, col, val = ctypes.POINTER(ctypes.c_uint)(), ctypes.POINTER(ctypes.c_uint)(), ctypes.POINTER(ctypes.c_double)()
    pl = np.ascontiguousarray(perplexity_list if perplexity_list is not None else [0.0], np.float64)
    rc = lib.fitsne_host_similarities(_dp(X), X.shape[0], X.shape[1], ctypes.c_double(perplexity), K_use, ctypes.c_double(sigma_use),
                                      len(pl) if perplexity_list is not None else 0, _dp(pl), int(nthreads or os.cpu_count() or 1),
                                      ctypes.byref(row), ctypes.byref(col), ctypes.byref(val))
    if rc != 0:
        raise RuntimeError("input_similarities failed (%d)" % rc)
    N = X.shape[0]
    r = np.ctypeslib.as_array(row, shape=(N + 1,)).copy()
    c = np.ctypeslib.as_array(col, shape=(int(r[-1]),)).copy()
    v = np.ctypeslib.as_array(val, shape=(int(r[-1]),)).copy()
    for ptr in (row, col, val):
        lib.fitsne_host_free(ptr)
    return r, c, v


def fast_tsne(X, theta=0.5, perplexity=30, map_dims=2, max_iter=750, stop_early_exag_iter=250, K=-1, sigma=-1, nbody_algo="FFT",
              knn_algo="annoy", mom_switch_iter=250, momentum=0.5, final_momentum=0.8, learning_rate="auto", early_exag_coeff=12,
              no_momentum_during_exag=False, n_trees=50, search_k=None, start_late_exag_iter="auto", late_exag_coeff=-1, nterms=3,
              intervals_per_integer=1, min_num_intervals=50, seed=-1, initialization="pca", load_affinities=None,
              perplexity_list=None, df=1, return_loss=False, nthreads=-1, max_step_norm=5, device=-1):
    """Drop-in for the reference wrapper's fast_tsne() (same parameters, fast_tsne.py:19-51) running in-process on a B200."""
    if nbody_algo != "FFT" or theta == 0:
        raise ValueError("this build accelerates the FFT-interpolation path only (nbody_algo='FFT', theta > 0)")
    if map_dims not in (1, 2):
        raise ValueError("FFT interpolation scheme supports only 1 or 2 output dimensions")
    X = np.array(X).astype(float)
    N = X.shape[0]
    if learning_rate == "auto":
        learning_rate = max(200, N / early_exag_coeff)
    if start_late_exag_iter == "auto":
        start_late_exag_iter = stop_early_exag_iter if late_exag_coeff > 0 else -1
    if max_step_norm == "none":
        max_step_norm = -1
    rng = np.random.default_rng(None if seed == -1 else seed)
    if isinstance(initialization, str) and initialization == "pca":
        Xc = X - X.mean(0)
        # leading principal components by SVD of the (thin) data matrix, scaled to std 1e-4 like the wrapper
        U, S, _ = np.linalg.svd(Xc, full_matrices=False) if min(Xc.shape) <= 2000 else _randomized_svd(Xc, map_dims, rng)
        Y0 = U[:, :map_dims] * S[:map_dims]
        Y0 = Y0 / np.std(Y0[:, 0]) * 0.0001
    elif isinstance(initialization, str) and initialization == "random":
        Y0 = rng.standard_normal((N, map_dims)) * 0.0001
    else:
        Y0 = np.array(initialization).astype(float).reshape(N, map_dims)
    if sigma > 0 and K > 0:
        perplexity = -1
    if N - 1 < 3 * (perplexity if perplexity_list is None else max(perplexity_list)):
        raise ValueError("Perplexity too large for the number of data points!")
    if load_affinities == "load":
        row = np.fromfile("P_row.dat", np.uint32); col = np.fromfile("P_col.dat", np.uint32); val = np.fromfile("P_val.dat", np.float64)
    else:
        row, col, val = input_similarities_device(X, float(perplexity), K, sigma, perplexity_list, device)
        if load_affinities == "save":
            row.tofile("P_row.dat"); col.tofile("P_col.dat"); val.tofile("P_val.dat")
    Y, costs = run_host(row, col, val, Y0, nterms=nterms, intervals_per_integer=intervals_per_integer,
                        min_num_intervals=min_num_intervals, df=df, device=device, max_iter=max_iter,
                        stop_lying_iter=stop_early_exag_iter, mom_switch_iter=mom_switch_iter,
                        start_late_exag_iter=start_late_exag_iter, momentum=momentum, final_momentum=final_momentum,
                        learning_rate=learning_rate, early_exag_coeff=early_exag_coeff, late_exag_coeff=late_exag_coeff,
                        max_step_norm=max_step_norm, no_momentum_during_exag=no_momentum_during_exag)
    if return_loss:
        loss = costs.copy()
        loss[np.arange(1, max_iter + 1) % 50 > 0] = np.nan          # like the wrapper (fast_tsne.py:328)
        return Y, loss
    return Y


def _randomized_svd(A, k, rng, oversample=8, iters=4):
    """Small randomized range finder for the PCA initialisation of large inputs (the wrapper uses sklearn's arpack)."""
    Q = np.linalg.qr(A @ rng.standard_normal((A.shape[1], k + oversample)))[0]
    for _ in range(iters):
        Q = np.linalg.qr(A @ (A.T @ Q))[0]
    Ub, S, Vt = np.linalg.svd(Q.T @ A, full_matrices=False)
    return Q @ Ub, S, Vt
