"""ctypes bindings for the oracle -- TEST INFRASTRUCTURE ONLY.

Two checkers live here:
  * ``Oracle``    -- oracle/libfitsne_oracle.so, our fp64 C restatement (fitsne_oracle.c)
  * ``Reference`` -- oracle/_ref/libfitsne_ref.so, the unmodified reference compiled by
                     ``make -C oracle ref`` (only where /root/reference exists; the built
                     .so travels to the GPU box)
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (fit-sne_b200/) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libfitsne_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libfitsne_ref.so")
REF_BIN = os.path.join(HERE, "_ref", "fast_tsne_ref")
REF_BIN_TIMED = os.path.join(HERE, "_ref", "fast_tsne_ref_timed")

_c_d = ctypes.c_double
_c_i = ctypes.c_int
_vp = ctypes.c_void_p


def build(ref=None):
    """Compile the C restatement (always) and oracle/_ref (when the reference tree is present)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if ref is None:
        ref = os.path.isdir("/root/reference/src")
    if ref:
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def _p(a):
    return a.ctypes.data_as(_vp)


def _csr(row_P, col_P, val_P):
    row = np.ascontiguousarray(row_P, dtype=np.uint32)
    col = np.ascontiguousarray(col_P, dtype=np.uint32)
    val = np.ascontiguousarray(val_P, dtype=np.float64)
    if col.size == 0:  # keep pointers valid for an empty graph
        col = np.zeros(1, np.uint32)
        val = np.zeros(1, np.float64)
    return row, col, val


class Oracle:
    def __init__(self, path=ORACLE_SO):
        if not os.path.exists(path):
            build(ref=False)
        self.lib = ctypes.CDLL(path)
        self.lib.fitsne_oracle_kl.restype = _c_d

    def grid(self, Y, ipi=1.0, min_int=50):
        Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(len(Y), -1)
        out = np.zeros(3)
        self.lib.fitsne_oracle_grid(_c_i(Y.shape[0]), _c_i(Y.shape[1]), _p(Y), _c_d(ipi), _c_i(min_int), _p(out))
        return float(out[0]), float(out[1]), int(out[2])

    def gradient(self, Y, row_P, col_P, val_P, nterms=3, ipi=1.0, min_int=50, df=1.0):
        Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(len(Y), -1)
        row, col, val = _csr(row_P, col_P, val_P)
        dC = np.zeros_like(Y)
        z = _c_d(0)
        rc = self.lib.fitsne_oracle_gradient(_c_i(Y.shape[0]), _c_i(Y.shape[1]), _p(row), _p(col), _p(val), _p(Y),
                                             _p(dC), _c_i(nterms), _c_d(ipi), _c_i(min_int), _c_d(df),
                                             ctypes.byref(z))
        if rc != 0:
            raise RuntimeError("fitsne_oracle_gradient rc=%d" % rc)
        return dC, z.value

    def kl(self, Y, row_P, col_P, val_P, sum_Q, df=1.0):
        Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(len(Y), -1)
        row, col, val = _csr(row_P, col_P, val_P)
        return self.lib.fitsne_oracle_kl(_c_i(Y.shape[0]), _c_i(Y.shape[1]), _p(row), _p(col), _p(val), _p(Y),
                                         _c_d(sum_Q), _c_d(df))

    def step(self, Y, uY, gains, dY, mode, momentum, learning_rate, max_step_norm):
        """In place on Y, uY, gains (float64, C-contiguous)."""
        for a in (Y, uY, gains):
            assert a.dtype == np.float64 and a.flags.c_contiguous
        dY = np.ascontiguousarray(dY, dtype=np.float64)
        N = Y.shape[0]
        d = Y.size // N
        self.lib.fitsne_oracle_step(_c_i(N), _c_i(d), _p(Y), _p(uY), _p(gains), _p(dY), _c_i(mode), _c_d(momentum),
                                    _c_d(learning_rate), _c_d(max_step_norm))

    def run(self, Y0, row_P, col_P, val_P, max_iter, stop_lying_iter=250, mom_switch_iter=250, momentum=0.5,
            final_momentum=0.8, learning_rate=200.0, early_exag_coeff=12.0, no_momentum_during_exag=False,
            start_late_exag_iter=-1, late_exag_coeff=-1.0, nterms=3, ipi=1.0, min_int=50, df=1.0,
            max_step_norm=5.0):
        Y = np.array(Y0, dtype=np.float64, order="C").reshape(len(Y0), -1)
        row, col, val = _csr(row_P, col_P, val_P)
        val = val.copy()
        costs = np.zeros(max_iter)
        rc = self.lib.fitsne_oracle_run(_c_i(Y.shape[0]), _c_i(Y.shape[1]), _p(row), _p(col), _p(val), _p(Y),
                                        _c_i(max_iter), _c_i(stop_lying_iter), _c_i(mom_switch_iter), _c_d(momentum),
                                        _c_d(final_momentum), _c_d(learning_rate), _c_d(early_exag_coeff), _p(costs),
                                        _c_i(int(no_momentum_during_exag)), _c_i(start_late_exag_iter),
                                        _c_d(late_exag_coeff), _c_i(nterms), _c_d(ipi), _c_i(min_int), _c_d(df),
                                        _c_d(max_step_norm))
        if rc != 0:
            raise RuntimeError("fitsne_oracle_run rc=%d" % rc)
        return Y, costs


class Reference:
    """The unmodified reference's object code (tsne.cpp / nbodyfft.cpp) behind ref_harness.cpp."""

    def __init__(self, path=REF_SO):
        os.environ.setdefault("MKL_NUM_THREADS", "1")  # reference FFTs are single-threaded
        self.lib = ctypes.CDLL(path)
        self.lib.ref_fft_gradient.restype = _c_d
        self.lib.ref_kl_fft.restype = _c_d

    @staticmethod
    def available(path=REF_SO):
        return os.path.exists(path)

    def gradient(self, Y, row_P, col_P, val_P, nterms=3, ipi=1.0, min_int=50, df=1.0, nthreads=1):
        Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(len(Y), -1)
        row, col, val = _csr(row_P, col_P, val_P)
        dC = np.zeros_like(Y)
        z = self.lib.ref_fft_gradient(_c_i(Y.shape[0]), _c_i(Y.shape[1]), _p(row), _p(col), _p(val), _p(Y), _p(dC),
                                      _c_i(nterms), _c_d(ipi), _c_i(min_int), ctypes.c_uint(nthreads), _c_d(df))
        return dC, z

    def kl(self, Y, row_P, col_P, val_P, sum_Q, df=1.0):
        Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(len(Y), -1)
        row, col, val = _csr(row_P, col_P, val_P)
        return self.lib.ref_kl_fft(_c_i(Y.shape[0]), _c_i(Y.shape[1]), _p(row), _p(col), _p(val), _p(Y), _c_d(sum_Q),
                                   ctypes.c_uint(1), _c_d(df))

    def run(self, Y0, row_P, col_P, val_P, max_iter, scratch_dir, stop_lying_iter=250, mom_switch_iter=250,
            momentum=0.5, final_momentum=0.8, learning_rate=200.0, early_exag_coeff=12.0,
            no_momentum_during_exag=False, start_late_exag_iter=-1, late_exag_coeff=-1.0, nterms=3, ipi=1.0,
            min_int=50, df=1.0, max_step_norm=5.0, nthreads=1):
        Y = np.array(Y0, dtype=np.float64, order="C").reshape(len(Y0), -1)
        row, col, val = _csr(row_P, col_P, val_P)
        os.makedirs(scratch_dir, exist_ok=True)
        row.tofile(os.path.join(scratch_dir, "P_row.dat"))
        col.tofile(os.path.join(scratch_dir, "P_col.dat"))
        val.tofile(os.path.join(scratch_dir, "P_val.dat"))
        costs = np.zeros(max_iter)
        rc = self.lib.ref_run_with_P(scratch_dir.encode(), _c_i(Y.shape[0]), _p(Y), _c_i(Y.shape[1]), _c_i(max_iter),
                                     _c_i(stop_lying_iter), _c_i(mom_switch_iter), _c_d(momentum),
                                     _c_d(final_momentum), _c_d(learning_rate), _c_d(early_exag_coeff), _p(costs),
                                     _c_i(int(no_momentum_during_exag)), _c_i(start_late_exag_iter),
                                     _c_d(late_exag_coeff), _c_i(nterms), _c_d(ipi), _c_i(min_int),
                                     ctypes.c_uint(nthreads), _c_d(df), _c_d(max_step_norm))
        if rc != 0:
            raise RuntimeError("ref_run_with_P rc=%d" % rc)
        return Y, costs
