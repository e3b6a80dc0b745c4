"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/fitsne_b200.h declares, and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "fitsne_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fitsne_[a-z_A-Z0-9]+)\s*\(", src)))


def test_header_declares_the_hot_path():
    names = declared_functions()
    for must in ("fitsne_create", "fitsne_gradient", "fitsne_step", "fitsne_kl", "fitsne_run", "fitsne_run_host",
                 "fitsne_destroy"):
        assert must in names


def test_library_exports_every_declared_symbol():
    import fitsne_b200
    lib = fitsne_b200.load_library()
    names = declared_functions()
    assert set(names) == set(fitsne_b200.EXPORTS)
    for n in names:
        assert hasattr(lib, n), n
    assert b"sm_100a" in lib.fitsne_version()


def test_struct_layouts_match_header():
    import fitsne_b200
    # natural C layout of the structs in the header (LP64)
    assert ctypes.sizeof(fitsne_b200.Config) == 40
    assert ctypes.sizeof(fitsne_b200.StepParams) == 40
    assert ctypes.sizeof(fitsne_b200.Schedule) == 72
    assert ctypes.sizeof(fitsne_b200.Stats) == 32 + 16 + 16 + 16 * 8 + 8


def _has_cuda():
    try:
        cudart = ctypes.CDLL("libcudart.so")
    except OSError:
        try:
            import torch
            return torch.cuda.is_available()
        except Exception:
            return False
    n = ctypes.c_int(0)
    return cudart.cudaGetDeviceCount(ctypes.byref(n)) == 0 and n.value > 0


@pytest.mark.skipif(_has_cuda(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback_without_device():
    import fitsne_b200
    row = np.arange(0, 11, dtype=np.uint32)
    col = np.arange(10, dtype=np.uint32)[::-1].copy()
    val = np.full(10, 0.1)
    with pytest.raises(fitsne_b200.FitsneError) as ei:
        fitsne_b200.FitSNE(row, col, val, np.random.randn(10, 2))
    assert ei.value.code == -2          # FITSNE_ENODEV
    assert "no CPU fallback" in str(ei.value)


def test_argument_errors_come_before_any_device_work():
    """Bad shard layouts are refused with FITSNE_EINVAL and a message, on any box (the checks precede device selection):
    more ranks than the peer tables hold, a rank outside the world, a rank whose ceil(N / world) block is empty."""
    import fitsne_b200
    lib = fitsne_b200.load_library()
    row = np.arange(0, 10, dtype=np.uint32)          # N = 9, one edge per row
    col = np.arange(9, dtype=np.uint32)[::-1].copy()
    val = np.full(9, 1.0 / 9)
    Y = np.random.default_rng(0).standard_normal((9, 2))
    cfg = fitsne_b200.Config(3, 1.0, 50, 1.0, 0, 0)
    ident = (ctypes.c_ubyte * 128)()

    def create(rank, world, b, e):
        ctx = ctypes.c_void_p()
        rc = lib.fitsne_create_sharded(ctypes.byref(cfg), 9, 2, row.ctypes.data_as(ctypes.c_void_p), col.ctypes.data_as(ctypes.c_void_p),
                                       val.ctypes.data_as(ctypes.c_void_p), Y.ctypes.data_as(ctypes.c_void_p), rank, world, b, e,
                                       ident, ctypes.byref(ctx))
        return rc, lib.fitsne_last_error(None).decode()

    rc, msg = create(0, 9, 0, 1)
    assert rc == -1 and "world size" in msg, (rc, msg)          # FITSNE_EINVAL
    rc, msg = create(4, 4, 0, 3)
    assert rc == -1 and "rank" in msg, (rc, msg)
    rc, msg = create(3, 4, 9, 9)                                 # ceil(9 / 4) = 3 rows per rank: rank 3 would own [9, 9)
    assert rc == -1 and "must own rows" in msg, (rc, msg)


def test_product_never_imports_the_oracle():
    """The product path must not route through oracle/ (it is test infrastructure)."""
    pkg = os.path.join(ROOT, "fit-sne_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".c")) or f == "Makefile":
                text = open(os.path.join(base, f), errors="ignore").read()
                assert "pyoracle" not in text and "fitsne_oracle" not in text and "libfitsne_ref" not in text, f
