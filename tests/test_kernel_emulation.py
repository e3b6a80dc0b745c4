"""The per-thread bodies of the spread and of the FFT stages are __host__ __device__ phase functions; the emulators under
tests/tools run the very same code on the host, thread by thread and phase by phase (what the barriers enforce), and check
it against direct fp64 computations.  CPU only (nvcc as a host compiler)."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="needs nvcc")
def test_spread2_host_emulation(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "spread_emul")
    src = os.path.join(ROOT, "tests", "tools", "spread_emul.cu")
    subprocess.check_call([nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe, src], timeout=600)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "SPREAD_EMUL_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-1000:]


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="needs nvcc")
def test_fft_stages_host_emulation(tmp_path):
    """fitsne_fft.cuh's Stockham stages (row passes, 1-D lines) and fitsne_conv.cuh's in-place column transform (forward
    with zero substitution -> digit-reversed spectra -> inverse of 3 or 4 slots) for every FFT length of the grid ladder,
    emulated thread by thread and stage by stage, against a direct fp64 DFT."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "fft_emul")
    src = os.path.join(ROOT, "tests", "tools", "fft_emul.cu")
    subprocess.check_call([nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe, src], timeout=600)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "FFT_EMUL_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-1000:]
