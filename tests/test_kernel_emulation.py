"""k_spread_chunks2 (one thread per chunk) is written as two phase functions; tests/tools/spread_emul.cu runs the very
same code on the host, thread by thread and phase by phase, and checks it bit for bit against a transcription of
k_spread_chunks' threads and, after the combine step, against a direct fp64 spread.  CPU only (nvcc as a host compiler)."""
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
    """fitsne_fft.cuh's Stockham stages (narrow plans = the shipped path, wide plans = radix 16/9, opt-in) for every FFT
    length of the grid ladder, emulated thread by thread and stage by stage, against a direct fp64 DFT."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "fft_emul")
    src = os.path.join(ROOT, "tests", "tools", "fft_emul.cu")
    subprocess.check_call([nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe, src], timeout=600)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "FFT_EMUL_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-1000:]


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="needs nvcc")
def test_sorted_spmv_host_emulation(tmp_path):
    """Column-sorted attractive term (opt-in FITSNE_FLAG_SORTED_SPMV): the layout kernels and the CTA phases of
    k_attract_sorted, emulated lane by lane, against a direct fp64 sum over the CSR; layout invariants checked too."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "spmv_emul")
    src = os.path.join(ROOT, "tests", "tools", "spmv_emul.cu")
    subprocess.check_call([nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe, src], timeout=600)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "SPMV_EMUL_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-1000:]
