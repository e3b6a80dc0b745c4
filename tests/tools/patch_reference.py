"""INTEGRATION.md form B, executable: insert the C-ABI call into a COPY of the reference's TSNE::run (the copy lives in a
scratch directory; /root/reference is never written and no reference source enters the repository), compile it with the
reference's own flags and link it against libfitsne_b200.so.  `python tests/tools/patch_reference.py <outdir>` prints the
path of the patched binary (oracle/_ref-style build: FFTW replaced by the MKL shim, as for the unmodified reference).
Test infrastructure: tests/test_integration_patch.py runs this in the build container and, on a GPU box, runs the binary."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("FITSNE_REFERENCE", "/root/reference")

MARKER = "    // If we are doing early exaggeration, we pre-multiply all the P by the coefficient of early exaggeration"
INCLUDE_AFTER = '#include "time_code.h"'
BLOCK = r'''
    /* ---- fitsne_b200: the FFT-interpolation loop (tsne.cpp:389-577) runs on the B200 through the C ABI ------------- */
    if (!exact && nbody_algorithm == 2) {
        fitsne_config cfg = { nterms, intervals_per_integer, min_num_intervals, df, /*device*/ -1, /*flags*/ 0 };
        fitsne_schedule s = { max_iter, stop_lying_iter, mom_switch_iter, start_late_exag_iter, momentum,
                              final_momentum, learning_rate, early_exag_coeff, late_exag_coeff, max_step_norm,
                              no_momentum_during_exag ? 1 : 0, /*verbose*/ 1 };
        /* host CSR P (unsigned/unsigned/double) and host Y in; host Y and costs[] out */
        int rc = fitsne_run_host(&cfg, &s, N, no_dims, row_P, col_P, val_P, Y, costs);
        if (rc != 0) { printf("fitsne_b200: %s\n", fitsne_last_error(NULL)); return -100 + rc; }
        free(dY); free(uY); free(gains);
        free(row_P); free(col_P); free(val_P);
        return 0;
    }
'''


def make(outdir):
    os.makedirs(outdir, exist_ok=True)
    src = open(os.path.join(REF, "src", "tsne.cpp")).read()
    assert src.count(MARKER) == 1 and src.count(INCLUDE_AFTER) == 1, "the reference's tsne.cpp does not look like the surveyed one"
    src = src.replace(INCLUDE_AFTER, INCLUDE_AFTER + '\n#include "fitsne_b200.h"          /* include/fitsne_b200.h of this repository */')
    src = src.replace(MARKER, BLOCK + MARKER)
    patched = os.path.join(outdir, "tsne_patched.cpp")
    open(patched, "w").write(src)
    import torch
    torchlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    lib = os.path.join(ROOT, "fit-sne_b200", "lib")
    exe = os.path.join(outdir, "fast_tsne_patched")
    flags = ["-std=c++11", "-O3", "-pthread", "-w", "-I" + os.path.join(ROOT, "oracle", "ref_build"), "-I" + os.path.join(REF, "src"),
             "-I" + os.path.join(ROOT, "include")]
    cmd = ["g++"] + flags + [patched, os.path.join(REF, "src", "nbodyfft.cpp"), os.path.join(REF, "src", "sptree.cpp"),
                             os.path.join(ROOT, "oracle", "ref_build", "fftw_shim_mkl.cpp"), "-o", exe,
                             "-L" + lib, "-lfitsne_b200", "-Wl,-rpath," + lib, "-L/usr/local/cuda/lib64", "-Wl,-rpath-link,/usr/local/cuda/lib64",
                             "-Wl,-rpath,/usr/local/cuda/lib64", "-L" + torchlib, "-ltorch_cpu", "-lc10", "-Wl,-rpath," + torchlib, "-lm"]
    try:
        subprocess.check_call(cmd)
    finally:
        os.remove(patched)          # the patched copy of the reference source is a build intermediate: it is not kept
    return exe


if __name__ == "__main__":
    print(make(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "oracle", "_ref", "patched")))
