/*
 * FFTW -> MKL DFTI shim -- TEST INFRASTRUCTURE ONLY (oracle build).
 *
 * Implements the five FFTW functions the reference's hot path calls (see
 * fftw3.h in this directory) with MKL's DFTI interface.  MKL is not installed
 * as a development package, but torch's libtorch_cpu.so exports the DFTI
 * symbols, so the prototypes and the handful of DFTI enum constants used are
 * declared by hand below (values from the public mkl_dfti.h).
 *
 * Run with MKL_NUM_THREADS=1 to mirror the reference, which never enables
 * FFTW threading (plans are plain fftw_plan_dft_* + FFTW_ESTIMATE).
 */
#include <cstdio>
#include <cstdlib>
#include "fftw3.h"

extern "C" {
typedef void *DFTI_HANDLE;
long DftiCreateDescriptor_d_1d(DFTI_HANDLE *, int domain, long n);
long DftiCreateDescriptor_d_md(DFTI_HANDLE *, int domain, long dim, long *n);
long DftiSetValue(DFTI_HANDLE, int param, ...);
long DftiCommitDescriptor(DFTI_HANDLE);
long DftiComputeForward(DFTI_HANDLE, void *, ...);
long DftiComputeBackward(DFTI_HANDLE, void *, ...);
long DftiFreeDescriptor(DFTI_HANDLE *);
}

enum {
    kConjEvenStorage = 10, kPlacement = 11, kInStrides = 12, kOutStrides = 13,
    kComplex = 32, kReal = 33, kComplexComplex = 39, kInplace = 43, kNotInplace = 44
};

struct oracle_fftw_plan_s {
    DFTI_HANDLE h;
    void *in;
    void *out;
    bool forward;
};

static void must(long status, const char *what) {
    if (status != 0) {
        fprintf(stderr, "fftw_shim_mkl: %s failed with DFTI status %ld\n", what, status);
        abort();
    }
}

static fftw_plan real_2d(int n0, int n1, void *in, void *out, bool forward) {
    auto *p = new oracle_fftw_plan_s{nullptr, in, out, forward};
    long len[2] = {n0, n1};
    long real_strides[3] = {0, n1, 1};
    long cplx_strides[3] = {0, n1 / 2 + 1, 1};
    must(DftiCreateDescriptor_d_md(&p->h, kReal, 2, len), "create(real,2d)");
    must(DftiSetValue(p->h, kPlacement, kNotInplace), "placement");
    must(DftiSetValue(p->h, kConjEvenStorage, kComplexComplex), "cce storage");
    must(DftiSetValue(p->h, kInStrides, forward ? real_strides : cplx_strides), "in strides");
    must(DftiSetValue(p->h, kOutStrides, forward ? cplx_strides : real_strides), "out strides");
    must(DftiCommitDescriptor(p->h), "commit");
    return p;
}

extern "C" fftw_plan fftw_plan_dft_r2c_2d(int n0, int n1, double *in, fftw_complex *out, unsigned) {
    return real_2d(n0, n1, in, out, true);
}

extern "C" fftw_plan fftw_plan_dft_c2r_2d(int n0, int n1, fftw_complex *in, double *out, unsigned) {
    return real_2d(n0, n1, in, out, false);
}

extern "C" fftw_plan fftw_plan_dft_1d(int n, fftw_complex *in, fftw_complex *out, int sign, unsigned) {
    auto *p = new oracle_fftw_plan_s{nullptr, in, out, sign == FFTW_FORWARD};
    must(DftiCreateDescriptor_d_1d(&p->h, kComplex, n), "create(complex,1d)");
    must(DftiSetValue(p->h, kPlacement, in == out ? kInplace : kNotInplace), "placement");
    must(DftiCommitDescriptor(p->h), "commit");
    return p;
}

extern "C" void fftw_execute(const fftw_plan p) {
    if (p->in == p->out) {
        must(p->forward ? DftiComputeForward(p->h, p->in) : DftiComputeBackward(p->h, p->in), "compute");
    } else {
        must(p->forward ? DftiComputeForward(p->h, p->in, p->out) : DftiComputeBackward(p->h, p->in, p->out),
             "compute");
    }
}

extern "C" void fftw_destroy_plan(fftw_plan p) {
    if (!p) return;
    DftiFreeDescriptor(&p->h);
    delete p;
}
