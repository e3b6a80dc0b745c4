/*
 * Minimal stand-in for <fftw3.h> -- TEST INFRASTRUCTURE ONLY (oracle build).
 *
 * FFTW is not installed in this image.  The reference's hot path uses exactly
 * five FFTW entry points (/root/reference/src/nbodyfft.cpp:64-68,165-168,181,
 * 194,211-212,301-304,401-404,410,423,432-433).  This header declares just
 * those, with FFTW's documented semantics (unnormalised transforms, r2c output
 * of n0 x (n1/2+1) complex values); fftw_shim_mkl.cpp implements them on top of
 * MKL DFTI as exported by torch's libtorch_cpu.so.
 *
 * Written from the FFTW API documentation; nothing here is taken from the
 * reference tree.
 */
#ifndef ORACLE_MINI_FFTW3_H
#define ORACLE_MINI_FFTW3_H

#ifdef __cplusplus
extern "C" {
#endif

typedef double fftw_complex[2];
typedef struct oracle_fftw_plan_s *fftw_plan;

#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)

fftw_plan fftw_plan_dft_r2c_2d(int n0, int n1, double *in, fftw_complex *out, unsigned flags);
fftw_plan fftw_plan_dft_c2r_2d(int n0, int n1, fftw_complex *in, double *out, unsigned flags);
fftw_plan fftw_plan_dft_1d(int n, fftw_complex *in, fftw_complex *out, int sign, unsigned flags);
void fftw_execute(const fftw_plan p);
void fftw_destroy_plan(fftw_plan p);

#ifdef __cplusplus
}
#endif
#endif
