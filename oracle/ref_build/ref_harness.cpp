/*
 * C entry points into the UNMODIFIED reference (oracle/_ref/libfitsne_ref.so)
 * -- TEST INFRASTRUCTURE ONLY.  Nothing under fit-sne_b200/ may link this.
 *
 * The reference's per-iteration functions are private members of class TSNE
 * (/root/reference/src/tsne.h:63-93).  The reference sources are compiled as
 * they lie (see ../Makefile); this translation unit re-opens the class with
 * `#define private public` so the tests can call the very same object code the
 * reference's TSNE::run calls at /root/reference/src/tsne.cpp:446-464,555.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>

#define private public
#include "tsne.h"
#undef private

extern "C" {

/* Dispatch exactly as TSNE::run does (tsne.cpp:446-464).  Returns current_sum_Q. */
double ref_fft_gradient(int N, int no_dims, unsigned int *row_P, unsigned int *col_P, double *val_P, double *Y,
                        double *dC, int nterms, double intervals_per_integer, int min_num_intervals,
                        unsigned int nthreads, double df) {
    TSNE t;
    t.current_sum_Q = 0;
    if (no_dims == 1) {
        if (df == 1.0)
            t.computeFftGradientOneD(nullptr, row_P, col_P, val_P, Y, N, no_dims, dC, nterms, intervals_per_integer,
                                     min_num_intervals, nthreads);
        else
            t.computeFftGradientOneDVariableDf(nullptr, row_P, col_P, val_P, Y, N, no_dims, dC, nterms,
                                               intervals_per_integer, min_num_intervals, nthreads, df);
    } else {
        if (df == 1.0)
            t.computeFftGradient(nullptr, row_P, col_P, val_P, Y, N, no_dims, dC, nterms, intervals_per_integer,
                                 min_num_intervals, nthreads);
        else
            t.computeFftGradientVariableDf(nullptr, row_P, col_P, val_P, Y, N, no_dims, dC, nterms,
                                           intervals_per_integer, min_num_intervals, nthreads, df);
    }
    return t.current_sum_Q;
}

/* evaluateErrorFft (tsne.cpp:1329-1355) with a caller-supplied current_sum_Q.
 * Call with nthreads=1: the reference accumulates C from all threads without
 * synchronisation (tsne.cpp:1349). */
double ref_kl_fft(int N, int no_dims, unsigned int *row_P, unsigned int *col_P, double *val_P, double *Y,
                  double sum_Q, unsigned int nthreads, double df) {
    TSNE t;
    t.current_sum_Q = sum_Q;
    return t.evaluateErrorFft(row_P, col_P, val_P, Y, N, no_dims, nthreads, df);
}

/* Exact O(N^2) gradient (tsne.cpp:1232-1282) on a dense P -- second oracle for small N. */
void ref_exact_gradient(int N, int no_dims, double *P_dense, double *Y, double *dC, double df) {
    TSNE t;
    t.computeExactGradient(P_dense, Y, N, no_dims, dC, df);
}

void ref_zero_mean(double *X, int N, int D) {
    TSNE t;
    t.zeroMean(X, N, D);
}

/* Full TSNE::run with P injected through the reference's own load_affinities=1
 * hook (tsne.cpp:236-281), which reads P_row.dat / P_col.dat / P_val.dat from the
 * CWD: the caller names a scratch directory holding those files. */
int ref_run_with_P(const char *p_dir, int N, double *Y, int no_dims, int max_iter, int stop_lying_iter,
                   int mom_switch_iter, double momentum, double final_momentum, double learning_rate,
                   double early_exag_coeff, double *costs, int no_momentum_during_exag, int start_late_exag_iter,
                   double late_exag_coeff, int nterms, double intervals_per_integer, int min_num_intervals,
                   unsigned int nthreads, double df, double max_step_norm) {
    char cwd[4096];
    if (!getcwd(cwd, sizeof cwd)) return -100;
    if (chdir(p_dir) != 0) return -101;
    double *X = (double *) calloc((size_t) N, sizeof(double)); /* N x 1 dummy; only zero-meaned, never used */
    TSNE t;
    /* perplexity = -1 and sigma/K set: the "manual kernel width" branch is never reached with load_affinities=1 */
    int rc = t.run(X, N, 1, Y, no_dims, -1.0, 0.5, 0, true, max_iter, stop_lying_iter, mom_switch_iter, momentum,
                   final_momentum, learning_rate, 1, 1.0, 2, 1, early_exag_coeff, costs,
                   no_momentum_during_exag != 0, start_late_exag_iter, late_exag_coeff, 1, 1, nterms,
                   intervals_per_integer, min_num_intervals, nthreads, 1, 0, nullptr, df, max_step_norm);
    free(X);
    if (chdir(cwd) != 0) return -102;
    return rc;
}

} /* extern "C" */
