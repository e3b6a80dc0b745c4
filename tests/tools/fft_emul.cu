// Host emulation of the shared-memory FFTs (fitsne_fft.cuh, fitsne_conv.cuh), CPU only, test infrastructure.
//  (1) Stockham (natural order; row passes, 1-D lines): for every FFT length the grid ladder can produce, run all
//      "threads" of a CTA through stage 0, then stage 1, ... -- what the barrier between stages enforces -- on interleaved
//      sequences in the padded / skewed shared-memory layout, and compare with a direct fp64 DFT.
//  (2) In-place column transform (2-D lengths): the [pos][4 slots] tile, forward decimation-in-frequency stages with the
//      zero-substituting first stage, checked against a direct DFT THROUGH THE DIGIT-REVERSAL the plan implies, then the
//      conjugate trick + inverse decimation-in-time stages, which must give back M x the input in natural order.
// Prints FFT_EMUL_OK.
#include "../../fit-sne_b200/csrc/fitsne_kernels.cuh"
#include "../../fit-sne_b200/csrc/fitsne_fft.cuh"
#include "../../fit-sne_b200/csrc/fitsne_conv.cuh"
#include <cmath>
#include <complex>
#include <cstdio>
#include <random>
#include <set>
#include <vector>
using namespace fk;

static double run_len(int M, int lines, bool *plan_ok, int *nstages) {
    FftPlan plan;
    *plan_ok = fft_make_plan(M, &plan);
    if (!*plan_ok) return 0;
    *nstages = plan.nstages;
    const int NS = fft_buf_len(M, lines);
    std::vector<float2> bufa((size_t) lines * NS, make_float2(NAN, NAN)), bufb(bufa), W(M);
    for (int k = 0; k < M; k++) { const double a = -2.0 * M_PI * (double) k / (double) M; W[k] = make_float2((float) cos(a), (float) sin(a)); }
    std::mt19937 rng(M * 7 + lines);
    std::uniform_real_distribution<float> U(-1.f, 1.f);
    std::vector<std::complex<double>> in((size_t) lines * M);
    for (int l = 0; l < lines; l++) for (int i = 0; i < M; i++) {
        const float2 v = make_float2(U(rng), U(rng));
        in[(size_t) l * M + i] = {v.x, v.y};
        bufa[(size_t) l * NS + fft_phys(i)] = v;
    }
    float2 *x = bufa.data(), *y = bufb.data();
    int n_cur = M, s = 1;
    const int nthreads = lines == 1 ? 512 : 256;      // k_fft_line / the row kernels of fitsne_conv.cuh
    for (int st = 0; st < plan.nstages; st++) {
        for (int tid = 0; tid < nthreads; tid++) fft_run_stage(x, y, NS, lines, plan, st, n_cur, s, W.data(), tid, nthreads);
        n_cur /= plan.radix[st]; s *= plan.radix[st];
        std::swap(x, y);
    }
    // direct DFT of line 0 and of the last line (fp64), a sample of 64 output bins each for long lengths
    double max_err = 0, max_ref = 0;
    for (int l : {0, lines - 1}) {
        const int step = M > 1024 ? M / 61 : 1;
        for (int k = 0; k < M; k += step) {
            std::complex<double> acc = 0;
            for (int i = 0; i < M; i++) {
                const double a = -2.0 * M_PI * (double) (((long long) i * k) % M) / (double) M;
                acc += in[(size_t) l * M + i] * std::complex<double>(cos(a), sin(a));
            }
            const float2 got = x[(size_t) l * NS + fft_phys(k)];
            max_err = std::max(max_err, std::abs(acc - std::complex<double>(got.x, got.y)));
            max_ref = std::max(max_ref, std::abs(acc));
        }
    }
    return max_err / max_ref;
}

// in-place column transform: returns the worse of (forward vs DFT, round trip vs M * input), relative
static double run_col(int M, int nz, bool *plan_ok, int inv_slots = COL_SLOTS, bool wide = false) {
    ColPlan pl;
    *plan_ok = col_make_plan(M, &pl, wide);
    if (!*plan_ok) return 0;
    std::vector<float2> W(M), x((size_t) M * COL_SLOTS, make_float2(NAN, NAN));
    for (int k = 0; k < M; k++) { const double a = -2.0 * M_PI * (double) k / (double) M; W[k] = make_float2((float) cos(a), (float) sin(a)); }
    std::mt19937 rng(M * 13 + nz);
    std::uniform_real_distribution<float> U(-1.f, 1.f);
    std::vector<std::complex<double>> in((size_t) COL_SLOTS * M, 0.0);
    for (int sl = 0; sl < COL_SLOTS; sl++) for (int i = 0; i < nz; i++) {      // positions >= nz stay poisoned: they must not be read
        const float2 v = make_float2(U(rng), U(rng));
        in[(size_t) sl * M + i] = {v.x, v.y};
        x[(size_t) i * COL_SLOTS + sl] = v;
    }
    const int nthreads = COL_THREADS;
    for (int tid = 0; tid < nthreads; tid++) col_run_fwd_stage<true>(x.data(), pl, 0, nz, W.data(), tid, nthreads);
    for (int st = 1; st < pl.nstages; st++)
        for (int tid = 0; tid < nthreads; tid++) col_run_fwd_stage<false>(x.data(), pl, st, nz, W.data(), tid, nthreads);
    // frequency held by position pos: digit d_st = (pos / m_st) % radix_st contributes d_st * tws_st
    auto freq = [&](int pos) { int k = 0; for (int st = 0; st < pl.nstages; st++) k += ((pos / pl.m[st]) % pl.radix[st]) * pl.tws[st]; return k; };
    double max_err = 0, max_ref = 0;
    std::vector<char> seen(M, 0);
    for (int pos = 0; pos < M; pos++) { const int k = freq(pos); if (k < 0 || k >= M || seen[k]) return 1.0; seen[k] = 1; }   // a permutation
    const int step = M > 512 ? M / 53 : 1;
    for (int sl : {0, COL_SLOTS - 1}) for (int pos = 0; pos < M; pos += step) {
        const int k = freq(pos);
        std::complex<double> acc = 0;
        for (int i = 0; i < nz; i++) {
            const double a = -2.0 * M_PI * (double) (((long long) i * k) % M) / (double) M;
            acc += in[(size_t) sl * M + i] * std::complex<double>(cos(a), sin(a));
        }
        const float2 got = x[(size_t) pos * COL_SLOTS + sl];
        max_err = std::max(max_err, std::abs(acc - std::complex<double>(got.x, got.y)));
        max_ref = std::max(max_ref, std::abs(acc));
    }
    double e1 = max_err / max_ref;
    // conjugate (what the Hadamard step does while storing), inverse stages, compare with M * input
    for (auto &v : x) v.y = -v.y;
    for (int st = pl.nstages - 1; st > 0; st--)
        for (int tid = 0; tid < nthreads; tid++) col_run_inv_stage<false>(x.data(), pl, st, W.data(), tid, nthreads, inv_slots);
    for (int tid = 0; tid < nthreads; tid++) col_run_inv_stage<true>(x.data(), pl, 0, W.data(), tid, nthreads, inv_slots);
    max_err = 0; max_ref = 0;
    for (int sl = 0; sl < inv_slots; sl++) for (int i = 0; i < M; i++) {
        const std::complex<double> want = in[(size_t) sl * M + i] * (double) M;
        const float2 got = x[(size_t) i * COL_SLOTS + sl];
        max_err = std::max(max_err, std::abs(want - std::complex<double>(got.x, got.y)));
        max_ref = std::max(max_ref, std::abs(want));
    }
    return std::max(e1, max_err / max_ref);
}

int main() {
    std::set<int> lens;
    for (int n = 32; n <= 8192; n += 2) { const int m = nice_fft_size(n); if (m <= 8192) lens.insert(m); }
    bool ok = true;
    int count = 0, ccount = 0;
    double worst = 0, worst_col = 0;
    for (int M : lens) {
        {   // plan families: both factor M exactly, the wide one never needs more stages, and an odd radix comes last when there is one
            int rn[FFT_MAX_STAGES], rw[FFT_MAX_STAGES];
            const int sn = fft_pick_radices(M, false, rn), sw = fft_pick_radices(M, true, rw);
            long long pn = 1, pw = 1;
            for (int i = 0; i < sn; i++) pn *= rn[i];
            for (int i = 0; i < sw; i++) pw *= rw[i];
            bool odd_in = false;
            for (int i = 0; i < sw; i++) odd_in = odd_in || (rw[i] & 1);
            if (sn == 0 || sw == 0 || pn != M || pw != M || sw > sn || (odd_in && !(rw[sw - 1] & 1))) {
                printf("M=%d: radix plans inconsistent (narrow %d stages, wide %d stages)\n", M, sn, sw); ok = false;
            }
        }
        for (int lines : {1, 2}) {
            if (lines > 1 && M > 4096) continue;
            bool pn; int sn = 0;
            const double en = run_len(M, lines, &pn, &sn);
            if (!pn || !(en < 3e-6)) { printf("M=%d lines=%d: %d stages err %.2e  FAILED\n", M, lines, sn, en); ok = false; }
            worst = std::max(worst, en);
            if (lines == 1) count++;
        }
        if (M <= 4096) {
            for (int nz : {M / 2, M / 2 - M / 7, M}) {
                bool pc;
                bool pw1, pw2;
                const double ec = std::max(std::max(run_col(M, nz, &pc), run_col(M, nz, &pc, 3)),
                                           std::max(run_col(M, nz, &pw1, COL_SLOTS, true), run_col(M, nz, &pw2, 3, true)));   // both plan families
                pc = pc && pw1 && pw2;
                if (!pc || !(ec < 4e-6)) { printf("M=%d nz=%d: in-place column transform err %.2e  FAILED\n", M, nz, ec); ok = false; }
                worst_col = std::max(worst_col, ec);
            }
            ccount++;
        }
    }
    printf("%d lengths (32..8192) Stockham, worst rel. error %.2e; %d lengths in-place column transform, worst %.2e\n", count, worst, ccount, worst_col);
    if (ok) printf("FFT_EMUL_OK\n");
    return ok ? 0 : 1;
}
