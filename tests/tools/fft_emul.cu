// Host emulation of the shared-memory Stockham FFT (fitsne_fft.cuh): for every FFT length the grid ladder can produce
// and for both plan kinds (narrow: radices 8/4/2/3/5; wide: + 16 and 9) run all 512 "threads" of a CTA through stage 0,
// then stage 1, ... -- what the barrier between stages enforces -- on `lines` interleaved sequences in the padded /
// skewed shared-memory layout, and compare with a direct fp64 DFT.  Test infrastructure; prints FFT_EMUL_OK.
#include "../../fit-sne_b200/csrc/fitsne_kernels.cuh"
#include "../../fit-sne_b200/csrc/fitsne_fft.cuh"
#include <cmath>
#include <complex>
#include <cstdio>
#include <random>
#include <set>
#include <vector>
using namespace fk;

template <bool WIDE>
static double run_len(int M, int lines, bool *plan_ok, int *nstages) {
    FftPlan plan;
    *plan_ok = fft_make_plan(M, &plan, WIDE);
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
    const int nthreads = FFT_THREADS;
    for (int st = 0; st < plan.nstages; st++) {
        for (int tid = 0; tid < nthreads; tid++) fft_run_stage<WIDE>(x, y, NS, lines, plan, st, n_cur, s, W.data(), tid, nthreads);
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

int main() {
    std::set<int> lens;
    for (int n = 32; n <= 8192; n += 2) { const int m = nice_fft_size(n); if (m <= 8192) lens.insert(m); }
    bool ok = true;
    int count = 0, fewer = 0;
    double worst_n = 0, worst_w = 0;
    for (int M : lens) {
        // lines as get_plans picks them: columns up to 8, rows up to 4 (2-D, M <= 4096); a single line in 1-D
        for (int lines : {1, 4, 8}) {
            if (lines > 1 && M > 4096) continue;
            if ((size_t) M * lines > (size_t) FFT_EPT * FFT_THREADS) continue;
            bool pn, pw; int sn = 0, sw = 0;
            const double en = run_len<false>(M, lines, &pn, &sn), ew = run_len<true>(M, lines, &pw, &sw);
            if (!pn || !pw || !(en < 3e-6) || !(ew < 3e-6)) { printf("M=%d lines=%d: narrow %d stages err %.2e, wide %d stages err %.2e  FAILED\n", M, lines, sn, en, sw, ew); ok = false; }
            worst_n = std::max(worst_n, en); worst_w = std::max(worst_w, ew);
            if (lines == 1) { count++; if (sw < sn) fewer++; }
        }
    }
    FftPlan a, b;
    fft_make_plan(1152, &a, false); fft_make_plan(1152, &b, true);
    printf("%d lengths (32..8192); wide plans have fewer stages for %d of them (1152: %d -> %d); worst rel. error narrow %.2e, wide %.2e\n",
           count, fewer, a.nstages, b.nstages, worst_n, worst_w);
    if (ok) printf("FFT_EMUL_OK\n");
    return ok ? 0 : 1;
}
