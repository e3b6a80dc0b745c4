// fitsne_fft.cuh -- hand-written shared-memory FFTs for the circulant kernel convolution (nbodyfft.cpp:150-217).
//
// Why not cuFFT: on sm_100 cuFFT finalises (JIT-compiles) kernels at plan-creation time -- 1-3 s per new FFT
// length on a machine that has not seen it -- and the grid size of a t-SNE run drifts through a dozen lengths,
// so plan creation cost more wall-clock than the whole optimisation.  These kernels
// need no plans, work for every length 2^a 3^b 5^c <= 4096 (8192 in 1-D), and let the convolution use:
//   * two real planes per complex transform (w1 + i*delta_x, ...), separated inside the Hadamard kernel;
//   * pruning: only the G non-zero rows of the zero-padded input are row-transformed, and only the G rows of the
//     output that the gather reads are inverse row-transformed;
//   * in-place passes: rows (contiguous) then columns (tiles of FFT_TC adjacent columns, 32-byte segments).
// Algorithm: Stockham autosort, decimation in frequency, mixed radix 8/4/2/3/5, ping-pong in shared memory, one
// __syncthreads per stage; twiddles from a per-length fp32 table computed in fp64.  The inverse transform is
// conj(FFT(conj(x))) (conjugation folded into the global loads/stores); normalisation is folded into the kernel
// samples (k_gen_kernels), as before.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fk {

constexpr int FFT_MAX_STAGES = 14;
constexpr int FFT_THREADS = 512;

// shared-memory index skew: one pad slot every 8 elements turns the stride-R / stride-8R writes of the first Stockham
// stages (which would hit 2 of the 16 eight-byte bank pairs) into conflict-free or 2-way patterns
__host__ __device__ __forceinline__ int fft_phys(int i) { return i + (i >> 3); }
// per-sequence buffer stride: skewed length rounded up so that stride % 16 == 16 / lines -- the `lines` sequences of a
// column tile then start in different 8-byte bank pairs and the transposing stage-in/out is conflict-free
__host__ __device__ __forceinline__ int fft_buf_len(int n, int lines) {
    int len = n + (n >> 3) + 1;
    const int want = lines >= 16 ? 1 : 16 / (lines < 1 ? 1 : lines);
    while ((len & 15) != (want & 15)) len++;
    return len;
}

// exact division of small non-negative integers (n < 2^20) by an invariant: shifts for powers of two, multiply-high else
struct FastDiv {
    uint32_t mul, shift;   // mul == 0: power of two, q = n >> shift;  else q = umulhi(n, mul) >> shift
};
__host__ inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    uint32_t l = 0;
    while ((1u << l) < d) l++;
    if ((1u << l) == d) { f.mul = 0; f.shift = l; return f; }
    // round-up method with a 32-bit multiplier: mul = floor(2^(31+l) / d) + 1, q = umulhi(n, mul) >> (l-1);
    // exact for n < 2^20 and d < 2^15 (checked exhaustively offline for the ranges used here)
    f.mul = (uint32_t) ((((uint64_t) 1 << (32 + l - 1)) / d) + 1);
    f.shift = l - 1;
    return f;
}
__device__ __forceinline__ int fastdiv(int n, FastDiv f) {
    return f.mul ? (int) (__umulhi((uint32_t) n, f.mul) >> f.shift) : (n >> f.shift);
}

struct FftPlan {
    int n;
    int nstages;
    int radix[FFT_MAX_STAGES];
    FastDiv div_s[FFT_MAX_STAGES];     // by s = product of the radices of the earlier stages
    int tpl_log2[FFT_MAX_STAGES];      // per stage: log2(butterfly slots per line group), see fft_stage
};

// wide == false: radices 8/4/2/3/5 (the round-1 plan).  wide == true: also 16 and 9, which turns the 4-5 stages of the
// lengths a t-SNE grid produces (1152 = 8*8*2*3*3, 1280 = 8*8*4*5, 1024 = 8*8*8*2) into 3 (16*8*9, 16*16*5, 16*16*4):
// every stage is a full pass over shared memory plus a barrier, so fewer, wider stages is less of everything but FMAs.
__host__ inline bool fft_make_plan(int n, FftPlan *p, bool wide = false) {
    p->n = n; p->nstages = 0;
    int m = n;
    auto take = [&](int r) { while (m % r == 0 && p->nstages < FFT_MAX_STAGES) { p->radix[p->nstages++] = r; m /= r; } };
    if (wide) {
        take(16);
        // leftover power of two next (8, 4 or 2), then nines, threes, fives
        take(8); take(4); take(2); take(9); take(3); take(5);
    } else {
        take(8); take(4); take(2); take(3); take(5);
    }
    if (m != 1) return false;
    int sacc = 1;
    for (int st = 0; st < p->nstages; st++) {
        p->div_s[st] = make_fastdiv((uint32_t) sacc);
        sacc *= p->radix[st];
        // butterfly slots per line group: the power of two >= the stage's butterflies per line (n / radix), 32..256
        // (narrow plans keep the round-1 value: the power of two >= n/8 for every stage)
        const int per = wide ? n / p->radix[st] : n / 8;
        int l2 = 5;
        while ((1 << l2) < per && l2 < 8) l2++;
        p->tpl_log2[st] = l2;
    }
    return true;
}

__global__ void k_fft_twiddles(float2 *__restrict__ W, int n) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double s, c;
    sincospi(-2.0 * (double) k / (double) n, &s, &c);
    W[k] = make_float2((float) c, (float) s);
}

__host__ __device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ float2 cmulf(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__host__ __device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)

// forward DFT of R points in registers, natural output order
template <int R>
__host__ __device__ __forceinline__ void dft_small(float2 (&a)[R]);

template <>
__host__ __device__ __forceinline__ void dft_small<2>(float2 (&a)[2]) {
    const float2 t = a[1];
    a[1] = csub(a[0], t); a[0] = cadd(a[0], t);
}
template <>
__host__ __device__ __forceinline__ void dft_small<4>(float2 (&a)[4]) {
    const float2 t0 = cadd(a[0], a[2]), t1 = csub(a[0], a[2]), t2 = cadd(a[1], a[3]), t3 = mul_mi(csub(a[1], a[3]));
    a[0] = cadd(t0, t2); a[1] = cadd(t1, t3); a[2] = csub(t0, t2); a[3] = csub(t1, t3);
}
template <>
__host__ __device__ __forceinline__ void dft_small<8>(float2 (&a)[8]) {
    // one radix-2 DIF split, then two 4-point DFTs: X[2j] = DFT4(a_k + a_{k+4})[j], X[2j+1] = DFT4((a_k - a_{k+4}) w8^k)[j]
    const float h = 0.70710678118654752440f;
    float2 e[4], o[4];
#pragma unroll
    for (int k = 0; k < 4; k++) { e[k] = cadd(a[k], a[k + 4]); o[k] = csub(a[k], a[k + 4]); }
    o[1] = make_float2(h * (o[1].x + o[1].y), h * (o[1].y - o[1].x));     // * (1 - i)/sqrt2
    o[2] = mul_mi(o[2]);                                                    // * (-i)
    o[3] = make_float2(h * (o[3].y - o[3].x), -h * (o[3].x + o[3].y));    // * (-1 - i)/sqrt2
    dft_small<4>(e);
    dft_small<4>(o);
#pragma unroll
    for (int j = 0; j < 4; j++) { a[2 * j] = e[j]; a[2 * j + 1] = o[j]; }
}
template <>
__host__ __device__ __forceinline__ void dft_small<3>(float2 (&a)[3]) {
    const float2 u = cadd(a[1], a[2]), v = csub(a[1], a[2]);
    const float2 c = make_float2(a[0].x - 0.5f * u.x, a[0].y - 0.5f * u.y);
    const float h = 0.86602540378443864676f;
    const float2 d = make_float2(h * v.y, -h * v.x);           // (-i) * (sqrt(3)/2) * v
    a[0] = cadd(a[0], u); a[1] = cadd(c, d); a[2] = csub(c, d);
}
template <>
__host__ __device__ __forceinline__ void dft_small<5>(float2 (&a)[5]) {
    const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
    const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
    const float2 u1 = cadd(a[1], a[4]), u2 = cadd(a[2], a[3]), v1 = csub(a[1], a[4]), v2 = csub(a[2], a[3]);
    const float2 p1 = make_float2(a[0].x + c1 * u1.x + c2 * u2.x, a[0].y + c1 * u1.y + c2 * u2.y);
    const float2 p2 = make_float2(a[0].x + c2 * u1.x + c1 * u2.x, a[0].y + c2 * u1.y + c1 * u2.y);
    const float2 q1 = mul_mi(make_float2(s1 * v1.x + s2 * v2.x, s1 * v1.y + s2 * v2.y));   // -i * q1
    const float2 q2 = mul_mi(make_float2(s2 * v1.x - s1 * v2.x, s2 * v1.y - s1 * v2.y));   // -i * q2
    a[0] = make_float2(a[0].x + u1.x + u2.x, a[0].y + u1.y + u2.y);
    a[1] = cadd(p1, q1); a[4] = csub(p1, q1); a[2] = cadd(p2, q2); a[3] = csub(p2, q2);
}

// Wide radices as one Cooley-Tukey step in registers: N = N1*N2, input n = N2*n1 + n2, output k = k1 + N1*k2,
//   X[k1 + N1*k2] = sum_n2 w_N^(n2*k1) * (sum_n1 x[N2*n1 + n2] w_N1^(n1*k1)) * w_N2^(n2*k2).
// The inner twiddles w_N^(n2*k1) are literals (multiplying by a literal complex number).
__host__ __device__ __forceinline__ float2 cmulc(float2 a, float c, float s) {      // a * (c - i*s), i.e. times exp(-i*theta)
    return make_float2(a.x * c + a.y * s, a.y * c - a.x * s);
}
template <>
__host__ __device__ __forceinline__ void dft_small<16>(float2 (&a)[16]) {
    // 16 = 4 x 4.  cos/sin of k*pi/8
    const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, h = 0.70710678118654752440f;
    float2 y[4][4];
#pragma unroll
    for (int n2 = 0; n2 < 4; n2++) {
        float2 t[4] = {a[n2], a[4 + n2], a[8 + n2], a[12 + n2]};
        dft_small<4>(t);
#pragma unroll
        for (int k1 = 0; k1 < 4; k1++) y[n2][k1] = t[k1];
    }
    // twiddles w16^(n2*k1): exponents 1,2,3 / 2,4,6 / 3,6,9
    y[1][1] = cmulc(y[1][1], c1, s1); y[1][2] = cmulc(y[1][2], h, h);     y[1][3] = cmulc(y[1][3], s1, c1);
    y[2][1] = cmulc(y[2][1], h, h);   y[2][2] = mul_mi(y[2][2]);          y[2][3] = cmulc(y[2][3], -h, h);
    y[3][1] = cmulc(y[3][1], s1, c1); y[3][2] = cmulc(y[3][2], -h, h);    y[3][3] = cmulc(y[3][3], -c1, -s1);
#pragma unroll
    for (int k1 = 0; k1 < 4; k1++) {
        float2 t[4] = {y[0][k1], y[1][k1], y[2][k1], y[3][k1]};
        dft_small<4>(t);
#pragma unroll
        for (int k2 = 0; k2 < 4; k2++) a[k1 + 4 * k2] = t[k2];
    }
}
template <>
__host__ __device__ __forceinline__ void dft_small<9>(float2 (&a)[9]) {
    // 9 = 3 x 3.  cos/sin of 40, 80 and 160 degrees
    const float c1 = 0.76604444311897803520f, s1 = 0.64278760968653932632f;
    const float c2 = 0.17364817766693034885f, s2 = 0.98480775301220805937f;
    const float c4 = -0.93969262078590838405f, s4 = 0.34202014332566873304f;
    float2 y[3][3];
#pragma unroll
    for (int n2 = 0; n2 < 3; n2++) {
        float2 t[3] = {a[n2], a[3 + n2], a[6 + n2]};
        dft_small<3>(t);
#pragma unroll
        for (int k1 = 0; k1 < 3; k1++) y[n2][k1] = t[k1];
    }
    y[1][1] = cmulc(y[1][1], c1, s1); y[1][2] = cmulc(y[1][2], c2, s2);
    y[2][1] = cmulc(y[2][1], c2, s2); y[2][2] = cmulc(y[2][2], c4, s4);
#pragma unroll
    for (int k1 = 0; k1 < 3; k1++) {
        float2 t[3] = {y[0][k1], y[1][k1], y[2][k1]};
        dft_small<3>(t);
#pragma unroll
        for (int k2 = 0; k2 < 3; k2++) a[k1 + 3 * k2] = t[k2];
    }
}

// One Stockham stage of radix R for `batch` independent length-N sequences stored at x + b*NS, written to y + b*NS.
// Sub-problem (n_cur, s): butterflies t in [0, N/R): p = t / s, q = t % s, m = n_cur / R,
//   a_k = x[q + s*(p + k*m)],  y[q + s*(R*p + j)] = (sum_k a_k w_R^{jk}) * W_N[p*j*s].   W lives in shared memory.
// Thread layout: tid = (line group lb, butterfly slot ts) with `tpl` slots per group.  A thread computes the (skewed)
// shared-memory offsets and the R-1 twiddles of its butterfly ONCE and re-uses them for every sequence of its line
// group -- the integer address arithmetic, not the floating-point work, is what dominated a one-butterfly-per-
// iteration formulation (ncu: IMAD/ISETP/LEA > 45 % of issued instructions, profiles/).
__host__ __device__ __forceinline__ int fastdiv_hd(int n, FastDiv f) {
#ifdef __CUDA_ARCH__
    return fastdiv(n, f);
#else
    return f.mul ? (int) ((((uint64_t) (uint32_t) n * (uint64_t) f.mul) >> 32) >> f.shift) : (n >> f.shift);
#endif
}

template <int R>
__host__ __device__ __forceinline__ void fft_stage(const float2 *__restrict__ x, float2 *__restrict__ y, int N, int NS, int batch, int n_cur, int s,
                                                   FastDiv div_s, int tpl_log2, const float2 *__restrict__ W, int tid, int nthreads) {
    const int m = n_cur / R;
    const int per = N / R;
    const int sm_ = s * m;
    const int tpl = 1 << tpl_log2;
    const int ts = tid & (tpl - 1), lb = tid >> tpl_log2, lgroups = (nthreads >> tpl_log2) > 1 ? (nthreads >> tpl_log2) : 1;
    for (int t = ts; t < per; t += tpl) {
        const int p = fastdiv_hd(t, div_s), q = t - p * s;
        const int i0 = q + s * p, o0 = q + s * R * p;
        int xi[R], yo[R];
        float2 w[R];
#pragma unroll
        for (int k = 0; k < R; k++) { xi[k] = fft_phys(i0 + k * sm_); yo[k] = fft_phys(o0 + s * k); }
        const int step = p * s;
#pragma unroll
        for (int j = 1; j < R; j++) w[j] = W[step * j];       // W[0] = 1 when p == 0
        for (int b = lb; b < batch; b += lgroups) {
            const float2 *xb = x + b * NS;
            float2 *yb = y + b * NS;
            float2 a[R];
#pragma unroll
            for (int k = 0; k < R; k++) a[k] = xb[xi[k]];
            dft_small<R>(a);
            yb[yo[0]] = a[0];
#pragma unroll
            for (int j = 1; j < R; j++) yb[yo[j]] = cmulf(a[j], w[j]);
        }
    }
}

// One thread's share of stage `st` (x -> y).  Host-callable: tests/tools/fft_emul.cu runs every thread of a CTA through a
// stage, then the next stage -- what the barrier in fft_smem enforces -- and compares with a direct DFT.
template <bool WIDE>
__host__ __device__ __forceinline__ void fft_run_stage(const float2 *x, float2 *y, int NS, int batch, const FftPlan &plan, int st,
                                                       int n_cur, int s, const float2 *__restrict__ W, int tid, int nthreads) {
    const int N = plan.n;
    const int r = plan.radix[st];
    const FastDiv ds = plan.div_s[st];
    const int tl = plan.tpl_log2[st];
    if (r == 8) fft_stage<8>(x, y, N, NS, batch, n_cur, s, ds, tl, W, tid, nthreads);
    else if (WIDE && r == 16) fft_stage<16>(x, y, N, NS, batch, n_cur, s, ds, tl, W, tid, nthreads);
    else if (r == 4) fft_stage<4>(x, y, N, NS, batch, n_cur, s, ds, tl, W, tid, nthreads);
    else if (r == 2) fft_stage<2>(x, y, N, NS, batch, n_cur, s, ds, tl, W, tid, nthreads);
    else if (WIDE && r == 9) fft_stage<9>(x, y, N, NS, batch, n_cur, s, ds, tl, W, tid, nthreads);
    else if (r == 3) fft_stage<3>(x, y, N, NS, batch, n_cur, s, ds, tl, W, tid, nthreads);
    else fft_stage<5>(x, y, N, NS, batch, n_cur, s, ds, tl, W, tid, nthreads);
}

// Runs all stages (ping-pong between a and b); returns the buffer holding the result.
template <bool WIDE>
__device__ __forceinline__ float2 *fft_smem(float2 *a, float2 *b, int NS, int batch, const FftPlan &plan, const float2 *__restrict__ W) {
    int n_cur = plan.n, s = 1;
    float2 *x = a, *y = b;
    for (int st = 0; st < plan.nstages; st++) {
        fft_run_stage<WIDE>(x, y, NS, batch, plan, st, n_cur, s, W, (int) threadIdx.x, (int) blockDim.x);
        __syncthreads();
        n_cur /= plan.radix[st]; s *= plan.radix[st];
        float2 *t = x; x = y; y = t;
    }
    return x;
}

// In-place FFT of `lines` rows (COLS=false: contiguous) or columns (COLS=true: `lines` adjacent columns, row pitch M)
// per CTA: one large CTA per SM, all of its global loads issued before the first use (FFT_EPT independent loads per
// thread, 64-byte column segments), Stockham stages in shared memory, store back.
//   rows: grid = (ceil(rows_total / lines), nplanes); prune_mask bit i set => plane i only needs rows < *g_rows
//         (g_rows points at GridParams::G on the device)
//   cols: grid = (ceil(M / lines), nplanes)
//   forward passes (inverse == 0): prune_mask bit i also says "plane i is a zero-padded G x G corner": everything outside
//   the corner is taken as zero WITHOUT being read (rows pass: columns >= G; columns pass: rows >= G), so the padding never
//   has to be written to memory and whatever an earlier iteration left there is ignored.
// Dynamic smem: (2 * lines * fft_buf_len(M, lines) + M) float2.  Requires M * lines <= FFT_EPT * FFT_THREADS.
constexpr int FFT_EPT = 24;

// WIDE: the plan may contain radix-16 / radix-9 stages (separate instantiation: the wide butterflies need ~2x the
// registers, which must not cost the narrow kernel its occupancy at small M)
template <bool COLS, bool WIDE = false>
__global__ void __launch_bounds__(FFT_THREADS) k_fft_pass(float2 *__restrict__ data, size_t plane, int rows_total, int lines,
                                                             FftPlan plan, const float2 *__restrict__ W, int inverse, unsigned prune_mask,
                                                             const int *__restrict__ g_rows, const int *__restrict__ ok,
                                                             const unsigned *__restrict__ skip_mask) {
    if (ok && !*ok) return;
    if (skip_mask && ((*skip_mask >> blockIdx.y) & 1u)) return;      // plane keeps its (cached) contents this iteration
    extern __shared__ float2 fft_sm[];
    // the stage loop indexes the plan dynamically: keep it in shared memory, not in the (slow to index) parameter bank
    __shared__ FftPlan plan_s;
    for (int i = threadIdx.x; i < (int) (sizeof(FftPlan) / 4); i += blockDim.x)
        reinterpret_cast<int *>(&plan_s)[i] = reinterpret_cast<const int *>(&plan)[i];
    const int M = plan.n, NS = fft_buf_len(M, lines);
    const int l0 = blockIdx.x * lines;
    int limit = COLS ? M : rows_total;
    if (!COLS && ((prune_mask >> blockIdx.y) & 1u)) limit = min(limit, *g_rows);
    if (l0 >= limit) return;
    const int nl = min(lines, limit - l0);
    const bool zpad = !inverse && ((prune_mask >> blockIdx.y) & 1u);
    const int gz = zpad ? *g_rows : M;                                 // data extent along the transformed axis
    float2 *bufa = fft_sm, *bufb = fft_sm + (size_t) lines * NS, *Ws = fft_sm + (size_t) 2 * lines * NS;
    for (int i = threadIdx.x; i < M; i += blockDim.x) Ws[i] = W[i];
    float2 *base = data + (size_t) blockIdx.y * plane + (COLS ? (size_t) l0 : (size_t) l0 * M);
    // staging without integer divisions: rows line by line (contiguous); columns with `lines` a power of two
    // (i -> pos = i >> lg, line = i & (lines-1); full tiles only, M is a multiple of 16)
    int lg = 0;
    while ((1 << lg) < lines) lg++;
    if (COLS) {
        const int total = M << lg;
        float2 v[FFT_EPT];
#pragma unroll
        for (int u = 0; u < FFT_EPT; u++) {
            const int i = threadIdx.x + u * blockDim.x;
            if (i < total) v[u] = (i >> lg) < gz ? base[(size_t) (i >> lg) * M + (i & (lines - 1))] : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < FFT_EPT; u++) {
            const int i = threadIdx.x + u * blockDim.x;
            if (i < total) {
                if (inverse) v[u].y = -v[u].y;
                bufa[(i & (lines - 1)) * NS + fft_phys(i >> lg)] = v[u];
            }
        }
    } else {
        for (int ln = 0; ln < nl; ln++) {
            const float2 *src = base + (size_t) ln * M;
            float2 *dst = bufa + ln * NS;
#pragma unroll 2
            for (int pos = threadIdx.x; pos < M; pos += blockDim.x) {
                float2 v = pos < gz ? src[pos] : make_float2(0.f, 0.f);
                if (inverse) v.y = -v.y;
                dst[fft_phys(pos)] = v;
            }
        }
    }
    __syncthreads();
    const float2 *res = fft_smem<WIDE>(bufa, bufb, NS, nl, plan_s, Ws);
    if (COLS) {
        const int total = M << lg;
#pragma unroll 4
        for (int u = 0; u < FFT_EPT; u++) {
            const int i = threadIdx.x + u * blockDim.x;
            if (i < total) {
                float2 v = res[(i & (lines - 1)) * NS + fft_phys(i >> lg)];
                if (inverse) v.y = -v.y;
                base[(size_t) (i >> lg) * M + (i & (lines - 1))] = v;
            }
        }
    } else {
        for (int ln = 0; ln < nl; ln++) {
            float2 *dstg = base + (size_t) ln * M;
            const float2 *srcs = res + ln * NS;
            for (int pos = threadIdx.x; pos < M; pos += blockDim.x) {
                float2 v = srcs[fft_phys(pos)];
                if (inverse) v.y = -v.y;
                dstg[pos] = v;
            }
        }
    }
}

}  // namespace fk
