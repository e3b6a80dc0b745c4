// fitsne_fft.cuh -- hand-written shared-memory FFT building blocks for the circulant kernel convolution
// (nbodyfft.cpp:150-217, :401-433).
//
// Why not cuFFT: on sm_100 cuFFT finalises (JIT-compiles) kernels at plan-creation time -- 1-3 s per new FFT
// length on a machine that has not seen it -- and the grid size of a t-SNE run drifts through a dozen lengths,
// so plan creation cost more wall-clock than the whole optimisation.  These kernels
// need no plans and work for every length 2^a 3^b 5^c <= 4096 (8192 in 1-D).
// Algorithm: Stockham autosort, decimation in frequency, mixed radix 8/4/2/3/5, ping-pong in shared memory, one
// __syncthreads per stage, natural order in and out; twiddles from a per-length fp32 table computed in fp64.  The
// inverse transform is conj(FFT(conj(x))); normalisation is folded into the kernel samples.
// Users: the row passes of the 2-D convolution (fitsne_conv.cuh; its column pass has its own in-place transform) and the
// 1-D convolution (k_fft_line below).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fk {

constexpr int FFT_MAX_STAGES = 14;

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

// radices of a length-n transform.  wide == false: 8, 4, 2, 3, 5.  wide == true: as few stages as radices up to 16 allow --
// 2^a as 16s plus one of (8 | 4 | 2 | 8 4), 3^b as 9s plus a 3, then the 5s (1152 = 16 8 9, 1280 = 16 16 5).  Even radices
// first, odd ones last.  Returns the stage count, 0 if n is not of the form 2^a 3^b 5^c.
__host__ inline int fft_pick_radices(int n, bool wide, int *radix) {
    int ns = 0, m = n;
    auto push = [&](int r) { if (ns < FFT_MAX_STAGES) radix[ns++] = r; };
    auto take = [&](int r) { while (m % r == 0 && ns < FFT_MAX_STAGES) { radix[ns++] = r; m /= r; } };
    if (wide) {
        int a = 0;
        while (m % 2 == 0) { m /= 2; a++; }
        const int k = a / 4, r = a % 4;
        if (r == 1 && k >= 1) { for (int i = 0; i < k - 1; i++) push(16); push(8); push(4); }
        else { for (int i = 0; i < k; i++) push(16); if (r == 1) push(2); else if (r == 2) push(4); else if (r == 3) push(8); }
        take(9); take(3); take(5);
    } else {
        take(8); take(4); take(2); take(3); take(5);
    }
    return m == 1 ? ns : 0;
}

// The Stockham transforms (rows, 1-D lines) keep the narrow radices.  Measured on B200 with radix-16 / radix-9 stages in
// fft_stage (1152 = 16 8 9): the row kernels grew from 64-80 to 127 registers, the convolution phase went 0.0925 -> 0.0908 ms
// and the iteration did not move (2 597 -> 2 589 it/s), while every length that keeps narrow radices paid for the lower
// occupancy (2 597 -> 2 510 it/s) -- unlike the in-place column transform, where a stage is a pass over the whole tile.
__host__ inline bool fft_make_plan(int n, FftPlan *p) {
    p->n = n;
    p->nstages = fft_pick_radices(n, false, p->radix);
    if (p->nstages == 0) return false;
    int sacc = 1;
    for (int st = 0; st < p->nstages; st++) {
        p->div_s[st] = make_fastdiv((uint32_t) sacc);
        sacc *= p->radix[st];
        // butterfly slots per line group: the power of two >= n/8, 32..256
        const int per = n / 8;
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

// a * (c - i s): multiplication by the constant twiddle exp(-i phi), c = cos phi, s = sin phi
__host__ __device__ __forceinline__ float2 cmul_cs(float2 a, float c, float s) { return make_float2(a.x * c + a.y * s, a.y * c - a.x * s); }

// Radix 16 = 4 x 4 and radix 9 = 3 x 3 (Cooley-Tukey inside the registers: sub-DFTs over n1 with stride R2, constant
// twiddles w_R^(n2 k1), sub-DFTs over n2; X[k1 + R1 k2]).  Used by the in-place column transform of fitsne_conv.cuh, where
// a stage is a full pass over the shared-memory tile: 1152 = 16 * 8 * 9 takes three passes instead of five (8 8 2 3 3).
template <>
__host__ __device__ __forceinline__ void dft_small<16>(float2 (&a)[16]) {
    const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, h = 0.70710678118654752440f;
    float2 y[4][4];                          // y[n2][k1]
#pragma unroll
    for (int n2 = 0; n2 < 4; n2++) {
        float2 t[4] = {a[n2], a[4 + n2], a[8 + n2], a[12 + n2]};
        dft_small<4>(t);
#pragma unroll
        for (int k1 = 0; k1 < 4; k1++) y[n2][k1] = t[k1];
    }
    // w16^(n2 k1)
    y[1][1] = cmul_cs(y[1][1], c1, s1);                                          // w^1
    y[1][2] = make_float2(h * (y[1][2].x + y[1][2].y), h * (y[1][2].y - y[1][2].x));   // w^2 = (1 - i)/sqrt2
    y[1][3] = cmul_cs(y[1][3], s1, c1);                                          // w^3
    y[2][1] = make_float2(h * (y[2][1].x + y[2][1].y), h * (y[2][1].y - y[2][1].x));   // w^2
    y[2][2] = mul_mi(y[2][2]);                                                   // w^4 = -i
    y[2][3] = make_float2(h * (y[2][3].y - y[2][3].x), -h * (y[2][3].x + y[2][3].y));  // w^6 = (-1 - i)/sqrt2
    y[3][1] = cmul_cs(y[3][1], s1, c1);                                          // w^3
    y[3][2] = make_float2(h * (y[3][2].y - y[3][2].x), -h * (y[3][2].x + y[3][2].y));  // w^6
    y[3][3] = cmul_cs(y[3][3], -c1, -s1);                                        // w^9 = -cos(pi/8) + i sin(pi/8)
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
    const float c1 = 0.76604444311897803520f, s1 = 0.64278760968653932632f;     // 40 degrees
    const float c2 = 0.17364817766693034885f, s2 = 0.98480775301220805937f;     // 80 degrees
    const float c4 = -0.93969262078590838405f, s4 = 0.34202014332566873304f;    // 160 degrees
    float2 y[3][3];                          // y[n2][k1]
#pragma unroll
    for (int n2 = 0; n2 < 3; n2++) {
        float2 t[3] = {a[n2], a[3 + n2], a[6 + n2]};
        dft_small<3>(t);
#pragma unroll
        for (int k1 = 0; k1 < 3; k1++) y[n2][k1] = t[k1];
    }
    y[1][1] = cmul_cs(y[1][1], c1, s1);      // w9^1
    y[1][2] = cmul_cs(y[1][2], c2, s2);      // w9^2
    y[2][1] = cmul_cs(y[2][1], c2, s2);      // w9^2
    y[2][2] = cmul_cs(y[2][2], c4, s4);      // w9^4
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
__host__ __device__ __forceinline__ void fft_run_stage(const float2 *x, float2 *y, int NS, int batch, const FftPlan &plan, int st,
                                                       int n_cur, int s, const float2 *__restrict__ W, int tid, int nthreads) {
    const int N = plan.n;
    const int r = plan.radix[st];
    const FastDiv ds = plan.div_s[st];
    const int tl = plan.tpl_log2[st];
    if (r == 8) fft_stage<8>(x, y, N, NS, batch, n_cur, s, ds, tl, W, tid, nthreads);
    else if (r == 4) fft_stage<4>(x, y, N, NS, batch, n_cur, s, ds, tl, W, tid, nthreads);
    else if (r == 2) fft_stage<2>(x, y, N, NS, batch, n_cur, s, ds, tl, W, tid, nthreads);
    else if (r == 3) fft_stage<3>(x, y, N, NS, batch, n_cur, s, ds, tl, W, tid, nthreads);
    else fft_stage<5>(x, y, N, NS, batch, n_cur, s, ds, tl, W, tid, nthreads);
}

#ifdef __CUDACC__
// Runs all stages (ping-pong between a and b); returns the buffer holding the result.  W may live in shared or global
// memory.  Every stage ends with a barrier.
__device__ __forceinline__ float2 *fft_smem(float2 *a, float2 *b, int NS, int batch, const FftPlan &plan, const float2 *__restrict__ W) {
    int n_cur = plan.n, s = 1;
    float2 *x = a, *y = b;
    for (int st = 0; st < plan.nstages; st++) {
        fft_run_stage(x, y, NS, batch, plan, st, n_cur, s, W, (int) threadIdx.x, (int) blockDim.x);
        __syncthreads();
        n_cur /= plan.radix[st]; s *= plan.radix[st];
        float2 *t = x; x = y; y = t;
    }
    return x;
}

// 1-D embeddings: in-place FFT of one length-M line per CTA (grid = number of lines, line l at data + l*M).
//   zero_from != nullptr (forward passes of the charge lines): elements >= *zero_from are taken as zero without being read.
// Dynamic smem: (2 * fft_buf_len(M, 1) + M) float2.
constexpr int FFT_THREADS = 512;
__global__ void __launch_bounds__(FFT_THREADS) k_fft_line(float2 *__restrict__ data, FftPlan plan, const float2 *__restrict__ W, int inverse,
                                                          unsigned zero_mask, const int *__restrict__ zero_from, const int *__restrict__ ok) {
    pdl_prologue();
    if (ok && !*ok) return;
    extern __shared__ float2 fft_sm[];
    __shared__ FftPlan plan_s;
    for (int i = threadIdx.x; i < (int) (sizeof(FftPlan) / 4); i += blockDim.x)
        reinterpret_cast<int *>(&plan_s)[i] = reinterpret_cast<const int *>(&plan)[i];
    const int M = plan.n, NS = fft_buf_len(M, 1);
    const int gz = ((zero_mask >> blockIdx.x) & 1u) ? *zero_from : M;
    float2 *bufa = fft_sm, *bufb = fft_sm + NS, *Ws = fft_sm + 2 * NS;
    for (int i = threadIdx.x; i < M; i += blockDim.x) Ws[i] = W[i];
    float2 *line = data + (size_t) blockIdx.x * M;
    for (int pos = threadIdx.x; pos < M; pos += blockDim.x) {
        float2 v = pos < gz ? line[pos] : make_float2(0.f, 0.f);
        if (inverse) v.y = -v.y;
        bufa[fft_phys(pos)] = v;
    }
    __syncthreads();
    const float2 *res = fft_smem(bufa, bufb, NS, 1, plan_s, Ws);
    for (int pos = threadIdx.x; pos < M; pos += blockDim.x) {
        float2 v = res[fft_phys(pos)];
        if (inverse) v.y = -v.y;
        line[pos] = v;
    }
}
#endif  // __CUDACC__

}  // namespace fk
