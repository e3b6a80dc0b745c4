// fitsne_conv.cuh -- the 2-D circulant kernel convolution (nbodyfft.cpp:150-217 + the kernel spectrum of
// precompute_2d, nbodyfft.cpp:52-68) as five kernels around ONE fused column pass.
//
//   charge side (critical path)                                   kernel side (in line; sharded runs: a side stream beside sort + spread)
//   k_conv_rows_fwd   rows r < G of the spread grid               k_kspec_rows   rows dr in [0, G) of the kernel lattice,
//                     (w1 + i dx), (dy + i wbb) -> x-spectra,                    sampled in fp64 IN the kernel (no sample pass
//                     separated into 4 half-spectra  S[r][kx][4]                 over HBM), x-transform  -> KR[dr][kx]
//   k_conv_cols       per kx <= M/2: TMA tile load, 4 forward     k_kspec_cols   per kx pair: mirrored columns built from KR,
//                     column FFTs in place, Hadamard with the                    y-transform in place -> KS[kx][pos]
//                     kernel spectra + Parseval sum_Q, 3 inverse
//                     column FFTs in place, TMA tile store
//   k_conv_rows_inv   rows r < G: half-spectra -> (v1, Bx, By)
//
// What makes it cheap:
//   * REAL-input structure is used on both axes.  Rows: two real grids per complex transform, separated with the
//     Z[k] +- conj(Z[-k]) identities INSIDE the row (natural order, shared memory), so only kx in [0, M/2] is kept.
//     Columns: one kx owns all four spectra at (ky, kx), so the Hadamard product and the Parseval terms are local to
//     the tile and the forward-columns -> Hadamard -> inverse-columns chain never leaves shared memory.
//   * The column FFT is IN PLACE (decimation in frequency forward, decimation in time back), so a tile is M x 32 bytes
//     (37 KB at M = 1152: four CTAs per SM, 577 tiles = one wave) instead of a ping-pong pair, and it uses radices up to 16
//     where that saves passes over the tile (1152 = 16 x 8 x 9: three stages each way instead of five).  The forward transform
//     leaves the frequencies in digit-reversed order; the Hadamard product is pointwise, the kernel spectra are produced
//     by the very same transform (k_kspec_cols) in the very same order, and the inverse consumes that order -- nobody
//     ever needs the permutation.
//   * Only G of the M rows are non-zero on input and needed on output: the first forward stage substitutes zeros
//     without reading them and the tile moves G rows (rounded to the TMA box) each way.
//   * The kernels are even (Ksq, Kb) or odd in exactly one axis (Kgrad_x, Kgrad_y = lattice offset * Ksq), so their
//     spectra are real / purely imaginary: rows dr < 0 and columns kx > M/2 follow by symmetry and are never computed,
//     Kgrad_y costs no x-transform at all (offset_y * Ksq~), and the Hadamard product reads ONE float4 per frequency.
//
// Shared-memory tile layout of the column kernels: x[pos * 4 + slot] (float2), slot = which of the tile's 4 sequences.
// A warp's 32 lanes cover 8 consecutive positions x 4 slots = 256 contiguous bytes in every stage whose butterfly
// stride is >= 8, and the plan puts the odd radices last so that the stride-1 stage is conflict-free too.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "fitsne_kernels.cuh"
#include "fitsne_fft.cuh"

namespace fk {

constexpr int COL_SLOTS = 4;          // sequences per column tile
constexpr int COL_THREADS = 256;      // threads per column CTA (4 CTAs share an SM).  Measured at M = 1152 (stage tasks 288 | 576 | 512):
                                      // 192, 224 and 256 threads give the same iteration time, 288 (3 CTAs per SM) is 1.6 % slower
constexpr int COL_THREADS_MAX = 512;  // second instantiation's launch bound: distributed convolution (a rank's few columns, one CTA per SM) launches wider CTAs
constexpr int COL_BOX_ROWS = 128;     // rows per TMA box (box = 128 rows x 32 bytes)
constexpr int ROW_THREADS = 256;

// ------------------------------------------------------------------------------------- in-place column FFT --
struct ColPlan {
    int n, nstages;
    int radix[FFT_MAX_STAGES];
    int m[FFT_MAX_STAGES];            // butterfly stride of stage st: n_cur / radix
    int tws[FFT_MAX_STAGES];          // twiddle stride: n / n_cur
    FastDiv div_m[FFT_MAX_STAGES];
};

// Radices: fft_pick_radices (fitsne_fft.cuh) -- narrow (8, 4, 2, 3, 5) or wide (up to 16: as few passes over the tile as
// possible).  Either way the even radices come first and an odd one last: see the layout note above.
__host__ inline bool col_make_plan(int n, ColPlan *p, bool wide = false) {
    p->n = n;
    p->nstages = fft_pick_radices(n, wide, p->radix);
    if (p->nstages == 0) return false;
    int n_cur = n;
    for (int st = 0; st < p->nstages; st++) {
        p->m[st] = n_cur / p->radix[st];
        p->tws[st] = n / n_cur;
        p->div_m[st] = make_fastdiv((uint32_t) p->m[st]);
        n_cur = p->m[st];
    }
    return n_cur == 1;
}

// (The table stays in global memory / L1: a copy in shared memory behind the tile, filled while the TMA load is in flight,
// was measured on B200 and changed nothing -- convolution phase 0.0930 vs 0.0928-0.0949 ms.)
__host__ __device__ __forceinline__ float2 ld_tw(const float2 *__restrict__ W, int i) {
#ifdef __CUDA_ARCH__
    return __ldg(W + i);
#else
    return W[i];
#endif
}

// Forward stage (decimation in frequency), in place:  a_j = x[base + j m],  x[base + j m] <- DFT_R(a)_j * w_ncur^(q j).
// FIRST: positions >= nz are zero and are not read.
template <int R, bool FIRST>
__host__ __device__ __forceinline__ void col_fwd_stage(float2 *x, int M, int m, int tws, FastDiv div_m, int nz,
                                                       const float2 *__restrict__ W, int tid, int nthreads) {
    const int ntask = (M / R) * COL_SLOTS;
    const int n_cur = m * R;
    for (int task = tid; task < ntask; task += nthreads) {
        const int slot = task & (COL_SLOTS - 1), t = task >> 2;
        const int b = FIRST ? 0 : fastdiv_hd(t, div_m), q = t - b * m;
        const int base = b * n_cur + q;
        float2 a[R];
#pragma unroll
        for (int j = 0; j < R; j++) {
            const int pos = base + j * m;
            a[j] = (!FIRST || pos < nz) ? x[pos * COL_SLOTS + slot] : make_float2(0.f, 0.f);
        }
        dft_small<R>(a);
        x[base * COL_SLOTS + slot] = a[0];
        if (m > 1) {
            const int step = q * tws;
#pragma unroll
            for (int j = 1; j < R; j++) x[(base + j * m) * COL_SLOTS + slot] = cmulf(a[j], ld_tw(W, step * j));
        } else {
#pragma unroll
            for (int j = 1; j < R; j++) x[(base + j) * COL_SLOTS + slot] = a[j];
        }
    }
}

// Inverse stage (decimation in time) on CONJUGATED data, in place: conj(inverse butterfly) = forward butterfly of the
// conjugated, twiddled inputs.  CONJ_OUT (last stage executed = plan stage 0): un-conjugate while storing.
// nslots (3 or 4): only the first nslots sequences of the tile are transformed (the convolution's inverse needs three).
template <int R, bool CONJ_OUT>
__host__ __device__ __forceinline__ void col_inv_stage(float2 *x, int M, int m, int tws, FastDiv div_m,
                                                       const float2 *__restrict__ W, int tid, int nthreads, int nslots) {
    const int ntask = (M / R) * nslots;
    const int n_cur = m * R;
    for (int task = tid; task < ntask; task += nthreads) {
        int slot, t;
        if (nslots == COL_SLOTS) { slot = task & (COL_SLOTS - 1); t = task >> 2; }
        else { t = (int) (((unsigned) task * 43691u) >> 17); slot = task - 3 * t; }      // task / 3, exact for task < 2^16
        const int b = fastdiv_hd(t, div_m), q = t - b * m;
        const int base = b * n_cur + q;
        float2 a[R];
        a[0] = x[base * COL_SLOTS + slot];
        if (m > 1) {
            const int step = q * tws;
#pragma unroll
            for (int j = 1; j < R; j++) a[j] = cmulf(x[(base + j * m) * COL_SLOTS + slot], ld_tw(W, step * j));
        } else {
#pragma unroll
            for (int j = 1; j < R; j++) a[j] = x[(base + j) * COL_SLOTS + slot];
        }
        dft_small<R>(a);
#pragma unroll
        for (int j = 0; j < R; j++) {
            if (CONJ_OUT) a[j].y = -a[j].y;
            x[(base + j * m) * COL_SLOTS + slot] = a[j];
        }
    }
}

template <bool FIRST>
__host__ __device__ __forceinline__ void col_run_fwd_stage(float2 *x, const ColPlan &pl, int st, int nz, const float2 *__restrict__ W,
                                                           int tid, int nthreads) {
    const int M = pl.n, m = pl.m[st], tws = pl.tws[st];
    const FastDiv dm = pl.div_m[st];
    switch (pl.radix[st]) {
        case 16: col_fwd_stage<16, FIRST>(x, M, m, tws, dm, nz, W, tid, nthreads); break;
        case 9: col_fwd_stage<9, FIRST>(x, M, m, tws, dm, nz, W, tid, nthreads); break;
        case 8: col_fwd_stage<8, FIRST>(x, M, m, tws, dm, nz, W, tid, nthreads); break;
        case 4: col_fwd_stage<4, FIRST>(x, M, m, tws, dm, nz, W, tid, nthreads); break;
        case 2: col_fwd_stage<2, FIRST>(x, M, m, tws, dm, nz, W, tid, nthreads); break;
        case 3: col_fwd_stage<3, FIRST>(x, M, m, tws, dm, nz, W, tid, nthreads); break;
        default: col_fwd_stage<5, FIRST>(x, M, m, tws, dm, nz, W, tid, nthreads); break;
    }
}
template <bool CONJ_OUT>
__host__ __device__ __forceinline__ void col_run_inv_stage(float2 *x, const ColPlan &pl, int st, const float2 *__restrict__ W,
                                                           int tid, int nthreads, int nslots = COL_SLOTS) {
    const int M = pl.n, m = pl.m[st], tws = pl.tws[st];
    const FastDiv dm = pl.div_m[st];
    switch (pl.radix[st]) {
        case 16: col_inv_stage<16, CONJ_OUT>(x, M, m, tws, dm, W, tid, nthreads, nslots); break;
        case 9: col_inv_stage<9, CONJ_OUT>(x, M, m, tws, dm, W, tid, nthreads, nslots); break;
        case 8: col_inv_stage<8, CONJ_OUT>(x, M, m, tws, dm, W, tid, nthreads, nslots); break;
        case 4: col_inv_stage<4, CONJ_OUT>(x, M, m, tws, dm, W, tid, nthreads, nslots); break;
        case 2: col_inv_stage<2, CONJ_OUT>(x, M, m, tws, dm, W, tid, nthreads, nslots); break;
        case 3: col_inv_stage<3, CONJ_OUT>(x, M, m, tws, dm, W, tid, nthreads, nslots); break;
        default: col_inv_stage<5, CONJ_OUT>(x, M, m, tws, dm, W, tid, nthreads, nslots); break;
    }
}

#ifdef __CUDACC__
// forward: natural order in (positions >= nz taken as zero), digit-reversed frequencies out.  Ends with a barrier.
__device__ __forceinline__ void col_fft_forward(float2 *x, const ColPlan &pl, int nz, const float2 *__restrict__ W) {
    col_run_fwd_stage<true>(x, pl, 0, nz, W, (int) threadIdx.x, (int) blockDim.x);
    __syncthreads();
    for (int st = 1; st < pl.nstages; st++) {
        col_run_fwd_stage<false>(x, pl, st, nz, W, (int) threadIdx.x, (int) blockDim.x);
        __syncthreads();
    }
}
// inverse (unnormalised) of CONJUGATED digit-reversed spectra: natural order, un-conjugated, out.  Ends with a barrier.
__device__ __forceinline__ void col_fft_inverse_conj(float2 *x, const ColPlan &pl, const float2 *__restrict__ W, int nslots) {
    for (int st = pl.nstages - 1; st > 0; st--) {
        col_run_inv_stage<false>(x, pl, st, W, (int) threadIdx.x, (int) blockDim.x, nslots);
        __syncthreads();
    }
    col_run_inv_stage<true>(x, pl, 0, W, (int) threadIdx.x, (int) blockDim.x, nslots);
    __syncthreads();
}

// ------------------------------------------------------------------------------ TMA / mbarrier wrappers --
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 2-D tiled tensor copy global -> shared (coordinates: c0 = innermost); completion is signalled on the mbarrier
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *tm, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *tm, int c0, int c1, const void *smem_src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit_and_wait() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ----------------------------------------------------------------------------------- charge side: rows --
// Buffers (2-D contexts):
//   chg[G*G] float4   spread result, node (row = y node, col = x node) at row*G + col: (w1, delta_x, delta_y, wbb), box units
//   S[rows][M/2+1][4] float2   per row r and kx: the four x-spectra, later (in place) v1~, Bx~, By~
//   pot[G*G] float4   (v1, Bx, By, 0) at the nodes, what the gather reads
// One CTA per grid row r < G (the launch covers M/2 rows; the rest exit): both packed sequences of the row through the
// Stockham FFT of fitsne_fft.cuh (natural order out), then the real/imaginary-part spectra are separated with the row's
// own mirror bins.
// Sharded (p2p != 0): the row is the SUM of every rank's partial spread grid, added in rank order while loading (peer
// memory over NVLink; identical bits on every rank) -- the grid all-reduce costs no pass of its own.  p2p == 1: every rank
// transforms every row (replicated convolution); p2p == 2: distributed convolution, my block of rows only, every bin
// stored into the S of the rank that owns its column.
__global__ void __launch_bounds__(ROW_THREADS) k_conv_rows_fwd(const float4 *__restrict__ chg, float2 *__restrict__ S, FftPlan plan,
                                                               const float2 *__restrict__ W, const GridParams *__restrict__ gpp,
                                                               PeerComm pc, int p2p, unsigned int *__restrict__ ticket) {
    pdl_prologue();
    extern __shared__ __align__(16) float2 row_sm[];
    __shared__ FftPlan plan_s;
    const int G = gpp->G;
    bool live = gpp->ok && (int) blockIdx.x < G;
    if (p2p == 2 && live) live = part_owner(blockIdx.x, part_block(G, pc.world, 1), pc.world) == pc.rank;   // distributed: my rows only
    if (live) {
    if (p2p && threadIdx.x == 0) peer_wait(pc.flags[pc.rank], FLAG_GRID, pc, *reinterpret_cast<volatile unsigned int *>(pc.seq));
    for (int i = threadIdx.x; i < (int) (sizeof(FftPlan) / 4); i += blockDim.x)
        reinterpret_cast<int *>(&plan_s)[i] = reinterpret_cast<const int *>(&plan)[i];
    const int M = plan.n, NS = fft_buf_len(M, 2), r = blockIdx.x;
    float2 *bufa = row_sm, *bufb = row_sm + 2 * NS;
    const float4 *src = chg + (size_t) r * G;
    if (p2p) __syncthreads();                         // thread 0 has seen every rank's "grid complete" flag
    for (int c = threadIdx.x; c < M; c += blockDim.x) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < G) {
            if (!p2p) v = src[c];
            else {
                float4 t[MAX_RANKS];
#pragma unroll
                for (int q = 0; q < MAX_RANKS; q++)       // all ranks' loads in flight together, then added in rank order
                    t[q] = q < pc.world ? __ldcg(reinterpret_cast<const float4 *>(pc.grid[q]) + (size_t) r * G + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int q = 0; q < MAX_RANKS; q++) { v.x += t[q].x; v.y += t[q].y; v.z += t[q].z; v.w += t[q].w; }
            }
        }
        const int ph = fft_phys(c);
        bufa[ph] = make_float2(v.x, v.y);            // w1 + i delta_x
        bufa[NS + ph] = make_float2(v.z, v.w);       // delta_y + i wbb
    }
    __syncthreads();
    const float2 *res = fft_smem(bufa, bufb, NS, 2, plan_s, W);
    const int H = M / 2 + 1;
    const int kblock = p2p == 2 ? part_block(H, pc.world, 2) : H;
    for (int kx = threadIdx.x; kx < H; kx += blockDim.x) {
        const int km = kx ? M - kx : 0;
        float2 w1, dx, dy, wb;
        unpack_pair(res[fft_phys(kx)], res[fft_phys(km)], w1, dx);
        unpack_pair(res[NS + fft_phys(kx)], res[NS + fft_phys(km)], dy, wb);
        // sharded: the bin goes to the S of the rank that owns column kx (a store on peer memory: the transpose of a
        // distributed 2-D FFT, done by the producing kernel)
        float2 *Sd = p2p == 2 ? pc.S[part_owner(kx, kblock, pc.world)] : S;
        float4 *dst = reinterpret_cast<float4 *>(Sd + ((size_t) r * H + kx) * COL_SLOTS);
        dst[0] = make_float4(w1.x, w1.y, dx.x, dx.y);
        dst[1] = make_float4(dy.x, dy.y, wb.x, wb.y);
    }
    }   // live
    if (p2p == 2) peer_signal_last(ticket, pc, FLAG_S1, live ? 2 : 0);     // last CTA: "my rows are in every column owner's S"
}

// rows r < G: x-half-spectra (v1~, Bx~, By~) -> real rows.  Z1 = v1~ + i Bx~ extended by Hermitian symmetry gives
// v1 + i Bx in one complex inverse transform; By takes the second.  Inverse = conj(FFT(conj .)).
// Sharded: my rows only (their bins were written into my S by the ranks that own the columns); the finished row of the
// potential grid goes to EVERY rank's pot (each rank gathers for its own points anywhere in the plane), and the first CTA
// adds the ranks' sum_Q partials in rank order.
__global__ void __launch_bounds__(ROW_THREADS) k_conv_rows_inv(const float2 *S, float4 *__restrict__ pot, FftPlan plan,
                                                               const float2 *__restrict__ W, const GridParams *__restrict__ gpp,
                                                               PeerComm pc, int p2p, int N, Scalars *__restrict__ sc,
                                                               unsigned int *__restrict__ ticket) {
    pdl_prologue();
    extern __shared__ __align__(16) float2 row_sm[];
    __shared__ FftPlan plan_s;
    const int G = gpp->G;
    bool live = gpp->ok && (int) blockIdx.x < G;
    if (p2p == 2 && live) {
        const bool mine = part_owner(blockIdx.x, part_block(G, pc.world, 1), pc.world) == pc.rank;
        if (threadIdx.x == 0 && (mine || blockIdx.x == 0)) {
            peer_wait(pc.flags[pc.rank], FLAG_S2, pc, *reinterpret_cast<volatile unsigned int *>(pc.seq));
            if (blockIdx.x == 0) {        // sum_Q = (sum over ranks of their columns' Parseval partials, in rank order) - N
                double tot = 0;
                for (int q = 0; q < pc.world; q++) tot += *reinterpret_cast<const volatile double *>(pc.zs[pc.rank] + q);
                const double Z = tot - (double) N;
                sc->Z = Z;
                sc->inv_Z = (float) (1.0 / Z);
            }
        }
        live = mine;
        __syncthreads();
    }
    if (live) {
    for (int i = threadIdx.x; i < (int) (sizeof(FftPlan) / 4); i += blockDim.x)
        reinterpret_cast<int *>(&plan_s)[i] = reinterpret_cast<const int *>(&plan)[i];
    const int M = plan.n, NS = fft_buf_len(M, 2), r = blockIdx.x, H = M / 2 + 1;
    float2 *bufa = row_sm, *bufb = row_sm + 2 * NS;
    const float4 *src = reinterpret_cast<const float4 *>(S + (size_t) r * H * COL_SLOTS);
    for (int kx = threadIdx.x; kx < H; kx += blockDim.x) {
        const float4 a = __ldcg(src + 2 * kx), b = __ldcg(src + 2 * kx + 1);    // (v1, Bx), (By, -)
        const int km = kx ? M - kx : 0;
        // conj(Z1[kx]) with Z1[kx] = v1 + i Bx;  conj(Z1[-kx]) with Z1[-kx] = conj(v1) + i conj(Bx)
        bufa[fft_phys(kx)] = make_float2(a.x - a.w, -(a.y + a.z));
        bufa[NS + fft_phys(kx)] = make_float2(b.x, -b.y);
        if (km != kx) {
            bufa[fft_phys(km)] = make_float2(a.x + a.w, -(a.z - a.y));
            bufa[NS + fft_phys(km)] = make_float2(b.x, b.y);
        }
    }
    __syncthreads();
    const float2 *res = fft_smem(bufa, bufb, NS, 2, plan_s, W);
    for (int c = threadIdx.x; c < G; c += blockDim.x) {
        const float2 o1 = res[fft_phys(c)], o2 = res[NS + fft_phys(c)];
        const float4 v = make_float4(o1.x, -o1.y, o2.x, 0.f);
        if (p2p != 2) pot[(size_t) r * G + c] = v;
        else {
#pragma unroll
            for (int q = 0; q < MAX_RANKS; q++) if (q < pc.world) pc.pot[q][(size_t) r * G + c] = v;
        }
    }
    }   // live
    if (p2p == 2) peer_signal_last(ticket, pc, FLAG_POT, live ? 2 : 0);    // last CTA: "my rows of the potential grid are everywhere"
}

// ---------------------------------------------------------------------------------- kernel side: rows --
// Row dr in [0, G) of the wrap-around node-offset lattice (nbodyfft.cpp:52-61 with M >= 2G): samples in fp64,
//   z[c] = (Ksq + Kgrad_x) + i Kb,   Ksq=(1+r2/df)^-(df+1), Kb=(1+r2/df)^-df, Kgrad_x = (dc/p) Ksq   (tsne.cpp:69-94)
// carrying the 1/M^2 normalisation (nbodyfft.cpp:202-203).  Ksq, Kb are even in dc, Kgrad_x is odd, so
//   Re Z = Ksq~,  Im Z = a~ + Kb~  with a~ odd:   KR[dr][kx] = (Ksq~, Kb~, a~, 0)  for kx <= M/2   (Kgrad_x~ = i a~).
__global__ void __launch_bounds__(ROW_THREADS) k_kspec_rows(float4 *__restrict__ KR, FftPlan plan, const float2 *__restrict__ W,
                                                            const GridParams *__restrict__ gpp, double df) {
    pdl_prologue();
    extern __shared__ __align__(16) float2 row_sm[];
    __shared__ FftPlan plan_s;
    const int G = gpp->G;
    if (!gpp->ok || (int) blockIdx.x >= G) return;
    for (int i = threadIdx.x; i < (int) (sizeof(FftPlan) / 4); i += blockDim.x)
        reinterpret_cast<int *>(&plan_s)[i] = reinterpret_cast<const int *>(&plan)[i];
    const int M = plan.n, NS = fft_buf_len(M, 1), dr = blockIdx.x;
    const double h2 = gpp->h * gpp->h, inv_norm = gpp->inv_norm, inv_p = 1.0 / (double) gpp->p;
    float2 *bufa = row_sm, *bufb = row_sm + NS;
    for (int c = threadIdx.x; c < M; c += blockDim.x) {
        const int dc = c < G ? c : (c > M - G ? c - M : 0);
        float2 z = make_float2(0.f, 0.f);
        if (c < G || c > M - G) {
            const double r2 = h2 * ((double) dc * (double) dc + (double) dr * (double) dr);
            double kb, ksq;
            if (df == 1.0) { kb = 1.0 / (1.0 + r2); ksq = kb * kb; }
            else { const double t = 1.0 + r2 / df; kb = pow(t, -df); ksq = pow(t, -(df + 1.0)); }
            kb *= inv_norm; ksq *= inv_norm;
            z = make_float2((float) (ksq + (double) dc * inv_p * ksq), (float) kb);
        }
        bufa[fft_phys(c)] = z;
    }
    __syncthreads();
    const float2 *res = fft_smem(bufa, bufb, NS, 1, plan_s, W);
    const int H = M / 2 + 1;
    float4 *dst = KR + (size_t) dr * H;
    for (int kx = threadIdx.x; kx < H; kx += blockDim.x) {
        const float2 zk = res[fft_phys(kx)], zm = res[fft_phys(kx ? M - kx : 0)];
        dst[kx] = make_float4(0.5f * (zk.x + zm.x), 0.5f * (zk.y + zm.y), 0.5f * (zk.y - zm.y), 0.f);
    }
}

// ------------------------------------------------------------------------------- kernel side: columns --
// Tile = two adjacent kx.  For each, the length-M column over the row offset (rows >= G mirrored from dr < 0, zero in the
// gap) of   zA = Ksq~ + i Kb~   (both even in dr -> real spectra Ksq^, Kb^)   and   zB = a~ - (dr/p) Ksq~   (even - odd:
// Re ZB = gx with Kgrad_x^ = i gx,  Im ZB = -gy with Kgrad_y^ = i gy).   KS[kx][pos] = (Ksq^, Kb^, gx, gy), pos = the
// forward transform's digit-reversed frequency slot -- the order k_conv_cols meets them in.
template <int BOUND>
__global__ void __launch_bounds__(BOUND, BOUND == COL_THREADS ? 4 : 1) k_kspec_cols(const float4 *__restrict__ KR, float4 *__restrict__ KS, ColPlan plan,
                                                            const float2 *__restrict__ W, const GridParams *__restrict__ gpp,
                                                            int rank, int world) {
    pdl_prologue();
    extern __shared__ __align__(128) float2 col_sm[];
    __shared__ ColPlan plan_s;
    if (!gpp->ok) return;
    // sharded: only the columns whose convolution this rank performs (blocks are even-sized: a tile never straddles two ranks)
    if (world > 1 && part_owner(2 * blockIdx.x, part_block(plan.n / 2 + 1, world, 2), world) != rank) return;
    for (int i = threadIdx.x; i < (int) (sizeof(ColPlan) / 4); i += blockDim.x)
        reinterpret_cast<int *>(&plan_s)[i] = reinterpret_cast<const int *>(&plan)[i];
    const int M = plan.n, G = gpp->G, H = M / 2 + 1;
    const int kx0 = 2 * blockIdx.x, kx1 = kx0 + 1;
    const bool has1 = kx1 < H;
    const float inv_p = 1.f / (float) gpp->p;
    float2 *x = col_sm;
    for (int ry = threadIdx.x; ry < M; ry += blockDim.x) {
        const int dr = ry < G ? ry : (ry > M - G ? ry - M : 0);
        const bool valid = ry < G || ry > M - G;
        const int ad = dr < 0 ? -dr : dr;
        const float uy = (float) dr * inv_p;
        float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
        if (valid) {
            v0 = __ldg(KR + (size_t) ad * H + kx0);
            if (has1) v1 = __ldg(KR + (size_t) ad * H + kx1);
        }
        float4 *d = reinterpret_cast<float4 *>(x + ry * COL_SLOTS);
        d[0] = make_float4(v0.x, v0.y, v0.z - uy * v0.x, 0.f);
        d[1] = make_float4(v1.x, v1.y, v1.z - uy * v1.x, 0.f);
    }
    __syncthreads();
    col_fft_forward(x, plan_s, M, W);
    for (int pos = threadIdx.x; pos < M; pos += blockDim.x) {
        const float4 *s = reinterpret_cast<const float4 *>(x + pos * COL_SLOTS);
        const float4 a = s[0], b = s[1];
        KS[(size_t) kx0 * M + pos] = make_float4(a.x, a.y, a.z, -a.w);
        if (has1) KS[(size_t) kx1 * M + pos] = make_float4(b.x, b.y, b.z, -b.w);
    }
}

// ------------------------------------------------------------------------- charge side: fused column pass --
// One CTA per kx in [0, M/2].  TMA brings rows [0, G) x 32 bytes of S (the column's four spectra w1~, dx~, dy~, wbb~) into
// the tile, four forward FFTs run in place, then per frequency slot (nbodyfft.cpp:184-191 for all terms at once)
//   v1 = Ksq^ w1^,   B_k = Kgrad_k^ w1^ - Ksq^ delta_k^ = i g_k w1^ - Ksq^ delta_k^
// and, by Parseval in fp64, this column's share of sum_Q (weight 2 for 0 < kx < M/2: the mirrored half is not stored)
//   df==1: <w1,Kb*w1> + 2<wbb,v1> + sum_k (4<delta_k,Kgrad_k*w1> - 2<delta_k,Ksq*delta_k>)   (tsne.cpp:1101-1110)
//   df!=1: <w1,Kb*w1>                                                                          (tsne.cpp:950-955)
// (delta, wbb, Kgrad, B are in box units: the bracketed terms carry bw^2).  Three inverse FFTs in place, TMA stores the
// tile back over its input.  The last CTA to finish adds the per-column partials in index order: sum_Q, 1/sum_Q.
template <int BOUND>
__global__ void __launch_bounds__(BOUND, BOUND == COL_THREADS ? 4 : 1) k_conv_cols(const __grid_constant__ CUtensorMap tmS, const float4 *__restrict__ KS,
                                                           ColPlan plan, const float2 *__restrict__ W, const GridParams *__restrict__ gpp,
                                                           int df_is_one, double *__restrict__ zpartial, int N, Scalars *__restrict__ sc,
                                                           unsigned int *__restrict__ ticket, PeerComm pc, int p2p) {
    pdl_prologue();
    extern __shared__ __align__(128) float2 col_sm[];
    __shared__ ColPlan plan_s;
    __shared__ __align__(8) uint64_t mbar;
    __shared__ double red[32];
    if (!gpp->ok) {                                      // (the flag still goes out: one per iteration keeps the ranks in step)
        if (p2p == 2) peer_signal_last(ticket, pc, FLAG_S2, 0);
        return;
    }
    const int M = plan.n, G = gpp->G, kx = blockIdx.x;
    // sharded: my block of columns only (the other ranks' CTAs leave; their partial-sum slots must read as zero), and not
    // before every rank's rows have landed in my S
    const bool dist = p2p == 2;
    const bool mine = !dist || part_owner(kx, part_block(M / 2 + 1, pc.world, 2), pc.world) == pc.rank;
    const int nbox = (G + COL_BOX_ROWS - 1) / COL_BOX_ROWS;
    float2 *x = col_sm;
    if (threadIdx.x == 0) { mbar_init(&mbar, 1); fence_mbar_init(); }
    for (int i = threadIdx.x; i < (int) (sizeof(ColPlan) / 4); i += blockDim.x)
        reinterpret_cast<int *>(&plan_s)[i] = reinterpret_cast<const int *>(&plan)[i];
    __syncthreads();
    double zacc = 0;
    if (mine) {
    if (threadIdx.x == 0) {
        if (dist) {
            peer_wait(pc.flags[pc.rank], FLAG_S1, pc, *reinterpret_cast<volatile unsigned int *>(pc.seq));
            asm volatile("fence.proxy.async;" ::: "memory");      // the peers' (generic-proxy) stores, then my TMA (async-proxy) reads
        }
        mbar_expect_tx(&mbar, (uint32_t) nbox * COL_BOX_ROWS * COL_SLOTS * (uint32_t) sizeof(float2));
        for (int i = 0; i < nbox; i++) tma_load_2d(x + (size_t) i * COL_BOX_ROWS * COL_SLOTS, &tmS, kx * 2 * COL_SLOTS, i * COL_BOX_ROWS, &mbar);
    }
    mbar_wait(&mbar, 0);
    col_fft_forward(x, plan_s, G, W);
    const double bw2 = gpp->bw * gpp->bw;
    const float4 *ks = KS + (size_t) kx * M;
    for (int pos = threadIdx.x; pos < M; pos += blockDim.x) {
        float4 *s = reinterpret_cast<float4 *>(x + pos * COL_SLOTS);
        const float4 k = __ldg(ks + pos);                   // (Ksq^, Kb^, gx, gy), all real
        const float4 a = s[0], b = s[1];
        const float2 w1 = make_float2(a.x, a.y), dx = make_float2(a.z, a.w), dy = make_float2(b.x, b.y), wb = make_float2(b.z, b.w);
        const float2 v1 = make_float2(k.x * w1.x, k.x * w1.y);
        const float2 gxw = make_float2(-k.z * w1.y, k.z * w1.x), gyw = make_float2(-k.w * w1.y, k.w * w1.x);   // i g w1
        const float2 Bx = make_float2(gxw.x - k.x * dx.x, gxw.y - k.x * dx.y), By = make_float2(gyw.x - k.x * dy.x, gyw.y - k.x * dy.y);
        double z = (double) k.y * ((double) w1.x * w1.x + (double) w1.y * w1.y);
        if (df_is_one) {
            const double zb = 2.0 * re_conj_mul(wb, v1) + 4.0 * (re_conj_mul(dx, gxw) + re_conj_mul(dy, gyw))
                              - 2.0 * (double) k.x * ((double) dx.x * dx.x + (double) dx.y * dx.y + (double) dy.x * dy.x + (double) dy.y * dy.y);
            z += bw2 * zb;
        }
        zacc += z;
        s[0] = make_float4(v1.x, -v1.y, Bx.x, -Bx.y);       // conjugated: the inverse runs as a forward transform
        s[1] = make_float4(By.x, -By.y, 0.f, 0.f);
    }
    __syncthreads();
    col_fft_inverse_conj(x, plan_s, W, 3);             // v1, Bx, By: the fourth slot is not transformed (nobody reads it)
    if (!dist) {
        fence_proxy_async_smem();
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int i = 0; i < nbox; i++) tma_store_2d(&tmS, kx * 2 * COL_SLOTS, i * COL_BOX_ROWS, x + (size_t) i * COL_BOX_ROWS * COL_SLOTS);
            tma_store_commit_and_wait();
        }
    } else {
        // sharded: row r of the convolved column belongs to the rank that owns row r -- 32-byte stores on peer memory
        const int H = M / 2 + 1, rblock = part_block(G, pc.world, 1);
        for (int r = threadIdx.x; r < G; r += blockDim.x) {
            const float4 *sv = reinterpret_cast<const float4 *>(x + r * COL_SLOTS);
            float4 *d = reinterpret_cast<float4 *>(pc.S[part_owner(r, rblock, pc.world)] + ((size_t) r * H + kx) * COL_SLOTS);
            d[0] = sv[0]; d[1] = sv[1];
        }
    }
    }   // mine
    const double rsum = block_sum(zacc * ((kx == 0 || 2 * kx == M) ? 1.0 : 2.0), red);
    if (threadIdx.x == 0) {
        zpartial[kx] = rsum;
        if (dist && mine) __threadfence_system();     // (after block_sum's barriers: the whole CTA's peer stores, before the ticket)
    }
    if (last_block_done(ticket)) {           // sum_Q = (sum of the partials, in index order) - N   (tsne.cpp:1110)
        double s2 = 0;
        for (int i = threadIdx.x; i < (int) gridDim.x; i += blockDim.x) s2 += ld_partial(zpartial + i);
        const double tot = block_sum(s2, red);
        if (threadIdx.x == 0) {
            if (!dist) {
                const double Z = tot - (double) N;
                sc->Z = Z;
                sc->inv_Z = (float) (1.0 / Z);
            } else {
                // my columns' share goes to slot [rank] of every rank's table; k_conv_rows_inv adds the slots in rank order
                for (int q = 0; q < pc.world; q++) *reinterpret_cast<volatile double *>(pc.zs[q] + pc.rank) = tot;
                // ... and "my columns are back in every row owner's S": the exchange flag, by the CTA that finished last
                __threadfence_system();
                const uint32_t seq = *reinterpret_cast<volatile unsigned int *>(pc.seq);
                for (int q = 0; q < pc.world; q++)
                    if (q != pc.rank) *reinterpret_cast<volatile uint32_t *>(pc.flags[q] + FLAG_S2 * pc.world + pc.rank) = seq;
            }
        }
    }
}
#endif  // __CUDACC__

}  // namespace fk
