// fitsne_kernels.cuh -- hand-written sm_100a kernels for FIt-SNE's per-iteration gradient loop.
//
// Everything the reference does per iteration (reference = /root/reference/src/...) in fp32 on the device:
//   bounds + zero-mean        tsne.cpp:1039-1049, :1851-1876        k_center_bounds, k_setup_grid (sharded: k_update_shard, k_center_shard)
//   point -> box, sort        nbodyfft.cpp:85-114                    k_bin, k_radix_sweep x 2 (look-back)
//   Lagrange spread           nbodyfft.cpp:123-147, :310-336         k_spread_chunks, k_spread_combine
//   kernel samples + spectra  nbodyfft.cpp:52-68, tsne.cpp:69-94     2-D: k_kspec_rows / k_kspec_cols (fitsne_conv.cuh); 1-D: k_gen_kernels_1d + k_fft_line
//   convolution + sum_Q       nbodyfft.cpp:150-217, tsne.cpp:1101-1110  2-D: k_conv_rows_fwd / k_conv_cols / k_conv_rows_inv; 1-D: k_hadamard_1d
//   gather + normalise        nbodyfft.cpp:222-239, tsne.cpp:1149-1151  k_gather
//   attractive + optimiser    tsne.cpp:1121-1137, :479-513           k_attract (or k_attract_tiles), k_update
//   KL                        tsne.cpp:1329-1355                      k_kl
//
// The repulsive part uses the "local offset" formulation documented in tests/device_model.py (identical
// algebra to the reference's {1,x,y,x^2+y^2} charges, but fp32-safe); binning and in-box coordinates are
// evaluated in fp64 with the reference's exact operation order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <float.h>
#include <type_traits>

namespace fk {

constexpr int PMAX = 16;          // max interpolation points per box and axis
constexpr int CHUNK = 8;          // points per spread work item (one thread walks a chunk).  Measured on B200 (spread_chunks, us, chunk
                                  // 1 / 2 / 4 / 8): N=1M 135 / 78 / 49 / 36, N=500k 76 / 47 / 34 / 25, N=125k 32 / 23 / 20 / 19 -- longer
                                  // chunks win at every size (fewer partials to park and stitch), so slices of a sharded run keep 8 too
constexpr int SORT_THREADS = 256;
constexpr int SORT_IPT = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_IPT;
constexpr int SORT_MAX_BITS = 11;     // widest radix digit
constexpr int SORT_ONE_PASS_BITS = 8; // keys this short sort in ONE pass.  (Measured on B200: a single 12-bit pass -- 4096 bins --
                                      // is slower than two 6-bit passes: 84 vs 70 us at N=1M; so only short 1-D keys qualify.)
constexpr int RED_BLOCKS = 1184;  // 8 x 148 SMs: partial-reduction width for bounds / column sums
constexpr int Z_BLOCKS_1D = 32;      // Parseval partials of the 1-D Hadamard kernel

// Device-resident description of this iteration's interpolation grid.  Rewritten every iteration by
// k_setup_grid from the bounds; all kernels read it from memory so that one captured CUDA graph serves every
// iteration that has the same n_boxes.
struct GridParams {
    double mn, mx;        // min_coord / max_coord as the reference computes them (tsne.cpp:1042-1049)
    double bw;            // (mx-mn)/B                       nbodyfft.cpp:16
    double bw2;           // (1*bw+mn) - (0*bw+mn)            nbodyfft.cpp:79-80
    double h;             // node spacing bw/p               nbodyfft.cpp:40
    double inv_norm;      // 1/M^dims, folded into the kernel spectra (nbodyfft.cpp:202-203)
    float bwf;
    int B, G, M, p, xbits, nb, ok;
    int sort_bits;        // radix digit width
    int sort_passes;      // 1: the whole key is one digit (<= SORT_MAX_BITS bits); 2: two LSD passes of sort_bits each
    int pad_[3];
    float s[PMAX];        // in-box node positions (k+1/2)/p, accumulated like nbodyfft.cpp:30-34
    float inv_den[PMAX];  // 1/prod_{j!=i}(s_i-s_j)          nbodyfft.cpp:313-321
};

struct StepParams {
    float alpha;          // exaggeration on P
    float momentum, lr, max_step_norm;
    int mode;             // FITSNE_STEP_*
    float inv_df;
};

struct Scalars {          // small device-resident results
    double Z;             // sum_Q (tsne.cpp:1112)
    float inv_Z;
    float pad;
    double mean[2];
    float bmin, bmax;     // bounds of the current Y (with the 2-D scan quirk)
    double kl;
    unsigned long long iter_done;   // optimiser steps that really executed (speculative launches that found a grid mismatch do not count)
};

// ---------------------------------------------------------------------------- peer-memory fabric (sharded) --
// One process per GPU; every rank maps the others' exchange buffers (CUDA IPC over NVLink / NVSwitch) and the three
// per-iteration exchanges are done by the kernels themselves, with loads / stores on peer memory -- no library
// collective inside the iteration:
//   grid reduction   k_conv_rows_fwd (2-D) / k_grid_sum_1d add the ranks' partial spread grids WHILE LOADING them, in rank
//                    order, so every rank gets the same bits;
//   statistics       k_update_shard writes its 128-byte record straight into every peer's table, k_center_shard reads them;
//   positions        every rank's centred slice of Y is pushed into the peers' Y by the copy engines (a side stream:
//                    DMA over NVLink, no SM involved) while the next iteration sorts and spreads; the SpMV waits for it.
// Ordering: flags[kind * world + r] in MY memory is written by rank r with the iteration's sequence number after its data
// for that exchange is in place (system-scope fence in between); consumers spin on their own flags.  A rank can never be
// more than one exchange ahead of its peers (each exchange needs everybody's signal), which is what makes single
// buffers safe: see DESIGN.md section 7 for the three hazard arguments.
//   convolution      (2-D) distributed like a 2-D FFT: rows of the grid and columns of the spectrum are dealt out to the
//                    ranks in contiguous blocks; k_conv_rows_fwd writes each x-spectrum bin into the S of the rank that
//                    owns its column, k_conv_cols writes each convolved bin into the S of the rank that owns its row,
//                    k_conv_rows_inv writes its rows of the potential grid into everybody's pot; sum_Q partials travel
//                    the same way.  The "transposes" are the kernels' own stores on peer memory.
constexpr int MAX_RANKS = 8;
constexpr int FLAG_Y = 0, FLAG_GRID = 1, FLAG_STATS = 2, FLAG_S1 = 3, FLAG_S2 = 4, FLAG_POT = 5, FLAG_KINDS = 6;
struct PeerComm {
    float *Y[MAX_RANKS];             // every rank's Y (full N x D floats); [rank] = my own
    void *grid[MAX_RANKS];           // every rank's partial spread grid: chg (2-D, float4 per node) or planes (1-D)
    void *stats[MAX_RANKS];          // every rank's ShardStats[world] table
    uint32_t *flags[MAX_RANKS];      // every rank's flag words [FLAG_KINDS * world]
    float2 *S[MAX_RANKS];            // 2-D: every rank's S (x-spectra / convolved half-spectra)
    float4 *pot[MAX_RANKS];          // 2-D: every rank's potential grid
    double *zs[MAX_RANKS];           // 2-D: every rank's table of per-rank sum_Q partials [world]
    unsigned int *seq;               // my sequence counter (bumped once per enqueued iteration by k_setup_grid)
    int rank, world;
};
// First statement of every kernel of the iteration chain (see launch_k in fitsne_capi.cu).  Launched with programmatic
// stream serialisation, a kernel may be set up while its predecessor drains; "wait" blocks until the
// predecessor grid has completed and its memory is visible -- nothing before it may touch global memory.  A no-op in a
// plain launch.  (No early "launch_dependents": measured on B200 with the trigger at the top of every kernel, the
// dependents' CTAs became resident beside multi-wave predecessors and took their SM slots -- 1M points 2518 -> 2073 it/s,
// 10M 349 -> 175 it/s.  The implicit trigger at grid exit keeps the order of the waves intact.)
__device__ __forceinline__ void pdl_prologue() {
#if defined(__CUDA_ARCH__)
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
// contiguous block partition of n items over the ranks (block size rounded up to a multiple of `align`)
__host__ __device__ __forceinline__ int part_block(int n, int world, int align) {
    const int b = (n + world - 1) / world;
    return (b + align - 1) / align * align;
}
__host__ __device__ __forceinline__ int part_owner(int i, int block, int world) {
    const int q = i / block;
    return q < world ? q : world - 1;
}

__device__ __forceinline__ void peer_wait(const uint32_t *flags, int kind, const PeerComm &pc, uint32_t seq) {
    // one thread: until every peer's flag of this kind has reached the iteration's sequence number
    for (int r = 0; r < pc.world; r++) {
        if (r == pc.rank) continue;
        const volatile uint32_t *f = flags + kind * pc.world + r;
        while ((int) (*f - seq) < 0) { }
    }
    __threadfence_system();
}

// tell every peer that my data for exchange `kind` is in place (one thread; everything written before is fenced first)
__global__ void k_peer_signal(PeerComm pc, int kind) {
    pdl_prologue();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    __threadfence_system();
    const uint32_t seq = *reinterpret_cast<volatile unsigned int *>(pc.seq);
    for (int r = 0; r < pc.world; r++)
        if (r != pc.rank) *reinterpret_cast<volatile uint32_t *>(pc.flags[r] + kind * pc.world + pc.rank) = seq;
}
// stream-level wait: one thread spins until every peer has signalled `kind` (and `kind2`, if >= 0) for this iteration
__global__ void k_peer_wait(PeerComm pc, int kind, int kind2) {
    pdl_prologue();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const uint32_t seq = *reinterpret_cast<volatile unsigned int *>(pc.seq);
    peer_wait(pc.flags[pc.rank], kind, pc, seq);
    if (kind2 >= 0) peer_wait(pc.flags[pc.rank], kind2, pc, seq);
}
// Producer-side signal, folded into the kernel that finishes the data: EVERY CTA of the grid calls this (all threads, on
// every exit path) after its last store; each CTA publishes its stores and takes a ticket, the CTA that arrives last
// raises my flag of this kind at every peer.  One graph node less per exchange than a separate one-thread k_peer_signal,
// and the flag leaves as soon as the last CTA is done instead of after the kernel has drained.
//   wrote: 0 = this CTA stored nothing the peers will read (no fence: a system-scope fence per CTA of a wide grid costs
//   more than the launch it replaces -- measured +12 us on k_spread_combine's 1184 CTAs), 1 = it stored into LOCAL memory
//   the peers read over NVLink (device-scope fence: their loads are served by this GPU's L2), 2 = it stored into PEER
//   memory (system-scope fence: the stores have to have landed before the ticket is taken).
__device__ __forceinline__ void peer_signal_last(unsigned int *ticket, const PeerComm &pc, int kind, int wrote) {
    __syncthreads();
    if (threadIdx.x == 0) {
        if (wrote == 2) __threadfence_system();
        else if (wrote == 1) __threadfence();
        if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
            *ticket = 0;
            __threadfence_system();
            const uint32_t seq = *reinterpret_cast<volatile unsigned int *>(pc.seq);
            for (int r = 0; r < pc.world; r++)
                if (r != pc.rank) *reinterpret_cast<volatile uint32_t *>(pc.flags[r] + kind * pc.world + pc.rank) = seq;
        }
    }
}

// 1-D sharded runs: planes 0, 1 (two packed complex lines of length M) <- sum over ranks of their partial lines, in rank
// order.  Every rank keeps its partial in `partial` (peer-readable) and writes the sum into its own FFT input.
__global__ void __launch_bounds__(256) k_grid_sum_1d(PeerComm pc, float2 *__restrict__ planes, int n /* 2 * M */, const int *__restrict__ ok) {
    pdl_prologue();
    if (!*ok) return;
    if (threadIdx.x == 0) peer_wait(pc.flags[pc.rank], FLAG_GRID, pc, *reinterpret_cast<volatile unsigned int *>(pc.seq));
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float2 acc = make_float2(0.f, 0.f);
    for (int r = 0; r < pc.world; r++) {
        const float2 v = __ldcg(reinterpret_cast<const float2 *>(pc.grid[r]) + i);
        acc.x += v.x; acc.y += v.y;
    }
    planes[i] = acc;
}

// ------------------------------------------------------------------------------------------ helpers --

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic block sum (fixed tree).  sm must hold 32 values.  Result valid in thread 0.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T *sm) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sm[w] = v;
    __syncthreads();
    T r = 0;
    if (w == 0) {
        r = lane < nw ? sm[lane] : T(0);
        r = warp_sum(r);
    }
    return r;
}

// "Last block done": every block deposits its partial result, then calls this; it returns true (in all threads) only
// in the block that arrives last, which can then reduce the partials in a FIXED order -- a deterministic two-level
// reduction without a second kernel launch.  The ticket counter resets itself for the next launch.
__device__ __forceinline__ bool last_block_done(unsigned int *counter) {
    // Contract: the block's partial result was written by THREAD 0 (block_sum / the warp-0 reductions leave it there).
    // Only that thread has to make it visible before taking a ticket; a __threadfence() in every thread would make all
    // warps wait for the acknowledgement of their own bulk stores (Y, uY, gains ...) before they can retire.
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int t = atomicAdd(counter, 1u);
        is_last = (t == gridDim.x - 1);
        if (is_last) { *counter = 0; __threadfence(); }
    }
    __syncthreads();
    return is_last;
}
// partials written by other blocks are read around L1 (ld.global.cg) in the last block
__device__ __forceinline__ double ld_partial(const double *p) { return __ldcg(p); }
__device__ __forceinline__ float2 ld_partial(const float2 *p) { return __ldcg(p); }

// box index and in-box coordinate of one coordinate, fp64, reference operation order
// (nbodyfft.cpp:86-113 / :350-363); __d*_rn keeps nvcc from contracting into FMAs the CPU does not use.
template <bool CLAMP_LOW>
__device__ __forceinline__ int box_of(float y, const GridParams &gp, float &u) {
    const double yd = (double) y;
    int idx = (int) __ddiv_rn(__dsub_rn(yd, gp.mn), gp.bw2);
    if (idx >= gp.B) idx = gp.B - 1;
    else if (CLAMP_LOW && idx < 0) idx = 0;
    if (!CLAMP_LOW && idx < 0) idx = 0;   // 1-D reference has no lower clamp (would index out of bounds); y>=min there
    const double lower = __dadd_rn(__dmul_rn((double) idx, gp.bw), gp.mn);
    u = (float) __ddiv_rn(__dsub_rn(yd, lower), gp.bw2);
    return idx;
}

// ------------------------------------------------------------------------- column sums, centring, bounds --

// Yout = Yin - mean (tsne.cpp:1851-1876) and the bounds of the centred values.
// 2-D min follows the reference's `if (>max) .. else if (<min)` scan (tsne.cpp:1045-1048): values in the strictly
// ascending prefix of the interleaved sequence x0,y0,x1,y1,... (ORIGINAL point order) only ever update max, so they are
// never considered for the min.  The prefix is a handful of values long (k values with probability 1/k!), so: every CTA
// leaves the first BOUNDS_HEAD points (original order) out of its minimum, and the last CTA to finish replays the scan on
// exactly those points.  Exact unless the prefix is longer than 2*BOUNDS_HEAD = 64 values (probability 1/64!).
// 1-D uses plain min/max (tsne.cpp:769-772).
constexpr int BOUNDS_HEAD = 32;
template <int D>
__global__ void __launch_bounds__(256) k_center_bounds(const float *__restrict__ Yin, float *__restrict__ Yout, int N,
                                                       int do_center, float2 *__restrict__ bounds_partial,
                                                       Scalars *__restrict__ sc, const uint32_t *__restrict__ orig_of,
                                                       const uint32_t *__restrict__ pos_of, const GridParams *__restrict__ gpp,
                                                       volatile float *host_bounds, unsigned int *__restrict__ ticket) {
    pdl_prologue();
    if (gpp && !gpp->ok) return;
    __shared__ float smf[64];
    double mean[2] = {0, 0};
    if (do_center) {                 // k_update's last block reduced the column sums of the new positions
        for (int d = 0; d < D; d++) mean[d] = sc->mean[d];
    }
    const long long head = D == 2 ? BOUNDS_HEAD : 0;
    float mn = INFINITY, mx = -INFINITY;
    const int per = (N + gridDim.x - 1) / gridDim.x;
    const int b = blockIdx.x * per, e = min(N, b + per);
    for (int i = b + threadIdx.x; i < e; i += blockDim.x) {
        if (D == 2) {
            float2 v = reinterpret_cast<const float2 *>(Yin)[i];
            v.x = (float) ((double) v.x - mean[0]);
            v.y = (float) ((double) v.y - mean[1]);
            if (do_center) reinterpret_cast<float2 *>(Yout)[i] = v;
            mx = fmaxf(mx, fmaxf(v.x, v.y));
            const long long o = orig_of ? (long long) orig_of[i] : (long long) i;      // original index of this point
            if (o >= head) mn = fminf(mn, fminf(v.x, v.y));
        } else {
            float v = (float) ((double) Yin[i] - mean[0]);
            if (do_center) Yout[i] = v;
            mx = fmaxf(mx, v);
            mn = fminf(mn, v);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { smf[w] = mn; smf[32 + w] = mx; }
    __syncthreads();
    if (w == 0) {
        const int nw = blockDim.x >> 5;
        mn = lane < nw ? smf[lane] : INFINITY;
        mx = lane < nw ? smf[32 + lane] : -INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if (lane == 0) bounds_partial[blockIdx.x] = make_float2(mn, mx);
    }
    // last block: combine the per-block bounds, replay the scan on the head points, publish (device scalars + host-mapped
    // words); closing a full optimiser step (gpp != nullptr) also bumps the executed-iterations counter
    if (last_block_done(ticket)) {
        __shared__ float headv[2 * BOUNDS_HEAD];
        const int nhead = D == 2 ? min(2 * BOUNDS_HEAD, 2 * N) : 0;
        if ((int) threadIdx.x < nhead) {      // centred values of the head points, recomputed from the input (same arithmetic)
            const int j = threadIdx.x;
            const size_t pos = pos_of ? (size_t) pos_of[j >> 1] : (size_t) (j >> 1);
            headv[j] = (float) ((double) Yin[pos * 2 + (j & 1)] - mean[j & 1]);
        }
        float bmn = INFINITY, bmx = -INFINITY;
        for (int i = threadIdx.x; i < (int) gridDim.x; i += blockDim.x) {
            const float2 v = ld_partial(bounds_partial + i);
            bmn = fminf(bmn, v.x); bmx = fmaxf(bmx, v.y);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            bmn = fminf(bmn, __shfl_xor_sync(0xffffffffu, bmn, o));
            bmx = fmaxf(bmx, __shfl_xor_sync(0xffffffffu, bmx, o));
        }
        __syncthreads();
        if (lane == 0) { smf[w] = bmn; smf[32 + w] = bmx; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int i = 1; i < (int) (blockDim.x >> 5); i++) { bmn = fminf(bmn, smf[i]); bmx = fmaxf(bmx, smf[32 + i]); }
            float run = -INFINITY;
            bool ascending = true;
            for (int j = 0; j < nhead; j++) {          // the reference's scan on the head: max only while strictly ascending
                const float v = headv[j];
                if (ascending && v > run) run = v;
                else { ascending = false; bmn = fminf(bmn, v); }
            }
            sc->bmin = bmn; sc->bmax = bmx;
            if (gpp) sc->iter_done += 1;
            if (host_bounds) {
                host_bounds[0] = bmn; host_bounds[1] = bmx;
                if (gpp) *reinterpret_cast<volatile unsigned long long *>(host_bounds + 4) = sc->iter_done;
            }
        }
    }
}

// FFT length for a grid of side G = n/2.  Any M >= 2G-1 gives the same linear convolution (the reference uses 2G,
// nbodyfft.cpp:155-156); M is the next 2^a 3^b 5^c that is a multiple of 16 (32 above 512) -- the lengths the
// mixed-radix shared-memory FFT handles, on a ladder coarse enough that CUDA graphs are re-captured rarely.
__host__ __device__ inline int nice_fft_size(int n) {
    const int q = n <= 512 ? 16 : 32;
    n = (n + q - 1) / q * q;
    for (;; n += q) {
        int m = n;
        while (m % 2 == 0) m /= 2;
        while (m % 3 == 0) m /= 3;
        while (m % 5 == 0) m /= 5;
        if (m == 1) return n;
    }
}

// n_boxes exactly as the reference picks it (tsne.cpp:1065-1077 in 2-D, :774 in 1-D)
__host__ __device__ inline int choose_n_boxes(double mn, double mx, double ipi, int min_int, int dims) {
    const int allowed[20] = {25, 36, 50, 55, 60, 65, 70, 75, 80, 85, 90, 96, 100, 110, 120, 130, 140, 150, 175, 200};
    double v = (mx - mn) / ipi;
    double m = (double) min_int > v ? (double) min_int : v;   // fmax(min_num_intervals, span/ipi)
    if (v != v) m = (double) min_int;
    int n = (int) m;
    if (dims == 2 && n < allowed[19]) {
        int c = 0;
        while (allowed[c] < n) c++;
        n = allowed[c];
    }
    return n;
}

// Sort layout for n_boxes = B: one pass when the key (dims*xbits bits) fits a single digit, else two LSD passes of
// ceil(key bits / 2).  The launch sequence is the same either way (kernels of the unused pass return at once).
__host__ __device__ inline void sort_layout(int B, int dims, int *bits, int *passes) {
    int xb = 0;
    while ((1 << xb) < B) xb++;
    const int kb = dims * xb > 1 ? dims * xb : 1;
    if (kb <= SORT_ONE_PASS_BITS) { *passes = 1; *bits = kb; }
    else { *passes = 2; *bits = (kb + 1) / 2; }
}
__host__ __device__ inline int sort_bits_for(int B, int dims) {
    int bits, passes;
    sort_layout(B, dims, &bits, &passes);
    return bits;
}

// Fill GridParams for a grid of B boxes/dim (the host's choice; verified against the device's own bounds).
__global__ void k_setup_grid(GridParams *__restrict__ gp, const Scalars *__restrict__ sc, const int *__restrict__ B_host, int M, int p, int dims,
                             double ipi, int min_int, int *__restrict__ mismatch, uint32_t *__restrict__ sort_totals,
                             uint32_t *__restrict__ work, unsigned int *__restrict__ sweep_tickets, unsigned int *__restrict__ comm_seq) {
    pdl_prologue();
    for (int i = threadIdx.x; i < 2 * (1 << SORT_MAX_BITS); i += blockDim.x) sort_totals[i] = 0;   // both passes
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (comm_seq) *comm_seq += 1;                    // sharded: this iteration's sequence number (also for no-op iterations)
    work[0] = 0;                                     // spread work list (boxes that span several chunks) starts empty
    sweep_tickets[0] = 0; sweep_tickets[1] = 0;      // tile tickets of the two sort passes
    // B_host > 0: the host sized the grid after reading the bounds (single-step API); 0: speculative launch, the device
    // sizes the grid itself and the whole iteration becomes a no-op if that grid does not belong to this graph's M
    const int hostB = *reinterpret_cast<const volatile int *>(B_host);
    const double mn = (double) sc->bmin, mx = (double) sc->bmax;
    const int want = choose_n_boxes(mn, mx, ipi, min_int, dims);
    const int B = hostB > 0 ? hostB : want;
    gp->ok = (want == B) && (mx > mn) && (nice_fft_size(2 * B * p) == M);
    if (!gp->ok) *mismatch = want;
    gp->mn = mn; gp->mx = mx;
    gp->B = B; gp->p = p; gp->G = B * p; gp->M = M;
    const double bw = __ddiv_rn(__dsub_rn(mx, mn), (double) B);
    gp->bw = bw;
    const double lo0 = __dadd_rn(__dmul_rn(0.0, bw), mn);
    gp->bw2 = __dsub_rn(__dadd_rn(__dmul_rn(1.0, bw), mn), lo0);
    gp->bwf = (float) bw;
    gp->h = __dmul_rn(__ddiv_rn(1.0, (double) p), bw);
    gp->inv_norm = dims == 2 ? 1.0 / ((double) M * (double) M) : 1.0 / (double) M;
    int xb = 0;
    while ((1 << xb) < B) xb++;
    gp->xbits = xb;
    sort_layout(B, dims, &gp->sort_bits, &gp->sort_passes);
    gp->pad_[0] = gp->pad_[1] = gp->pad_[2] = 0;
    gp->nb = dims == 2 ? B * B : B;
    double s[PMAX];
    const double hh = 1.0 / (double) p;
    s[0] = hh / 2;
    for (int i = 1; i < p; i++) s[i] = s[i - 1] + hh;
    for (int i = 0; i < p; i++) {
        double den = 1;
        for (int j = 0; j < p; j++) if (i != j) den *= s[i] - s[j];
        gp->s[i] = (float) s[i];
        gp->inv_den[i] = (float) (1.0 / den);
    }
}

// -------------------------------------------------------------------------------------------- binning --
// Stable LSD radix sort of (box key, point index, in-box coordinates) in THREE launches (one pass when the key has
// <= SORT_ONE_PASS_BITS bits, else two passes of `sort_bits` bits each):
//   k_bin            box id + in-box coordinate (fp64, reference order); global totals of BOTH digits (shared-memory
//                    histograms, one integer atomic per digit and CTA); the last CTA turns the totals into exclusive digit
//                    bases; every CTA clears its tile's look-back words
//   k_radix_sweep#0  one CTA per tile of SORT_TILE keys: warp-level stable ranks, tile totals per digit published as ONE
//   k_radix_sweep#1  word (ready bit | count), exclusive prefix over the earlier tiles by summing their words as they
//                    appear, stable scatter.  Tiles take their index from an atomic ticket, so a tile's predecessors are
//                    always running or done: no deadlock, no separate histogram / offset kernels, no host involvement.
// The in-box coordinates ride through both scatters (8 bytes per point) instead of being recomputed from Y[perm] in fp64
// after the sort.  Counts are integers, so the result does not depend on timing: stability (ties keep point-index order)
// makes the box-sorted order, hence the spread's summation order, bitwise repeatable.  key = (by << xbits) | bx.
// (The point re-ordering every few hundred iterations sorts 22-bit Morton keys with the older four-kernel pass below.)

#ifdef __CUDA_ARCH__
#define FK_ATOMIC_ADD(ptr, v) atomicAdd((ptr), (v))
#else
#define FK_ATOMIC_ADD(ptr, v) fk_host_fetch_add((ptr), (v))
template <typename T>
static inline T fk_host_fetch_add(T *p, T v) { const T old = *p; *p = old + v; return old; }
#endif

constexpr int BIN_THREADS = 512;                        // k_bin: SORT_TILE / 512 = 8 points per thread; every tile's CTA is resident at once
constexpr int SWEEP_THREADS = 512;                      // k_radix_sweep: 16 warps x 8 keys per lane
constexpr int SWEEP_IPT = SORT_TILE / SWEEP_THREADS;
constexpr uint32_t SWEEP_AGGREGATE = 0x40000000u, SWEEP_INCLUSIVE = 0x80000000u, SWEEP_COUNT_MASK = 0x3fffffffu;   // look-back word = status | count

__device__ __forceinline__ void tile_hist_flush(const uint32_t *cnt, int nb, uint32_t *hist, int tiles, uint32_t *totals) {
    for (int i = threadIdx.x; i < nb; i += blockDim.x) {
        const uint32_t v = cnt[i];
        if (hist) hist[(size_t) i * tiles + blockIdx.x] = v;
        if (v) atomicAdd(&totals[i], v);
    }
}

// exclusive scan of totals[0..nb) into bases[0..nb) by one CTA (nb <= 2048)
__device__ __forceinline__ void block_excl_scan(const uint32_t *__restrict__ totals, uint32_t *__restrict__ bases, int nb, uint32_t *sm /*[33]*/) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint32_t carry = 0;
    for (int base = 0; base < nb; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const uint32_t x = i < nb ? __ldcg(totals + i) : 0u;
        uint32_t inc = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        __syncthreads();
        if (lane == 31) sm[w] = inc;
        __syncthreads();
        uint32_t wbase = 0;
        for (int i2 = 0; i2 < w; i2++) wbase += sm[i2];
        if (i < nb) bases[i] = carry + wbase + inc - x;
        uint32_t tot = 0;
        for (int i2 = 0; i2 < nw; i2++) tot += sm[i2];
        carry += tot;
    }
}

template <int D>
__global__ void __launch_bounds__(BIN_THREADS) k_bin(const float *__restrict__ Y, int first, int n,
                                                     const GridParams *__restrict__ gpp, uint32_t *__restrict__ keys_two_pass,
                                                     uint32_t *__restrict__ keys_one_pass, float *__restrict__ ubuf_two_pass,
                                                     float *__restrict__ ubuf_one_pass, uint32_t *__restrict__ totals,
                                                     uint32_t *__restrict__ bases, uint32_t *__restrict__ state, int tiles,
                                                     unsigned int *__restrict__ ticket) {
    pdl_prologue();
    __shared__ uint32_t cnt[2 << SORT_MAX_BITS];       // digit 0 | digit 1
    __shared__ GridParams gps;
    __shared__ uint32_t scan_sm[33];
    for (int i = threadIdx.x; i < (int) (sizeof(GridParams) / 4); i += blockDim.x)
        reinterpret_cast<int *>(&gps)[i] = reinterpret_cast<const int *>(gpp)[i];
    __syncthreads();
    const GridParams &gp = gps;
    if (!gp.ok) return;
    const bool one = gp.sort_passes == 1;
    uint32_t *keys = one ? keys_one_pass : keys_two_pass;      // whatever feeds the first sweep that really runs
    float *ubuf = one ? ubuf_one_pass : ubuf_two_pass;
    const int bits = gp.sort_bits, nb = 1 << bits;
    const uint32_t mask = (uint32_t) nb - 1;
    constexpr int NBMAX = 1 << SORT_MAX_BITS;
    uint32_t *cnt1 = cnt + NBMAX;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) {
        cnt[i] = 0; cnt1[i] = 0;
        state[((size_t) blockIdx.x) * NBMAX + i] = 0;                          // look-back words of this tile, pass 0
        state[((size_t) tiles + blockIdx.x) * NBMAX + i] = 0;                  // ... pass 1
    }
    __syncthreads();
    const int base = blockIdx.x * SORT_TILE;
#pragma unroll
    for (int r = 0; r < SORT_TILE / BIN_THREADS; r++) {
        const int k = base + r * BIN_THREADS + threadIdx.x;
        if (k < n) {
            uint32_t key;
            if (D == 2) {
                const float2 y = reinterpret_cast<const float2 *>(Y)[first + k];
                float2 u;
                const int bx = box_of<true>(y.x, gp, u.x);
                const int by = box_of<true>(y.y, gp, u.y);
                key = ((uint32_t) by << gp.xbits) | (uint32_t) bx;
                reinterpret_cast<float2 *>(ubuf)[k] = u;
            } else {
                float u;
                key = (uint32_t) box_of<false>(Y[first + k], gp, u);
                ubuf[k] = u;
            }
            keys[k] = key;
            atomicAdd(&cnt[key & mask], 1u);
            if (!one) atomicAdd(&cnt1[(key >> bits) & mask], 1u);
        }
    }
    __syncthreads();
    // one-pass layout: this histogram (whole key) is the one the single sweep uses -> pass-1 slot
    if (one) tile_hist_flush(cnt, nb, nullptr, tiles, totals + NBMAX);
    else {
        tile_hist_flush(cnt, nb, nullptr, tiles, totals);
        tile_hist_flush(cnt1, nb, nullptr, tiles, totals + NBMAX);
    }
    __threadfence();
    __syncthreads();                              // all of this CTA's atomics are performed before thread 0 takes the ticket
    if (last_block_done(ticket)) {                // digit totals -> exclusive digit bases, both passes
        if (!one) block_excl_scan(totals, bases, nb, scan_sm);
        __syncthreads();
        block_excl_scan(totals + NBMAX, bases + NBMAX, nb, scan_sm);
    }
}

// One pass of the per-iteration sort; see the section comment.  vals_in == nullptr or first executed pass: values are the
// identity (+ val_base).  u_in/u_out: `dims` floats per element that travel with it.
__global__ void __launch_bounds__(SWEEP_THREADS) k_radix_sweep(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                                                               uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, int n,
                                                               int pass, const uint32_t *__restrict__ bases, uint32_t *state, int tiles,
                                                               unsigned int *__restrict__ ticket, uint32_t val_base,
                                                               const GridParams *__restrict__ gpp, const float *__restrict__ u_in,
                                                               float *__restrict__ u_out, int dims) {
    pdl_prologue();
    if (!gpp->ok) return;
    const bool one = gpp->sort_passes == 1;
    if (one && pass == 0) return;
    if (one || pass == 0) vals_in = nullptr;
    extern __shared__ uint32_t smem[];
    __shared__ int tile_s;
    const int bits = gpp->sort_bits, shift = one ? 0 : pass * bits;
    const int nb = 1 << bits;
    constexpr int NW = SWEEP_THREADS / 32, NBMAX = 1 << SORT_MAX_BITS;
    uint32_t *gbase = smem;                                          // [nb]
    uint16_t *cnt = reinterpret_cast<uint16_t *>(smem + nb);         // [NW][nb]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) tile_s = (int) atomicAdd(ticket, 1u);
    for (int i = threadIdx.x; i < NW * nb; i += SWEEP_THREADS) cnt[i] = 0;
    __syncthreads();
    const int tile = tile_s;
    const uint32_t mask = (uint32_t) nb - 1;
    const int base = tile * SORT_TILE + w * (SWEEP_IPT * 32);
    uint32_t key[SWEEP_IPT], rank[SWEEP_IPT];
    uint16_t *mycnt = cnt + w * nb;
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < SWEEP_IPT; r++) {
        const int i = base + r * 32 + lane;
        key[r] = i < n ? keys_in[i] : 0xffffffffu;
    }
#pragma unroll
    for (int r = 0; r < SWEEP_IPT; r++) {
        const int i = base + r * 32 + lane;
        const bool valid = i < n;
        const uint32_t d = valid ? ((key[r] >> shift) & mask) : (uint32_t) nb;   // sentinel digit groups the tail
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (lane == leader && valid) {
            old = mycnt[d];
            mycnt[d] = (uint16_t) (old + __popc(peers));
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[r] = old + __popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();
    // per digit: exclusive prefix over this tile's warps, publish the tile total (AGGREGATE), then look back over the
    // earlier tiles -- sixteen words in flight -- adding aggregates until a tile that already knows its INCLUSIVE prefix
    // is met, and publish this tile's inclusive prefix in turn (decoupled look-back: the walk is bounded by the number of
    // tiles in flight, not by the number of tiles).
    uint32_t *mystate = state + ((size_t) pass * tiles + tile) * NBMAX;
    const uint32_t *prev = state + (size_t) pass * tiles * NBMAX;
    for (int d = threadIdx.x; d < nb; d += SWEEP_THREADS) {
        uint32_t run = 0;
#pragma unroll
        for (int ww = 0; ww < NW; ww++) {
            const uint32_t t = cnt[ww * nb + d];
            cnt[ww * nb + d] = (uint16_t) run;
            run += t;
        }
        if (tile > 0) *reinterpret_cast<volatile uint32_t *>(mystate + d) = SWEEP_AGGREGATE | run;
        uint32_t excl = 0;
        int t = tile - 1;
        while (t >= 0) {
            uint32_t v[16];
#pragma unroll
            for (int j = 0; j < 16; j++)           // tiles before the first count as "inclusive prefix 0"
                v[j] = t - j >= 0 ? *reinterpret_cast<const volatile uint32_t *>(prev + (size_t) (t - j) * NBMAX + d) : SWEEP_INCLUSIVE;
            bool finished = false, stopped = false;
            int used = 0;
#pragma unroll
            for (int j = 0; j < 16; j++) {         // consume in order: stop at a word that is not published yet (re-read from there)
                const uint32_t status = v[j] >> 30;
                if (!finished && !stopped) {
                    if (status == 0) stopped = true;
                    else { excl += v[j] & SWEEP_COUNT_MASK; used++; finished = status == 2; }
                }
            }
            if (finished) break;
            t -= used;
        }
        *reinterpret_cast<volatile uint32_t *>(mystate + d) = SWEEP_INCLUSIVE | (excl + run);
        gbase[d] = bases[(size_t) pass * NBMAX + d] + excl;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SWEEP_IPT; r++) {
        const int i = base + r * 32 + lane;
        if (i < n) {
            const uint32_t d = (key[r] >> shift) & mask;
            const uint32_t pos = gbase[d] + mycnt[d] + rank[r];
            keys_out[pos] = key[r];
            vals_out[pos] = vals_in ? vals_in[i] : (val_base + (uint32_t) i);
            if (dims == 2) reinterpret_cast<float2 *>(u_out)[pos] = reinterpret_cast<const float2 *>(u_in)[i];
            else u_out[pos] = u_in[i];
        }
    }
}

// per-tile histogram of digit `pass` of already computed keys (point re-ordering only: the per-iteration box sort gets
// its histograms from k_bin and k_radix_scatter)
__global__ void __launch_bounds__(SORT_THREADS) k_radix_hist(const uint32_t *__restrict__ keys, int n, int pass,
                                                             uint32_t *__restrict__ hist, int tiles, uint32_t *__restrict__ totals,
                                                             const GridParams *__restrict__ gpp) {
    if (!gpp->ok || gpp->sort_passes == 1) return;
    __shared__ uint32_t cnt[1 << SORT_MAX_BITS];
    const int bits = gpp->sort_bits, shift = pass * bits;
    const int nb = 1 << bits;
    for (int i = threadIdx.x; i < nb; i += SORT_THREADS) cnt[i] = 0;
    __syncthreads();
    const int base = blockIdx.x * SORT_TILE;
    const uint32_t mask = (uint32_t) nb - 1;
#pragma unroll
    for (int r = 0; r < SORT_IPT; r++) {
        const int i = base + r * SORT_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&cnt[(keys[i] >> shift) & mask], 1u);
    }
    __syncthreads();
    tile_hist_flush(cnt, nb, hist, tiles, totals);
}

// One CTA per digit value d: hist[d][t] <- (sum of totals[d' < d]) + (exclusive prefix over tiles t' < t), in place.
__global__ void __launch_bounds__(256) k_radix_offsets(uint32_t *__restrict__ hist, int tiles, const uint32_t *__restrict__ totals,
                                                       int pass, const GridParams *__restrict__ gpp) {
    if (!gpp->ok) return;
    const bool one = gpp->sort_passes == 1;
    if (one && pass == 0) return;
    const int nb = 1 << gpp->sort_bits;
    const int d = blockIdx.x;
    if (d >= nb) return;
    __shared__ uint32_t wsum[8];
    __shared__ uint32_t carry_s;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    // base = sum of totals of smaller digits
    uint32_t acc = 0;
    for (int i = threadIdx.x; i < d; i += 256) acc += totals[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) wsum[w] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int i = 0; i < 8; i++) t += wsum[i];
        carry_s = t;
    }
    __syncthreads();
    uint32_t *rowp = hist + (size_t) d * tiles;
    for (int base = 0; base < tiles; base += 256) {
        const int i = base + threadIdx.x;
        const uint32_t x = i < tiles ? rowp[i] : 0u;
        uint32_t inc = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        __syncthreads();            // wsum / carry_s reads of the previous round are done
        if (lane == 31) wsum[w] = inc;
        __syncthreads();
        uint32_t wbase = 0;
        for (int i2 = 0; i2 < w; i2++) wbase += wsum[i2];
        const uint32_t carry = carry_s;
        if (i < tiles) rowp[i] = carry + wbase + inc - x;
        __syncthreads();
        if (threadIdx.x == 255) carry_s = carry + wbase + inc;
        __syncthreads();
    }
}

// Stable scatter by digit `pass` from precomputed offsets.  vals_in == nullptr: values are the identity (+ val_base).
__global__ void __launch_bounds__(SORT_THREADS) k_radix_scatter(const uint32_t *__restrict__ keys_in,
                                                                const uint32_t *__restrict__ vals_in,
                                                                uint32_t *__restrict__ keys_out,
                                                                uint32_t *__restrict__ vals_out, int n, int pass,
                                                                const uint32_t *__restrict__ hist, int tiles,
                                                                uint32_t val_base, const GridParams *__restrict__ gpp) {
    if (!gpp->ok) return;
    const bool one = gpp->sort_passes == 1;
    if (one && pass == 0) return;
    if (one || pass == 0) vals_in = nullptr;              // first scatter of the sort: values are the identity (+ val_base)
    extern __shared__ uint32_t smem[];
    const int bits = gpp->sort_bits, shift = one ? 0 : pass * bits;
    const int nb = 1 << bits;
    constexpr int NW = SORT_THREADS / 32;
    uint32_t *gbase = smem;                                          // [nb]
    uint16_t *cnt = reinterpret_cast<uint16_t *>(smem + nb);         // [NW][nb]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < nb; i += SORT_THREADS) gbase[i] = hist[(size_t) i * tiles + blockIdx.x];
    for (int i = threadIdx.x; i < NW * nb; i += SORT_THREADS) cnt[i] = 0;
    __syncthreads();
    const uint32_t mask = (uint32_t) nb - 1;
    const int base = blockIdx.x * SORT_TILE + w * (SORT_IPT * 32);
    uint32_t key[SORT_IPT], rank[SORT_IPT];
    uint16_t *mycnt = cnt + w * nb;
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < SORT_IPT; r++) {
        const int i = base + r * 32 + lane;
        const bool valid = i < n;
        key[r] = valid ? keys_in[i] : 0xffffffffu;
        const uint32_t d = valid ? ((key[r] >> shift) & mask) : (uint32_t) nb;   // sentinel digit groups the tail
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (lane == leader && valid) {
            old = mycnt[d];
            mycnt[d] = (uint16_t) (old + __popc(peers));
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[r] = old + __popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();
    // exclusive prefix over warps, per digit
    for (int d = threadIdx.x; d < nb; d += SORT_THREADS) {
        uint32_t run = 0;
#pragma unroll
        for (int ww = 0; ww < NW; ww++) {
            const uint32_t t = cnt[ww * nb + d];
            cnt[ww * nb + d] = (uint16_t) run;
            run += t;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SORT_IPT; r++) {
        const int i = base + r * 32 + lane;
        if (i < n) {
            const uint32_t d = (key[r] >> shift) & mask;
            const uint32_t pos = gbase[d] + mycnt[d] + rank[r];
            keys_out[pos] = key[r];
            vals_out[pos] = vals_in ? vals_in[i] : (val_base + (uint32_t) i);
        }
    }
}

// --------------------------------------------------------------------------------------------- spread --

template <int P>
__host__ __device__ __forceinline__ float lagrange1(const GridParams &gp, int p, int j, float u) {
    float v = gp.inv_den[j];
    if (P > 0) {
#pragma unroll
        for (int k = 0; k < P; k++) if (k != j) v *= (u - gp.s[k]);
    } else {
        for (int k = 0; k < p; k++) if (k != j) v *= (u - gp.s[k]);
    }
    return v;
}

// Spread (nbodyfft.cpp:123-147), deterministic and free of float atomics.  The box-sorted points are cut into fixed
// chunks of CHUNK consecutive points; ONE THREAD per chunk walks its points in order with all p^D node accumulators
// (L, L*bx, L*by, L*|b|^2) in registers (phase 2), then the chunks of a CTA (SP2_THREADS chunks = SP2_POINTS points) are
// stitched together in shared memory (phase 3):
//   * a box that lies entirely inside a chunk is finished in phase 2 and written straight to the (pre-zeroed) grid;
//   * the first / last segment of a chunk that continues from / into the neighbouring chunk is parked in shared memory
//     (H / T partial of the chunk); in phase 3 the thread of the chunk where such a box starts adds the partials of the
//     following chunks in chunk order and writes the box -- no round trip through global memory;
//   * only a box that crosses a CTA boundary leaves a partial in a global slot (two per CTA) and is appended to a work
//     list by the CTA that holds its head; k_spread_combine adds those few slots in CTA order.
//   * box boundaries are detected on the way (one look at the key before and after the chunk): box_range[] = [first, end)
//     of every non-empty box -- which only the combine step needs -- is a by-product, not a separate pass.
// Everything else in the grid is zero (empty box), so nothing ever iterates over the grid's nodes.  Every sum has a fixed
// order given the sorted order, which the stable sort makes unique: bitwise repeatable.
//   2-D: node = a*p + b, a = y node, b = x node (grid row = y node, column = x node).  1-D: (L, L*b, L*b^2, 0).
// A chunk's keys are 64 contiguous bytes and its in-box coordinates 128 (2-D): the walk reads them straight from global
// memory -- one cache line per thread, L1 hits after the first touch -- so shared memory holds nothing but the partials.
// Written as phase functions so that tests/tools/spread_emul.cu can run the very same code on the host (all threads of a
// block through phase 2, then phase 3) and check it against a direct spread.
template <int D>
__host__ __device__ __forceinline__ int key_to_box(uint32_t key, const GridParams &gp) {
    return D == 2 ? (int) (key >> gp.xbits) * gp.B + (int) (key & ((1u << gp.xbits) - 1u)) : (int) key;
}

// Spread results.  delta and wbb are kept in BOX UNITS (offsets / box width): every component is then O(w1) whatever the
// embedding's scale, which is what makes packing two real grids into one fp32 complex transform safe (a 1e-5-scale grid
// packed beside an O(1) grid would lose 5 digits in the separation).
//   2-D: one float4 per node, chg[(y node) * G + (x node)] = (w1, delta_x, delta_y, wbb): the dense G x G grid that
//        k_conv_rows_fwd reads row by row (and that sharded runs all-reduce as it is);
//   1-D: two packed complex lines of length M, plane 0 = (w1, delta), plane 1 = (wbb, 0), the FFT input itself.
template <int D>
__host__ __device__ __forceinline__ void store_node(void *__restrict__ grid, size_t stride, size_t off, float4 v) {
    if (D == 2) {
        reinterpret_cast<float4 *>(grid)[off] = v;
    } else {
        float2 *dst = reinterpret_cast<float2 *>(grid);
        dst[off] = make_float2(v.x, v.y);
        dst[stride + off] = make_float2(v.z, v.w);
    }
}

// offset of (box, node) in the spread grid
template <int D>
__host__ __device__ __forceinline__ size_t node_offset(int box, int node, const GridParams &gp, int p) {
    if (D == 2) {
        const int by = box / gp.B, bx = box - by * gp.B;
        const int a = node / p, b = node - a * p;
        return (size_t) (by * p + a) * (size_t) gp.G + (size_t) (bx * p + b);
    }
    return (size_t) box * p + node;
}

constexpr int SP2_THREADS = 128;                // chunks per block
constexpr int SP2_POINTS = SP2_THREADS * CHUNK;

// all nodes of one box segment: to the grid (finished box; one box -> base offset computed once) or to a partial slot.
// P > 0: compile-time node count (registers); P == 0: run-time p (nterms > 4 in 2-D, > 5 in 1-D; local-memory accumulators)
template <int D, int P, int MAXN>
__host__ __device__ __forceinline__ void spread2_flush(const float4 (&acc)[MAXN], int p, bool to_grid, int box, const GridParams &gp,
                                                       void *__restrict__ grid, size_t stride, float4 *__restrict__ slot) {
    const int nodes = D == 2 ? p * p : p;
    if (to_grid) {
        const size_t base = node_offset<D>(box, 0, gp, p);
        const size_t rs = (size_t) gp.G;
#pragma unroll
        for (int j = 0; j < MAXN; j++) {
            if (P == 0 && j >= nodes) break;
            const size_t off = D == 2 ? base + (size_t) (j / p) * rs + (size_t) (j % p) : base + (size_t) j;
            store_node<D>(grid, stride, off, acc[j]);
        }
    } else {
#pragma unroll
        for (int j = 0; j < MAXN; j++) {
            if (P == 0 && j >= nodes) break;
            slot[j] = acc[j];
        }
    }
}

// per-chunk bookkeeping between phase 2 and phase 3
constexpr int SP2_HVALID = 1;     // the chunk's first segment continues a box from the previous chunk: partial H
constexpr int SP2_TVALID = 2;     // the chunk's last segment continues into the next chunk (and is not the H segment): partial T
constexpr int SP2_THROUGH = 4;    // the H segment is the whole chunk AND continues into the next chunk
struct Sp2Meta {
    int first_box[SP2_THREADS], last_box[SP2_THREADS];
    int flags[SP2_THREADS];
};

// phase 2: thread t walks chunk c = blk*SP2_THREADS + t.  part = this CTA's partials, [chunk][H|T][node].
template <int D, int P>
__host__ __device__ __forceinline__ void spread2_chunk(int t, int blk, const float *__restrict__ sorted_u, const uint32_t *__restrict__ skeys,
                                                       int n, const GridParams &gp, float4 *__restrict__ part, Sp2Meta &meta,
                                                       void *__restrict__ grid, uint2 *__restrict__ box_range, int chunk = CHUNK) {
    constexpr int PP = P > 0 ? P : PMAX;
    constexpr int MAXN = D == 2 ? PP * PP : PP;
    const int p = P > 0 ? P : gp.p;
    const int nodes = D == 2 ? p * p : p;
    const int c = blk * SP2_THREADS + t;
    const int kb = c * chunk;
    meta.flags[t] = 0;
    if (kb >= n) return;
    const int ke = kb + chunk < n ? kb + chunk : n;
    const size_t stride = (size_t) gp.M;              // 1-D: plane 0 -> plane 1
    float4 *myH = part + (size_t) t * 2 * nodes, *myT = myH + nodes;
    const uint32_t *kp = skeys + kb;
    float4 acc[MAXN];
#pragma unroll
    for (int j = 0; j < MAXN; j++) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    int cur = key_to_box<D>(kp[0], gp);
    const int prev = kb > 0 ? key_to_box<D>(skeys[kb - 1], gp) : -1;
    const bool head_cont = prev == cur;               // the first segment continues a box from the previous chunk
    bool started = !head_cont;                        // the segment being accumulated started inside this chunk
    if (started) box_range[cur].x = (uint32_t) kb;
    meta.first_box[t] = cur;
    // software pipeline: the next point's key and coordinates are requested before this point's ~100 instructions
    uint32_t key_next = kp[0];
    float2 u_next = make_float2(0.f, 0.f);
    if (D == 2) u_next = reinterpret_cast<const float2 *>(sorted_u)[kb]; else u_next.x = sorted_u[kb];
    for (int k = kb; k < ke; k++) {
        const uint32_t key_k = key_next;
        const float2 u_k = u_next;
        if (k + 1 < ke) {
            key_next = kp[k + 1 - kb];
            if (D == 2) u_next = reinterpret_cast<const float2 *>(sorted_u)[k + 1]; else u_next.x = sorted_u[k + 1];
        }
        const int box = key_to_box<D>(key_k, gp);
        if (box != cur) {
            // segment of `cur` ended inside the chunk: finished box unless it started before the chunk (-> H partial)
            spread2_flush<D, P, MAXN>(acc, p, started, cur, gp, grid, stride, myH);
            box_range[cur].y = (uint32_t) k;              // [first, end) of every non-empty box: written where the change is seen,
            box_range[box].x = (uint32_t) k;              // so empty boxes cost nothing (their entries are never read)
#pragma unroll
            for (int j = 0; j < MAXN; j++) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            cur = box;
            started = true;
        }
        if (D == 2) {
            const float2 u = u_k;
            float Lx[PP], Ly[PP], ox[PP], oy[PP];
#pragma unroll
            for (int j = 0; j < PP; j++) {
                if (P == 0 && j >= p) break;
                Lx[j] = lagrange1<P>(gp, p, j, u.x); Ly[j] = lagrange1<P>(gp, p, j, u.y);
                ox[j] = u.x - gp.s[j]; oy[j] = u.y - gp.s[j];          // offsets in BOX UNITS (x bw later)
            }
#pragma unroll
            for (int a = 0; a < PP; a++) {
                if (P == 0 && a >= p) break;
#pragma unroll
                for (int b = 0; b < PP; b++) {
                    if (P == 0 && b >= p) break;
                    const float L = Ly[a] * Lx[b];
                    float4 &q = acc[a * p + b];
                    q.x += L;
                    q.y += L * ox[b];
                    q.z += L * oy[a];
                    q.w += L * (ox[b] * ox[b] + oy[a] * oy[a]);
                }
            }
        } else {
            const float u = u_k.x;
#pragma unroll
            for (int a = 0; a < PP; a++) {
                if (P == 0 && a >= p) break;
                const float L = lagrange1<P>(gp, p, a, u);
                const float o = u - gp.s[a];
                float4 &q = acc[a];
                q.x += L;
                q.y += L * o;
                q.z += L * o * o;
            }
        }
    }
    // last segment
    const int next = ke < n ? key_to_box<D>(skeys[ke], gp) : gp.nb;
    const bool ends = next != cur;
    if (ends) box_range[cur].y = (uint32_t) ke;
    meta.last_box[t] = cur;
    int fl = head_cont ? SP2_HVALID : 0;
    if (started) {                                    // started in this chunk: finished, or the head (T) of a longer box
        spread2_flush<D, P, MAXN>(acc, p, ends, cur, gp, grid, stride, myT);
        if (!ends) fl |= SP2_TVALID;
    } else {                                          // the whole chunk is one segment of a box that started earlier (H)
        spread2_flush<D, P, MAXN>(acc, p, false, cur, gp, grid, stride, myH);
        if (!ends) fl |= SP2_THROUGH;
    }
    meta.flags[t] = fl;
}

// phase 3: thread t finishes the boxes whose in-CTA run of partials starts at chunk t (T of chunk t, or -- thread 0 -- the
// H run the CTA inherits from its predecessor).  A run that stays inside the CTA is final -> grid; one that crosses a CTA
// boundary goes to the CTA's global slot (0: continues from the previous CTA, 1: continues into the next) and, if its head
// is here, onto the work list.
template <int D, int P>
__host__ __device__ __forceinline__ void spread2_stitch(int t, int blk, int n, const GridParams &gp, const float4 *__restrict__ part,
                                                        const Sp2Meta &meta, void *__restrict__ grid, float4 *__restrict__ cslots,
                                                        uint32_t *__restrict__ work, int chunk = CHUNK) {
    constexpr int PP = P > 0 ? P : PMAX;
    constexpr int MAXN = D == 2 ? PP * PP : PP;
    const int p = P > 0 ? P : gp.p;
    const int nodes = D == 2 ? p * p : p;
    const int nch = min(SP2_THREADS, (n - blk * SP2_THREADS * chunk + chunk - 1) / chunk);       // chunks of this CTA
    if (t >= nch) return;
    const size_t stride = (size_t) gp.M;
    float4 *myc = cslots + (size_t) blk * 2 * nodes;
    for (int which = (t == 0 ? 0 : 1); which < 2; which++) {
        // which == 0 (thread 0 only): the run that starts with H[0]; which == 1: the run that starts with T[t]
        const int fl = meta.flags[t];
        if (which == 0 ? !(fl & SP2_HVALID) : !(fl & SP2_TVALID)) continue;
        float4 acc[MAXN];
#pragma unroll
        for (int j = 0; j < MAXN; j++) {
            if (P == 0 && j >= nodes) break;
            acc[j] = part[((size_t) t * 2 + which) * nodes + j];
        }
        // follow the H partials of the next chunks while the box goes on
        bool open = which == 0 ? (fl & SP2_THROUGH) != 0 : true;      // the box continues beyond the chunk just added
        int tt = t;
        while (open && tt + 1 < nch) {
            tt++;
            const float4 *h = part + (size_t) tt * 2 * nodes;
#pragma unroll
            for (int j = 0; j < MAXN; j++) {
                if (P == 0 && j >= nodes) break;
                const float4 v = h[j];
                acc[j].x += v.x; acc[j].y += v.y; acc[j].z += v.z; acc[j].w += v.w;
            }
            open = (meta.flags[tt] & SP2_THROUGH) != 0;
        }
        const int box = which == 0 ? meta.first_box[0] : meta.last_box[t];
        if (which == 1 && !open) {
            spread2_flush<D, P, MAXN>(acc, p, true, box, gp, grid, stride, nullptr);      // started and ended inside the CTA
        } else {
            spread2_flush<D, P, MAXN>(acc, p, false, box, gp, grid, stride, myc + (size_t) which * nodes);
            if (which == 1) {                         // head of a box that crosses into the next CTA
                const uint32_t e = FK_ATOMIC_ADD(work, 1u);
                work[1 + e] = (uint32_t) box;
            }
        }
    }
}

// dynamic shared memory of k_spread_chunks: the partials of the CTA's chunks (compile-time node counts only; the run-time-p
// fallback keeps them in a global scratch area)
template <int D, int P>
__host__ __device__ constexpr size_t spread_smem_bytes() {
    return (size_t) SP2_THREADS * 2 * (P > 0 ? (D == 2 ? P * P : P) : 0) * sizeof(float4);
}

#ifdef __CUDACC__
template <int D, int P>
__global__ void __launch_bounds__(SP2_THREADS) k_spread_chunks(const float *__restrict__ sorted_u, const uint32_t *__restrict__ skeys, int n,
                                                               const GridParams *__restrict__ gpp, float4 *__restrict__ cslots,
                                                               float4 *__restrict__ gpart, void *__restrict__ grid,
                                                               uint2 *__restrict__ box_range, uint32_t *__restrict__ work, int chunk) {
    pdl_prologue();
    extern __shared__ __align__(16) unsigned char sp_raw[];
    __shared__ GridParams gps;
    __shared__ Sp2Meta meta;
    for (int i = threadIdx.x; i < (int) (sizeof(GridParams) / 4); i += blockDim.x)
        reinterpret_cast<int *>(&gps)[i] = reinterpret_cast<const int *>(gpp)[i];
    __syncthreads();
    if (!gps.ok) return;
    const int p = P > 0 ? P : gps.p;
    const int nodes = D == 2 ? p * p : p;
    float4 *part = P > 0 ? reinterpret_cast<float4 *>(sp_raw) : gpart + (size_t) blockIdx.x * SP2_THREADS * 2 * nodes;
    spread2_chunk<D, P>(threadIdx.x, blockIdx.x, sorted_u, skeys, n, gps, part, meta, grid, box_range, chunk);
    __syncthreads();                                  // (block-scope barrier also orders the global-memory partials of P == 0)
    spread2_stitch<D, P>(threadIdx.x, blockIdx.x, n, gps, part, meta, grid, cslots, work, chunk);
}

// Combine the boxes that cross CTA boundaries of k_spread_chunks: work[0] = number of listed boxes, work[1..] = the boxes
// (in arrival order, which does not matter: every box is summed on its own).  One thread per (box, node) adds the CTA slots
// in CTA order; a box that spans COMBINE_COOP CTAs or more is summed by the whole warp (lane-strided + fixed shuffle tree).
// The summation order depends only on the number of CTAs the box spans: deterministic.  Persistent grid.
constexpr int COMBINE_COOP = 32;
#endif

// Host-callable core: lane `lane` of `lanes` of one (box, node); CTA b0 holds the box's head in slot 1, later CTAs use slot 0
template <int D>
__host__ __device__ __forceinline__ float4 combine_node_lane(const float4 *__restrict__ cslots, int b0, int b1, int nodes, int node, int lane, int lanes) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = b0 + lane; b <= b1; b += lanes) {
        const float4 v = cslots[((size_t) b * 2 + (b == b0 ? 1 : 0)) * nodes + node];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    return acc;
}

#ifdef __CUDACC__
template <int D>
__global__ void __launch_bounds__(256) k_spread_combine(const float4 *__restrict__ cslots, const uint2 *__restrict__ box_range,
                                                        const GridParams *__restrict__ gpp, const uint32_t *__restrict__ work,
                                                        void *__restrict__ grid, unsigned int *__restrict__ ticket, PeerComm pc, int p2p,
                                                        int chunk) {
    pdl_prologue();
    const GridParams &gp = *gpp;
    const int p = gp.p, nodes = D == 2 ? p * p : p;
    const int lane = threadIdx.x & 31;
    const int ntask = gp.ok ? (int) work[0] * nodes : 0;
    const int stride = gridDim.x * blockDim.x;
    for (int t0 = blockIdx.x * blockDim.x + threadIdx.x - lane; t0 < ntask; t0 += stride) {
        const int task = t0 + lane;
        int box = 0, node = 0, b0 = 0, b1 = -1;
        if (task < ntask) {
            const int e = task / nodes;
            node = task - e * nodes;
            box = (int) work[1 + e];
            const uint2 r = box_range[box];
            b0 = (int) r.x / (SP2_THREADS * chunk); b1 = ((int) r.y - 1) / (SP2_THREADS * chunk);
        }
        const bool coop = b1 - b0 + 1 >= COMBINE_COOP;
        if (task < ntask && !coop)
            store_node<D>(grid, (size_t) gp.M, node_offset<D>(box, node, gp, p), combine_node_lane<D>(cslots, b0, b1, nodes, node, 0, 1));
        // very long boxes (most points in one box): the whole warp takes them one after the other
        unsigned todo = __ballot_sync(0xffffffffu, coop);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int bbox = __shfl_sync(0xffffffffu, box, src), bnode = __shfl_sync(0xffffffffu, node, src);
            const int c0 = __shfl_sync(0xffffffffu, b0, src), c1 = __shfl_sync(0xffffffffu, b1, src);
            float4 acc = combine_node_lane<D>(cslots, c0, c1, nodes, bnode, lane, 32);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
                acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
                acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
                acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
            }
            if (lane == 0) store_node<D>(grid, (size_t) gp.M, node_offset<D>(bbox, bnode, gp, p), acc);
        }
    }
    // sharded, peer fabric: my partial grid is complete -- tell the peers (they add the partials inside k_conv_rows_fwd's
    // loads / k_grid_sum_1d).  Sent whether or not the grid is valid: a flag per iteration keeps the ranks in step.
    if (p2p) peer_signal_last(ticket, pc, FLAG_GRID, (int) (blockIdx.x * blockDim.x) < ntask ? 1 : 0);
}
#endif

// ------------------------------------------------------------------------- kernel samples (1-D embeddings) --
// 2-D embeddings sample and transform their kernels inside fitsne_conv.cuh (k_kspec_rows / k_kspec_cols).  1-D: real
// kernels on the wrap-around node-offset line, offsets d in (-G, G) stored at index d mod M (the reference's 2G
// circulant embedding, nbodyfft.cpp:280-300, with M >= 2G), packed two per complex line:
//   plane 2 = (Ksq, Kb)   plane 3 = (Kgrad, 0)
//   Ksq=(1+r2/df)^-(df+1)   Kgrad = (R/bw)*Ksq  (box units)   Kb=(1+r2/df)^-df      (tsne.cpp:69-94)
// Values carry the 1/M inverse-FFT normalisation (nbodyfft.cpp:427-430).
__global__ void __launch_bounds__(256) k_gen_kernels_1d(const GridParams *__restrict__ gpp, double df, float2 *__restrict__ planes) {
    pdl_prologue();
    const GridParams &gp = *gpp;
    if (!gp.ok) return;
    const int M = gp.M, G = gp.G;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= M) return;
    const int dc = c < G ? c : (c > M - G ? c - M : 0);
    float2 k1 = make_float2(0.f, 0.f), k2 = k1;
    if (c < G || c > M - G) {
        const double r2 = gp.h * gp.h * (double) dc * (double) dc;
        double kb, ksq;
        if (df == 1.0) { kb = 1.0 / (1.0 + r2); ksq = kb * kb; }
        else { const double t = 1.0 + r2 / df; kb = pow(t, -df); ksq = pow(t, -(df + 1.0)); }
        kb *= gp.inv_norm; ksq *= gp.inv_norm;
        k1 = make_float2((float) ksq, (float) kb);
        k2 = make_float2((float) ((double) dc / (double) gp.p * ksq), 0.f);   // gradient kernel in box units: R / bw = offset / p
    }
    planes[2 * (size_t) M + c] = k1;
    planes[3 * (size_t) M + c] = k2;
}

// ------------------------------------------------------------------------------ Hadamard + sum_Q terms --
// A packed spectrum Z = FFT(A + iB) of two REAL arrays separates as  A^[k] = (Z[k] + conj(Z[-k]))/2,
// B^[k] = (Z[k] - conj(Z[-k]))/(2i).
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
__device__ __forceinline__ double re_conj_mul(float2 a, float2 b) { return (double) a.x * (double) b.x + (double) a.y * (double) b.y; }
// split Z[k], Z[-k] of a packed pair into the spectra of its real part (A) and imaginary part (B) at k
__device__ __forceinline__ void unpack_pair(float2 zk, float2 zm, float2 &A, float2 &B) {
    A = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
    B = make_float2(0.5f * (zk.y + zm.y), -0.5f * (zk.x - zm.x));
}

// 1-D embeddings.  Input: the length-M spectra of the four packed lines; one thread owns the frequency pair (k, -k), so
// everything is done in place:
//   plane 0 <- V1 = v1^ + i*B^      v1 = Ksq*w1,  B = Kgrad*w1 - Ksq*delta     (nbodyfft.cpp:410-420)
// and, by Parseval in fp64, the sum_Q terms
//   df==1: <w1,Kb*w1> + 2<wbb,v1> + 4<delta,Kgrad*w1> - 2<delta,Ksq*delta>   (tsne.cpp:809-818)
//   df!=1: <w1,Kb*w1>                                                        (tsne.cpp:700-706)
__global__ void __launch_bounds__(256) k_hadamard_1d(float2 *__restrict__ planes, const GridParams *__restrict__ gpp,
                                                     int df_is_one, double *__restrict__ zpartial, int N, Scalars *__restrict__ sc,
                                                     unsigned int *__restrict__ ticket) {
    pdl_prologue();
    __shared__ double sm[32];
    const GridParams &gp = *gpp;
    if (!gp.ok) return;
    const int M = gp.M;
    float2 *Z1 = planes, *Z2 = planes + M;
    const float2 *K1 = planes + 2 * (size_t) M, *K2 = planes + 3 * (size_t) M;
    const double bw2 = gp.bw * gp.bw;
    double zacc = 0;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < M; e += gridDim.x * blockDim.x) {
        const int em = (M - e) % M;
        if (em < e) continue;                       // the partner thread owns this pair
        const double wt = em == e ? 1.0 : 2.0;
        float2 w1, d1, wbb, unused, ksq, kb, kg1, kg2;
        unpack_pair(Z1[e], Z1[em], w1, d1);
        unpack_pair(Z2[e], Z2[em], wbb, unused);
        unpack_pair(K1[e], K1[em], ksq, kb);
        unpack_pair(K2[e], K2[em], kg1, kg2);
        const float2 v1 = cmul(ksq, w1);
        const float2 kgw1 = cmul(kg1, w1), ksd1 = cmul(ksq, d1);
        const float2 B1 = make_float2(kgw1.x - ksd1.x, kgw1.y - ksd1.y);
        // delta, wbb, Kgrad and B are in box units: the physical sum_Q terms carry bw^2
        double z = re_conj_mul(w1, cmul(kb, w1));
        double zb = 0;
        if (df_is_one) zb += 2.0 * re_conj_mul(wbb, v1) + 4.0 * re_conj_mul(d1, kgw1) - 2.0 * re_conj_mul(d1, ksd1);
        // V1 = v1 + i*B1 at k; at -k both spectra are conjugated (real arrays): V1[-k] = conj(v1) + i*conj(B1)
        Z1[e] = make_float2(v1.x - B1.y, v1.y + B1.x);
        if (em != e) Z1[em] = make_float2(v1.x + B1.y, -v1.y + B1.x);
        zacc += wt * (z + bw2 * zb);
    }
    const double r = block_sum(zacc, sm);
    if (threadIdx.x == 0) zpartial[blockIdx.x] = r;
    if (last_block_done(ticket)) {           // sum_Q = (sum of the partials, in index order) - N   (tsne.cpp:818)
        double s2 = 0;
        for (int i = threadIdx.x; i < (int) gridDim.x; i += blockDim.x) s2 += ld_partial(zpartial + i);
        const double tot = block_sum(s2, sm);
        if (threadIdx.x == 0) {
            const double Z = tot - (double) N;
            sc->Z = Z;
            sc->inv_Z = (float) (1.0 / Z);
        }
    }
}

// --------------------------------------------------------------------------------------------- gather --
// One thread per box-sorted point: F_rep/Z = (1/Z) sum_nodes L * (a_k * v1 + B_k) with a = y - X_node;
// written to the point's ORIGINAL index (frep[perm[k]]), so the update runs coalesced in point order.
// 2-D: `field` = pot, one float4 (v1, Bx, By, 0) per node of the G x G grid; 1-D: packed line (v1, B).
template <int D, int P>
__global__ void __launch_bounds__(256) k_gather(const float *__restrict__ sorted_u, const uint32_t *__restrict__ skeys,
                                                const uint32_t *__restrict__ perm, int n,
                                                const GridParams *__restrict__ gpp, const Scalars *__restrict__ sc,
                                                const void *__restrict__ field, float *__restrict__ frep) {
    pdl_prologue();
    __shared__ GridParams gps;
    for (int i = threadIdx.x; i < (int) (sizeof(GridParams) / 4); i += blockDim.x)
        reinterpret_cast<int *>(&gps)[i] = reinterpret_cast<const int *>(gpp)[i];
    __syncthreads();
    const GridParams &gp = gps;
    if (!gp.ok) return;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int p = P > 0 ? P : gp.p;
    const float bw = gp.bwf, inv_Z = sc->inv_Z;
    const uint32_t key = skeys[k];
    if (D == 2) {
        const int G = gp.G;
        const int by = (int) (key >> gp.xbits), bx = (int) (key & ((1u << gp.xbits) - 1u));
        const float2 u = reinterpret_cast<const float2 *>(sorted_u)[k];
        float Lx[P > 0 ? P : PMAX], ox[P > 0 ? P : PMAX];
#pragma unroll(P > 0 ? P : 1)
        for (int b = 0; b < (P > 0 ? P : PMAX); b++) {
            if (b < p) { Lx[b] = lagrange1<P>(gp, p, b, u.x); ox[b] = u.x - gp.s[b]; }
        }
        float fx = 0.f, fy = 0.f;
        const float4 *g0 = reinterpret_cast<const float4 *>(field) + (size_t) (by * p) * G + bx * p;
#pragma unroll(P > 0 ? P : 1)
        for (int a = 0; a < (P > 0 ? P : PMAX); a++) {
            if (a < p) {
                const float Ly = lagrange1<P>(gp, p, a, u.y);
                const float oy = u.y - gp.s[a];
                const float4 *row = g0 + (size_t) a * G;
#pragma unroll(P > 0 ? P : 1)
                for (int b = 0; b < (P > 0 ? P : PMAX); b++) {
                    if (b < p) {
                        const float L = Ly * Lx[b];
                        const float4 vb = __ldg(row + b);
                        fx += L * (ox[b] * vb.x + vb.y);
                        fy += L * (oy * vb.x + vb.z);
                    }
                }
            }
        }
        // offsets and B planes are in box units: one factor bw restores the physical force
        reinterpret_cast<float2 *>(frep)[perm[k]] = make_float2(fx * bw * inv_Z, fy * bw * inv_Z);
    } else {
        const float u = sorted_u[k];
        const float2 *g0 = reinterpret_cast<const float2 *>(field) + (size_t) key * p;                  // plane 0 = (v1, B)
        float f = 0.f;
        for (int a = 0; a < p; a++) {
            const float L = lagrange1<P>(gp, p, a, u);
            const float o = u - gp.s[a];
            const float2 vb = __ldg(g0 + a);
            f += L * (o * vb.x + vb.y);
        }
        frep[perm[k]] = f * bw * inv_Z;
    }
}

// ------------------------------------------------------------------- attractive term + optimiser step --
// k_attract: LPR lanes cooperate on one CSR row: attr_i = sum_j p_ij q_ij (y_i - y_j), q = 1/(1+d2/df)
// (tsne.cpp:1121-1137; exaggeration is applied later as a scalar).  It depends on Y and P only; it runs in line after the
// gather (beside other kernels it saturates every SM's load/store pipe and nothing is gained: see enqueue_iteration).
// Row offsets are local to this rank's edge slice: edges of row i are [row_P[i]-edge_base, row_P[i+1]-edge_base).
// Persistent form: a fixed grid (8 CTAs per SM, set by the host) strides over the row groups.
template <int D, int LPR>
__global__ void __launch_bounds__(256) k_attract(const uint32_t *__restrict__ row_P, const uint2 *__restrict__ edges, uint32_t edge_base,
                                                 const float *__restrict__ Y, int row_begin, int row_end, float inv_df,
                                                 float *__restrict__ attr) {
    pdl_prologue();
    constexpr int RPB = 256 / LPR;                       // rows per CTA per trip
    const int sub = threadIdx.x % LPR;
    const int nrows = row_end - row_begin;
    for (int base = blockIdx.x * RPB; base < nrows; base += gridDim.x * RPB) {
        const int row = row_begin + base + threadIdx.x / LPR;
        const bool active = row < row_end;
        float ax = 0.f, ay = 0.f;
        if (active) {
            float yix, yiy = 0.f;
            if (D == 2) { const float2 yi = reinterpret_cast<const float2 *>(Y)[row]; yix = yi.x; yiy = yi.y; }
            else yix = Y[row];
            const uint32_t e0 = row_P[row] - edge_base, e1 = row_P[row + 1] - edge_base;
#pragma unroll 4
            for (uint32_t e = e0 + sub; e < e1; e += LPR) {
                const uint2 ed = __ldg(edges + e);                  // one 8-byte load per edge: (column, weight)
                const uint32_t j = ed.x;
                const float pv = __uint_as_float(ed.y);
                if (D == 2) {
                    const float2 yj = __ldg(reinterpret_cast<const float2 *>(Y) + j);
                    const float dx = yix - yj.x, dy = yiy - yj.y;
                    const float q = __fdividef(pv, 1.f + (dx * dx + dy * dy) * inv_df);
                    ax += q * dx; ay += q * dy;
                } else {
                    const float dx = yix - __ldg(Y + j);
                    const float q = __fdividef(pv, 1.f + dx * dx * inv_df);
                    ax += q * dx;
                }
            }
        }
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) {
            ax += __shfl_xor_sync(0xffffffffu, ax, o);
            if (D == 2) ay += __shfl_xor_sync(0xffffffffu, ay, o);
        }
        if (active && sub == 0) {
            if (D == 2) reinterpret_cast<float2 *>(attr)[row] = make_float2(ax, ay);
            else attr[row] = ax;
        }
    }
}

// k_update: dC = alpha*attr - F_rep/Z (tsne.cpp:1153-1154), then either
//   UPDATE=false: write dC (parity entry point), or
//   UPDATE=true : gains / momentum / clipping / Y += uY (tsne.cpp:479-513) into Ynext (un-centred).
__device__ __forceinline__ float sgnf(float x) { return x == 0.f ? 0.f : (x < 0.f ? -1.f : 1.f); }

template <int D, bool UPDATE>
__device__ __forceinline__ void update_row(int row, const float *__restrict__ Y, const float *__restrict__ attr,
                                           const float *__restrict__ frep, const StepParams &sp, float *__restrict__ dC_out,
                                           float *__restrict__ uY, float *__restrict__ gains, float *__restrict__ Ynext,
                                           float &new0, float &new1) {
    float d0, d1 = 0.f, yix, yiy = 0.f;
    if (D == 2) {
        const float2 at = reinterpret_cast<const float2 *>(attr)[row], fr = reinterpret_cast<const float2 *>(frep)[row];
        const float2 yi = reinterpret_cast<const float2 *>(Y)[row];
        d0 = sp.alpha * at.x - fr.x; d1 = sp.alpha * at.y - fr.y;
        yix = yi.x; yiy = yi.y;
    } else {
        d0 = sp.alpha * attr[row] - frep[row];
        yix = Y[row];
    }
    if (!UPDATE) {
        if (D == 2) reinterpret_cast<float2 *>(dC_out)[row] = make_float2(d0, d1);
        else dC_out[row] = d0;
        return;
    }
    if (sp.mode == 2) {   // plain gradient descent, no learning rate (tsne.cpp:489)
        new0 = yix - d0; new1 = yiy - d1;
        if (D == 2) reinterpret_cast<float2 *>(Ynext)[row] = make_float2(new0, new1);
        else Ynext[row] = new0;
        return;
    }
    float u0, u1 = 0.f, g0, g1 = 1.f;
    if (D == 2) {
        const float2 u = reinterpret_cast<const float2 *>(uY)[row], g = reinterpret_cast<const float2 *>(gains)[row];
        u0 = u.x; u1 = u.y; g0 = g.x; g1 = g.y;
    } else { u0 = uY[row]; g0 = gains[row]; }
    g0 = (sgnf(d0) != sgnf(u0)) ? (g0 + .2f) : (g0 * .8f);
    if (g0 < .01f) g0 = .01f;
    u0 = sp.momentum * u0 - sp.lr * g0 * d0;
    if (D == 2) {
        g1 = (sgnf(d1) != sgnf(u1)) ? (g1 + .2f) : (g1 * .8f);
        if (g1 < .01f) g1 = .01f;
        u1 = sp.momentum * u1 - sp.lr * g1 * d1;
    }
    if (sp.mode == 0 && sp.max_step_norm > 0.f) {
        const float step = sqrtf(u0 * u0 + u1 * u1);
        if (step > sp.max_step_norm) { const float f = sp.max_step_norm / step; u0 *= f; u1 *= f; }
    }
    new0 = yix + u0; new1 = yiy + u1;
    if (D == 2) {
        reinterpret_cast<float2 *>(gains)[row] = make_float2(g0, g1);
        reinterpret_cast<float2 *>(uY)[row] = make_float2(u0, u1);
        reinterpret_cast<float2 *>(Ynext)[row] = make_float2(new0, new1);
    } else {
        gains[row] = g0; uY[row] = u0; Ynext[row] = new0;
    }
}

// Persistent form: gridDim.x CTAs, each walking a contiguous slice of the rows, so that the column sums of the new
// positions (for the zero-mean step, tsne.cpp:1851-1876) ride along in registers and cost ONE block reduction per CTA at
// the very end; the last CTA to finish adds the per-CTA partials in index order and publishes the means.  (A first
// attempt with one 256-row CTA per block reduction was 8 us slower than a separate column-sum pass; this one has the
// separate pass's summation order -- slice per CTA, thread-strided, fixed tree -- without its second read of Ynext.)
template <int D, bool UPDATE>
__global__ void __launch_bounds__(256) k_update(const float *__restrict__ Y, const float *__restrict__ attr,
                                                const float *__restrict__ frep, int row_begin, int row_end,
                                                const StepParams *__restrict__ spp, const GridParams *__restrict__ gpp,
                                                float *__restrict__ dC_out, float *__restrict__ uY, float *__restrict__ gains,
                                                float *__restrict__ Ynext, double *__restrict__ colsum_partial, int N_total,
                                                Scalars *__restrict__ sc, unsigned int *__restrict__ ticket) {
    pdl_prologue();
    if (!gpp->ok) return;
    const StepParams sp = *spp;
    const int per = (row_end - row_begin + gridDim.x - 1) / gridDim.x;
    const int b = row_begin + blockIdx.x * per, e = min(row_end, b + per);
    double s0 = 0, s1 = 0;
    for (int row = b + threadIdx.x; row < e; row += blockDim.x) {
        float new0 = 0.f, new1 = 0.f;           // this row's new (un-centred) position
        update_row<D, UPDATE>(row, Y, attr, frep, sp, dC_out, uY, gains, Ynext, new0, new1);
        s0 += new0; s1 += new1;
    }
    if (!UPDATE || colsum_partial == nullptr) return;
    __shared__ double smu[32];
    const double r0 = block_sum(s0, smu);
    if (threadIdx.x == 0) colsum_partial[blockIdx.x * D] = r0;
    if (D == 2) {
        const double r1 = block_sum(s1, smu);
        if (threadIdx.x == 0) colsum_partial[blockIdx.x * D + 1] = r1;
    }
    if (last_block_done(ticket)) {
        for (int d = 0; d < D; d++) {
            double s = 0;
            for (int i = threadIdx.x; i < (int) gridDim.x; i += blockDim.x) s += ld_partial(colsum_partial + i * D + d);
            const double r = block_sum(s, smu);
            if (threadIdx.x == 0) sc->mean[d] = r / (double) N_total;
        }
    }
}


// ------------------------------------------------------------------------- sharded zero-mean + bounds --
// Multi-GPU tail of an optimiser step.  Each rank has just written the new, un-centred positions of ITS points
// (Ynext[row_begin..row_end)).  Instead of all-gathering Y first and then reducing over all N points on every GPU, each
// rank reduces its own slice (inside k_update_shard), the per-rank records (128 bytes) are all-gathered, and every rank derives
// the same global column means and bounds from the same bytes (k_center_shard) while centring only its own slice.  The
// positions themselves travel at the start of the NEXT iteration (copy-engine pushes into the peers' Y on a side stream, or
// an NCCL all-gather without peer access), overlapped with that iteration's sort / spread, which only read the local slice.
//
// The 2-D scan quirk (tsne.cpp:1045-1048, see k_center_bounds) acts on the centred, interleaved sequence from flat
// index 0, i.e. on the head of rank 0's slice -- but the means are only known after the exchange.  Rank 0 therefore
// ships the raw new positions of its first SHARD_HEAD points and leaves them out of its slice minimum; every rank
// replays the scan on that head.  Exact unless the strictly ascending prefix is longer than 2*SHARD_HEAD values
// (probability 1/16! for continuous data); a longer prefix is cut there.
constexpr int SHARD_HEAD = 8;
struct ShardStats {              // one per rank, 128 bytes
    double sum[2];               // column sums of the rank's new (un-centred) positions
    float mn[2], mx[2];          // per-dimension min (head points excluded on rank 0) / max (all points)
    float head[2 * SHARD_HEAD];  // rank 0: first points, interleaved; unused elsewhere
    int nhead;                   // number of valid head VALUES (flat), 0 on ranks > 0 and in 1-D
    int pad_[7];
};
static_assert(sizeof(ShardStats) == 128, "ShardStats is exchanged as 128 raw bytes");

// Optimiser step of my slice + its statistics in ONE pass: the new positions are still in registers when their column sums
// and bounds are taken (persistent grid like k_update: contiguous slice per CTA, thread-strided, fixed trees).  The CTA
// that finishes last reduces the per-CTA partials, builds the 128-byte record and ships it -- every step of that tail
// spread over the block's threads: a single thread walking a few hundred partials, the head and 32 remote words one
// dependent L2 / NVLink latency at a time cost 40 us per iteration on 2 x B200.
template <int D>
__global__ void __launch_bounds__(256) k_update_shard(const float *__restrict__ Y, const float *__restrict__ attr,
                                                      const float *__restrict__ frep, int row_begin, int row_end, int rank,
                                                      const StepParams *__restrict__ spp, const GridParams *__restrict__ gpp,
                                                      float *__restrict__ dC_out, float *__restrict__ uY, float *__restrict__ gains,
                                                      float *__restrict__ Ynext, double *__restrict__ sum_partial,
                                                      float4 *__restrict__ mm_partial, ShardStats *__restrict__ out,
                                                      unsigned int *__restrict__ ticket, PeerComm pc, int p2p,
                                                      const uint32_t *__restrict__ orig_of, const uint32_t *__restrict__ pos_of) {
    pdl_prologue();
    if (!gpp->ok) return;
    __shared__ double smd[32];
    __shared__ float4 smm[8];
    __shared__ ShardStats rec;
    const StepParams sp = *spp;
    const int n = row_end - row_begin;
    // the head = the first SHARD_HEAD points in the caller's ORIGINAL order (rank 0 owns them; after a re-ordering of the
    // slice they sit at pos_of[0..])
    const int nhead_pts = (rank == 0 && D == 2) ? min(SHARD_HEAD, n) : 0;
    const int per = (n + gridDim.x - 1) / gridDim.x;
    const int b = blockIdx.x * per, e = min(n, b + per);
    double s0 = 0, s1 = 0;
    float4 mm = make_float4(INFINITY, INFINITY, -INFINITY, -INFINITY);      // (min0, min1, max0, max1)
    for (int i = b + threadIdx.x; i < e; i += blockDim.x) {
        float v0 = 0.f, v1 = 0.f;               // this row's new (un-centred) position
        update_row<D, true>(row_begin + i, Y, attr, frep, sp, dC_out, uY, gains, Ynext, v0, v1);
        s0 += v0;
        if (D == 2) {
            s1 += v1;
            mm.z = fmaxf(mm.z, v0); mm.w = fmaxf(mm.w, v1);
            const int o = orig_of ? (int) orig_of[row_begin + i] - row_begin : i;
            if (o >= nhead_pts) { mm.x = fminf(mm.x, v0); mm.y = fminf(mm.y, v1); }
        } else {
            mm.x = fminf(mm.x, v0); mm.z = fmaxf(mm.z, v0);
        }
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    auto reduce_mm = [&](float4 v) -> float4 {           // result in thread 0
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            v.x = fminf(v.x, __shfl_xor_sync(0xffffffffu, v.x, o)); v.y = fminf(v.y, __shfl_xor_sync(0xffffffffu, v.y, o));
            v.z = fmaxf(v.z, __shfl_xor_sync(0xffffffffu, v.z, o)); v.w = fmaxf(v.w, __shfl_xor_sync(0xffffffffu, v.w, o));
        }
        __syncthreads();
        if (lane == 0) smm[w] = v;
        __syncthreads();
        if (threadIdx.x == 0)
            for (int i = 1; i < (int) (blockDim.x >> 5); i++) {
                v.x = fminf(v.x, smm[i].x); v.y = fminf(v.y, smm[i].y); v.z = fmaxf(v.z, smm[i].z); v.w = fmaxf(v.w, smm[i].w);
            }
        return v;
    };
    const double r0 = block_sum(s0, smd);
    const double r1 = D == 2 ? block_sum(s1, smd) : 0.0;
    mm = reduce_mm(mm);
    if (threadIdx.x == 0) {
        sum_partial[blockIdx.x * 2] = r0; sum_partial[blockIdx.x * 2 + 1] = r1;
        mm_partial[blockIdx.x] = mm;
    }
    // (last_block_done's device-scope fence in thread 0 follows block-wide barriers: it also covers the Ynext rows the other
    // threads of this CTA stored, which the head below reads back)
    if (!last_block_done(ticket)) return;
    double t0 = 0, t1 = 0;
    float4 a = make_float4(INFINITY, INFINITY, -INFINITY, -INFINITY);
    for (int i = threadIdx.x; i < (int) gridDim.x; i += blockDim.x) {
        t0 += __ldcg(sum_partial + 2 * i); t1 += __ldcg(sum_partial + 2 * i + 1);
        const float4 v = __ldcg(mm_partial + i);
        a.x = fminf(a.x, v.x); a.y = fminf(a.y, v.y); a.z = fmaxf(a.z, v.z); a.w = fmaxf(a.w, v.w);
    }
    t0 = block_sum(t0, smd);
    t1 = block_sum(t1, smd);
    a = reduce_mm(a);
    if (threadIdx.x == 0) {
        rec.sum[0] = t0; rec.sum[1] = t1;
        rec.mn[0] = a.x; rec.mn[1] = a.y; rec.mx[0] = a.z; rec.mx[1] = a.w;
        rec.nhead = nhead_pts * 2;
        for (int i = 0; i < 7; i++) rec.pad_[i] = 0;
    }
    if ((int) threadIdx.x >= 32 && (int) threadIdx.x < 32 + 2 * SHARD_HEAD) {
        const int i = threadIdx.x - 32;
        float v = 0.f;
        if (i < nhead_pts * 2) {
            const size_t pt = pos_of ? (size_t) pos_of[row_begin + (i >> 1)] : (size_t) row_begin + (i >> 1);
            v = __ldcg(Ynext + pt * 2 + (i & 1));
        }
        rec.head[i] = v;
    }
    __syncthreads();
    // the record: slot [rank] of my own table and -- peer fabric -- of every peer's (one 4-byte remote store per lane and
    // peer), then the flag; k_center_shard on the other side spins on it
    if (threadIdx.x < sizeof(ShardStats) / 4) {
        const uint32_t wv = reinterpret_cast<const uint32_t *>(&rec)[threadIdx.x];
        reinterpret_cast<uint32_t *>(out)[threadIdx.x] = wv;
        if (p2p)
            for (int r = 0; r < pc.world; r++)
                if (r != pc.rank)
                    reinterpret_cast<volatile uint32_t *>(reinterpret_cast<ShardStats *>(pc.stats[r]) + pc.rank)[threadIdx.x] = wv;
    }
    if (p2p) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            const uint32_t seq = *reinterpret_cast<volatile unsigned int *>(pc.seq);
            for (int r = 0; r < pc.world; r++)
                if (r != pc.rank) *reinterpret_cast<volatile uint32_t *>(pc.flags[r] + FLAG_STATS * pc.world + pc.rank) = seq;
        }
    }
}

// Every rank: global means and bounds from the all-gathered records (identical bytes => identical results on all
// ranks), centre the local slice Y[row] = Ynext[row] - mean (tsne.cpp:1851-1876), publish the bounds.  Thread r of every
// CTA waits for rank r's record and fetches it (the waits and the loads of all ranks overlap), thread 0 adds in rank order.
template <int D>
__global__ void __launch_bounds__(256) k_center_shard(const float *__restrict__ Ynext, float *__restrict__ Y, int row_begin, int row_end,
                                                      int N, const ShardStats *all, int world,
                                                      const GridParams *__restrict__ gpp, Scalars *__restrict__ sc,
                                                      volatile float *host_bounds, PeerComm pc, int p2p) {
    pdl_prologue();
    if (!gpp->ok) return;
    __shared__ double mean_s[2];
    __shared__ double sum_s[2 * MAX_RANKS];
    __shared__ float mm_s[4 * MAX_RANKS];
    __shared__ float head_s[2 * SHARD_HEAD];
    const volatile ShardStats *va = all;           // (peers may have written these records: no cached / read-only loads)
    if ((int) threadIdx.x < world) {
        const int r = threadIdx.x;
        if (p2p && r != pc.rank) {
            const uint32_t seq = *reinterpret_cast<volatile unsigned int *>(pc.seq);
            const volatile uint32_t *f = pc.flags[pc.rank] + FLAG_STATS * pc.world + r;
            while ((int) (*f - seq) < 0) { }
            __threadfence_system();
        }
        sum_s[2 * r] = va[r].sum[0]; sum_s[2 * r + 1] = va[r].sum[1];
        if (blockIdx.x == 0) {
            mm_s[4 * r] = va[r].mn[0]; mm_s[4 * r + 1] = va[r].mn[1]; mm_s[4 * r + 2] = va[r].mx[0]; mm_s[4 * r + 3] = va[r].mx[1];
        }
    }
    __syncthreads();
    if (blockIdx.x == 0 && (int) threadIdx.x < 2 * SHARD_HEAD) head_s[threadIdx.x] = va[0].head[threadIdx.x];   // (record 0 is in: barrier above)
    if (threadIdx.x == 0) {
        double t0 = 0, t1 = 0;
        for (int r = 0; r < world; r++) { t0 += sum_s[2 * r]; t1 += sum_s[2 * r + 1]; }
        mean_s[0] = t0 / (double) N; mean_s[1] = t1 / (double) N;
    }
    __syncthreads();
    const double m0 = mean_s[0], m1 = mean_s[1];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        // centring is monotonic per dimension ((float)((double) y - mean)), so the bounds of the centred values are the
        // centred per-dimension bounds
        float bmn = INFINITY, bmx = -INFINITY;
        for (int r = 0; r < world; r++) {
            for (int d = 0; d < D; d++) {
                const double m = d ? m1 : m0;
                bmn = fminf(bmn, (float) ((double) mm_s[4 * r + d] - m));
                bmx = fmaxf(bmx, (float) ((double) mm_s[4 * r + 2 + d] - m));
            }
        }
        const int nh = va[0].nhead;
        float run = -INFINITY;
        bool ascending = true;
        for (int i = 0; i < nh; i++) {            // replay of the `if (>max) .. else if (<min)` scan on the head
            const float v = (float) ((double) head_s[i] - ((i & 1) ? m1 : m0));
            if (ascending && v > run) run = v;   // still in the strictly ascending prefix: max only
            else { ascending = false; bmn = fminf(bmn, v); }
        }
        sc->mean[0] = m0; sc->mean[1] = m1;
        sc->bmin = bmn; sc->bmax = bmx;
        sc->iter_done += 1;
        if (host_bounds) {
            host_bounds[0] = bmn; host_bounds[1] = bmx;
            *reinterpret_cast<volatile unsigned long long *>(host_bounds + 4) = sc->iter_done;
        }
    }
    const int i = row_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= row_end) return;
    if (D == 2) {
        float2 v = reinterpret_cast<const float2 *>(Ynext)[i];
        v.x = (float) ((double) v.x - m0);
        v.y = (float) ((double) v.y - m1);
        reinterpret_cast<float2 *>(Y)[i] = v;
    } else {
        Y[i] = (float) ((double) Ynext[i] - m0);
    }
}

// ------------------------------------------------------------- locality re-ordering + tiled attractive term --
// k_attract over a plain CSR is bound by L1 gather wavefronts (one 128-byte line per random neighbour), not by
// DRAM.  The tiled path removes that bound: every few hundred iterations the points are physically re-ordered
// along a Morton curve of the current embedding (neighbours in P are neighbours in Y by then), the edges are
// regrouped into tiles (row chunk of <= TILE_ROWS points) x (column block of TILE_COLS points), and
// k_attract_tiles keeps the chunk's own positions, its accumulators and one column block of Y in shared memory:
// neighbour gathers become shared-memory reads, the edge stream (8 B/edge) is the only HBM traffic.
// Row sums are accumulated as 64-bit fixed point (2^-60) with shared-memory integer atomics: integer addition is
// associative, so the result does not depend on the order edges arrive in -- bitwise repeatable.
constexpr int TILE_ROWS = 4096;    // max points per row chunk  (row-local index fits 16 bits)
constexpr int TILE_COLS = 14336;   // points per column block    (112 KB of float2)

struct TileGeom {
    int rows_per_chunk, nchunks, ncb;   // chunk c = rows [c*rows_per_chunk, ...), column block b = points [b*TILE_COLS, ...)
};

__device__ __forceinline__ uint32_t spread_bits11(uint32_t v) {   // 11 bits -> every other bit of 22
    v &= 0x7ffu;
    v = (v | (v << 8)) & 0x00ff00ffu;
    v = (v | (v << 4)) & 0x0f0f0f0fu;
    v = (v | (v << 2)) & 0x33333333u;
    v = (v | (v << 1)) & 0x55555555u;
    return v;
}

// 22-bit locality key of every point: Morton code of the position quantised to 2048 cells per axis (2-D) or the
// 22-bit quantised coordinate (1-D).
template <int D>
__global__ void __launch_bounds__(256) k_morton_keys(const float *__restrict__ Y, int n, const Scalars *__restrict__ sc,
                                                     uint32_t *__restrict__ keys) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float mn = sc->bmin, inv = 1.f / fmaxf(sc->bmax - sc->bmin, 1e-30f);
    if (D == 2) {
        const float2 y = reinterpret_cast<const float2 *>(Y)[i];
        const int qx = min(2047, max(0, (int) ((y.x - mn) * inv * 2048.f)));
        const int qy = min(2047, max(0, (int) ((y.y - mn) * inv * 2048.f)));
        keys[i] = spread_bits11((uint32_t) qx) | (spread_bits11((uint32_t) qy) << 1);
    } else {
        keys[i] = (uint32_t) min(4194303, max(0, (int) ((Y[i] - mn) * inv * 4194304.f)));
    }
}

// after sorting: perm[k] = previous position of the point that moves to position k.
// orig_of[k] = ORIGINAL index of the point now at k; pos_of[o] = current position of original point o;
// rank[prev] = new position (used to relabel the CSR columns).
__global__ void __launch_bounds__(256) k_reorder_maps(const uint32_t *__restrict__ perm, int n, uint32_t *__restrict__ rank,
                                                      const uint32_t *__restrict__ orig_old, uint32_t *__restrict__ orig_new,
                                                      uint32_t *__restrict__ pos_of) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t prev = perm[k];
    rank[prev] = (uint32_t) k;
    const uint32_t o = orig_old ? orig_old[prev] : prev;
    orig_new[k] = o;
    pos_of[o] = (uint32_t) k;
}

// Sharded contexts re-order WITHIN their slice (a block-diagonal permutation: ownership of the rows does not change, so no
// CSR row ever has to move between ranks).  perm[k] = previous LOCAL position of the point that moves to local position k;
// the maps are global (slice offset `base`), every rank fills its own slice and the slices are all-gathered afterwards.
__global__ void __launch_bounds__(256) k_reorder_maps_local(const uint32_t *__restrict__ perm, int nloc, uint32_t base,
                                                            uint32_t *__restrict__ rank, const uint32_t *__restrict__ orig_old,
                                                            uint32_t *__restrict__ orig_new, uint32_t *__restrict__ pos_of) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nloc) return;
    const uint32_t prev = base + perm[k];
    rank[prev] = base + (uint32_t) k;
    const uint32_t o = orig_old ? orig_old[prev] : prev;     // stays inside [base, base + nloc)
    orig_new[base + k] = o;
    pos_of[o] = base + (uint32_t) k;
}
// Relabel this rank's CSR rows: 8 lanes per OLD local row; new local row = rank[row] - row_begin, columns through the
// (all-gathered) global rank map.  PASS 0: row lengths; PASS 1: copy the edges to their new offsets.
__global__ void __launch_bounds__(256) k_relabel_csr_local(int pass, const uint32_t *__restrict__ row_old, uint32_t edge_base,
                                                           const uint2 *__restrict__ edges_old, const uint32_t *__restrict__ rank,
                                                           int row_begin, int nloc, uint32_t *__restrict__ new_len,
                                                           const uint32_t *__restrict__ row_new, uint2 *__restrict__ edges_new) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int sub = gid & 7, il = gid >> 3;
    if (il >= nloc) return;
    const uint32_t r = rank[row_begin + il] - (uint32_t) row_begin;
    const uint32_t e0 = row_old[row_begin + il] - edge_base, e1 = row_old[row_begin + il + 1] - edge_base;
    if (pass == 0) {
        if (sub == 0) new_len[r] = e1 - e0;
    } else {
        const uint32_t nb = row_new[r];
        for (uint32_t e = e0 + sub; e < e1; e += 8) {
            const uint2 ed = edges_old[e];
            edges_new[nb + (e - e0)] = make_uint2(rank[ed.x], ed.y);
        }
    }
}
// row_P entries of the local rows from the scanned local lengths
__global__ void __launch_bounds__(256) k_local_row_offsets(const uint32_t *__restrict__ row_new, int nloc, int row_begin, uint32_t edge_base,
                                                           uint32_t *__restrict__ row_P) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r <= nloc) row_P[row_begin + r] = edge_base + row_new[r];
}

template <int D>
__global__ void __launch_bounds__(256) k_permute_rows(const float *__restrict__ in, float *__restrict__ out,
                                                      const uint32_t *__restrict__ perm, int n) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    if (D == 2) reinterpret_cast<float2 *>(out)[k] = reinterpret_cast<const float2 *>(in)[perm[k]];
    else out[k] = in[perm[k]];
}

// host <-> device transfers in ORIGINAL point order: dev[k] corresponds to host[orig_of[k]]
template <int D>
__global__ void __launch_bounds__(256) k_d2f_ordered(const double *__restrict__ in, float *__restrict__ out,
                                                     const uint32_t *__restrict__ orig_of, int n) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const size_t o = orig_of[k];
    for (int d = 0; d < D; d++) out[(size_t) k * D + d] = (float) in[o * D + d];
}
template <int D>
__global__ void __launch_bounds__(256) k_f2d_ordered(const float *__restrict__ in, double *__restrict__ out,
                                                     const uint32_t *__restrict__ orig_of, int n) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const size_t o = orig_of[k];
    for (int d = 0; d < D; d++) out[o * D + d] = (double) in[(size_t) k * D + d];
}
template <int D>
__global__ void __launch_bounds__(256) k_f2f_ordered(const float *__restrict__ in, float *__restrict__ out,
                                                     const uint32_t *__restrict__ orig_of, int n) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const size_t o = orig_of[k];
    for (int d = 0; d < D; d++) out[o * D + d] = in[(size_t) k * D + d];
}

// single-CTA exclusive scan (re-ordering time only); out[n] = total
__global__ void __launch_bounds__(1024) k_scan_excl(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, int n) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry_s, total_s;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 4096) {
        uint32_t v[4], tsum = 0;
        const int i0 = base + threadIdx.x * 4;
#pragma unroll
        for (int j = 0; j < 4; j++) { const int i = i0 + j; const uint32_t x = i < n ? in[i] : 0u; v[j] = tsum; tsum += x; }
        uint32_t inc = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) wsum[w] = inc;
        __syncthreads();
        if (w == 0) {
            const uint32_t ws = wsum[lane];
            uint32_t winc = ws;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += t; }
            wsum[lane] = winc - ws;
            if (lane == 31) total_s = winc;
        }
        __syncthreads();
        const uint32_t excl = carry_s + wsum[w] + (inc - tsum);
#pragma unroll
        for (int j = 0; j < 4; j++) { const int i = i0 + j; if (i < n) out[i] = excl + v[j]; }
        __syncthreads();
        if (threadIdx.x == 0) carry_s += total_s;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry_s;
}

// Relabel the CSR into the new point order.  One 8-lane group per OLD row i (new row r = rank[i]).
//   PASS 0: new_len[r] = len(i); tile_cnt[tile(r, rank[col])]++
//   PASS 1: copy the row's edges to new_row_P[r] (relabelled columns) and -- tile_pack != nullptr -- scatter them into their tiles
__global__ void __launch_bounds__(256) k_relabel_csr(int pass, const uint32_t *__restrict__ row_old, const uint2 *__restrict__ edges_old,
                                                     const uint32_t *__restrict__ rank, int n,
                                                     TileGeom tg, uint32_t *__restrict__ new_len, uint32_t *__restrict__ tile_cnt,
                                                     const uint32_t *__restrict__ row_new, uint2 *__restrict__ edges_new,
                                                     const uint32_t *__restrict__ tile_start,
                                                     uint32_t *__restrict__ tile_cur, uint32_t *__restrict__ tile_pack,
                                                     float *__restrict__ tile_val) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int sub = gid & 7, i = gid >> 3;
    if (i >= n) return;
    const uint32_t r = rank[i];
    const uint32_t e0 = row_old[i], e1 = row_old[i + 1];
    const uint32_t rc = r / (uint32_t) tg.rows_per_chunk, rl = r - rc * (uint32_t) tg.rows_per_chunk;
    if (pass == 0) {
        if (sub == 0) new_len[r] = e1 - e0;
        for (uint32_t e = e0 + sub; e < e1; e += 8) {
            const uint32_t c = rank[edges_old[e].x];
            atomicAdd(&tile_cnt[(size_t) rc * tg.ncb + c / TILE_COLS], 1u);
        }
    } else {
        const uint32_t nb = row_new[r];
        for (uint32_t e = e0 + sub; e < e1; e += 8) {
            const uint2 ed = edges_old[e];
            const uint32_t c = rank[ed.x];
            const float v = __uint_as_float(ed.y);
            edges_new[nb + (e - e0)] = make_uint2(c, ed.y);
            if (tile_pack == nullptr) continue;              // the CSR path was chosen: nobody reads the tiles
            const uint32_t cb = c / TILE_COLS;
            const size_t t = (size_t) rc * tg.ncb + cb;
            const uint32_t pos = tile_start[t] + atomicAdd(&tile_cur[t], 1u);
            tile_pack[pos] = (rl << 16) | (c - cb * TILE_COLS);
            tile_val[pos] = v;
        }
    }
}

__global__ void __launch_bounds__(256) k_count_nonempty(const uint32_t *__restrict__ tile_cnt, size_t ntiles, uint32_t *__restrict__ out) {
    const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    const bool ne = t < ntiles && tile_cnt[t] != 0;
    const uint32_t m = __ballot_sync(0xffffffffu, ne);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(out, (uint32_t) __popc(m));
}

// One CTA per row chunk.  Shared memory: own positions [R] | fixed-point accumulators [R][D] | column block [TILE_COLS].
// Accumulation: 32-bit fixed point scaled by fix32 = 2^30 / (max row sum of P * max(1, sqrt(df))) (|q dx| / p <= sqrt(df) / 2).
template <int D>
__global__ void __launch_bounds__(1024, 1) k_attract_tiles(const float *__restrict__ Y, int n, TileGeom tg,
                                                           const uint32_t *__restrict__ tile_start,
                                                           const uint32_t *__restrict__ tile_pack, const float *__restrict__ tile_val,
                                                           float inv_df, float fix32, float *__restrict__ attr) {
    pdl_prologue();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int R = tg.rows_per_chunk;
    using YT = typename std::conditional<D == 2, float2, float>::type;
    YT *yrow = reinterpret_cast<YT *>(smem_raw);
    long long *acc = reinterpret_cast<long long *>(smem_raw + (size_t) TILE_ROWS * sizeof(YT));
    YT *ycol = reinterpret_cast<YT *>(smem_raw + (size_t) TILE_ROWS * sizeof(YT) + (size_t) TILE_ROWS * D * sizeof(long long));
    const int rc = blockIdx.x;
    const int r0 = rc * R, rn = min(R, n - r0);
    const YT *Yt = reinterpret_cast<const YT *>(Y);
    for (int i = threadIdx.x; i < rn; i += blockDim.x) yrow[i] = Yt[r0 + i];
    for (int i = threadIdx.x; i < rn * D; i += blockDim.x) acc[i] = 0;
    const uint32_t *ts = tile_start + (size_t) rc * tg.ncb;
    for (int cb = 0; cb < tg.ncb; cb++) {
        const uint32_t e0 = ts[cb], e1 = ts[cb + 1];
        if (e0 == e1) continue;                       // uniform across the CTA
        __syncthreads();                              // previous block's readers are done (also covers the init above)
        const int c0 = cb * TILE_COLS, cn = min(TILE_COLS, n - c0);
        for (int i = threadIdx.x; i < cn; i += blockDim.x) ycol[i] = Yt[c0 + i];
        __syncthreads();
        // 4 independent edges per thread per trip: keeps enough 8-byte loads in flight to stream at HBM rate
        for (uint32_t eb = e0 + threadIdx.x; eb < e1; eb += 4 * blockDim.x) {
            uint32_t pk[4];
            float pv[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t e = eb + u * blockDim.x;
                pk[u] = e < e1 ? tile_pack[e] : 0u;
                pv[u] = e < e1 ? tile_val[e] : 0.f;      // weight 0 contributes exactly 0
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (eb + u * blockDim.x >= e1) break;
                const uint32_t rl = pk[u] >> 16, cl = pk[u] & 0xffffu;
                if (D == 2) {
                    const float2 yi = reinterpret_cast<const float2 *>(yrow)[rl], yj = reinterpret_cast<const float2 *>(ycol)[cl];
                    const float dx = yi.x - yj.x, dy = yi.y - yj.y;
                    const float q = pv[u] / (1.f + (dx * dx + dy * dy) * inv_df);
                    atomicAdd(reinterpret_cast<int *>(&acc[2 * rl]), __float2int_rn(q * dx * fix32));
                    atomicAdd(reinterpret_cast<int *>(&acc[2 * rl + 1]), __float2int_rn(q * dy * fix32));
                } else {
                    const float dx = reinterpret_cast<const float *>(yrow)[rl] - reinterpret_cast<const float *>(ycol)[cl];
                    const float q = pv[u] / (1.f + dx * dx * inv_df);
                    atomicAdd(reinterpret_cast<int *>(&acc[rl]), __float2int_rn(q * dx * fix32));
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < rn; i += blockDim.x) {
        float a0, a1 = 0.f;
        a0 = (float) ((double) *reinterpret_cast<int *>(&acc[(D == 2 ? 2 : 1) * i]) / (double) fix32);
        if (D == 2) a1 = (float) ((double) *reinterpret_cast<int *>(&acc[2 * i + 1]) / (double) fix32);
        if (D == 2) reinterpret_cast<float2 *>(attr)[r0 + i] = make_float2(a0, a1);
        else attr[r0 + i] = a0;
    }
}

// ------------------------------------------------------------------------------------------------ KL --
// sum_edges (alpha p) log((alpha p + FLT_MIN) / (q + FLT_MIN)), q = (1+d2/df)^-df / sum_Q  (tsne.cpp:1340-1348);
// per-row sums in fp64, per-block partials reduced in fixed order by k_finalize_kl.
template <int D>
__global__ void __launch_bounds__(256) k_kl(const uint32_t *__restrict__ row_P, const uint2 *__restrict__ edges,
                                            uint32_t edge_base, const float *__restrict__ Y,
                                            int row_begin, int row_end, double alpha, double df,
                                            const Scalars *__restrict__ sc, double *__restrict__ partial) {
    __shared__ double sm[32];
    const double Z = sc->Z;
    const int lane = threadIdx.x & 31;
    const int warps_total = (gridDim.x * blockDim.x) >> 5;
    double acc = 0;
    for (int row = row_begin + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5); row < row_end; row += warps_total) {
        float yix, yiy = 0.f;
        if (D == 2) { const float2 yi = reinterpret_cast<const float2 *>(Y)[row]; yix = yi.x; yiy = yi.y; }
        else yix = Y[row];
        const uint32_t e0 = row_P[row] - edge_base, e1 = row_P[row + 1] - edge_base;
        for (uint32_t e = e0 + lane; e < e1; e += 32) {
            const uint2 ed = __ldg(edges + e);
            const uint32_t j = ed.x;
            const double pv = alpha * (double) __uint_as_float(ed.y);
            double d2;
            if (D == 2) {
                const float2 yj = reinterpret_cast<const float2 *>(Y)[j];
                const double dx = (double) yix - (double) yj.x, dy = (double) yiy - (double) yj.y;
                d2 = dx * dx + dy * dy;
            } else {
                const double dx = (double) yix - (double) Y[j];
                d2 = dx * dx;
            }
            double q = 1.0 / (1.0 + d2 / df);
            if (df != 1.0) q = pow(q, df);
            q /= Z;
            acc += pv * log((pv + (double) FLT_MIN) / (q + (double) FLT_MIN));
        }
    }
    const double r = block_sum(acc, sm);
    if (threadIdx.x == 0) partial[blockIdx.x] = r;
}

__global__ void __launch_bounds__(256) k_finalize_kl(const double *__restrict__ partial, int nparts, Scalars *__restrict__ sc) {
    __shared__ double sm[32];
    double s = 0;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) s += partial[i];
    const double r = block_sum(s, sm);
    if (threadIdx.x == 0) sc->kl = r;
}

// max row sum of P for the automatic exaggeration coefficient (tsne.cpp:392-399)
__global__ void __launch_bounds__(256) k_row_sum_max(const uint32_t *__restrict__ row_P, const uint2 *__restrict__ edges,
                                                     uint32_t edge_base, int row_begin, int row_end,
                                                     double *__restrict__ partial) {
    __shared__ double sm[32];
    double mx = 0;
    for (int row = row_begin + blockIdx.x * blockDim.x + threadIdx.x; row < row_end; row += gridDim.x * blockDim.x) {
        double s = 0;
        for (uint32_t e = row_P[row] - edge_base; e < row_P[row + 1] - edge_base; e++) s += (double) __uint_as_float(edges[e].y);
        mx = fmax(mx, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < (int) (blockDim.x >> 5); i++) mx = fmax(mx, sm[i]);
        partial[blockIdx.x] = mx;
    }
}

// device CSR edge word: (column u32, weight f32 bits) -- one 8-byte load per edge in the SpMV / KL kernels
__global__ void k_pack_edges(const uint32_t *__restrict__ col, const double *__restrict__ val, uint2 *__restrict__ edges, size_t n) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) edges[i] = make_uint2(col[i], __float_as_uint((float) val[i]));
}

// fp64 host data <-> fp32 device data
__global__ void k_d2f(const double *__restrict__ in, float *__restrict__ out, size_t n) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float) in[i];
}
__global__ void k_f2d(const float *__restrict__ in, double *__restrict__ out, size_t n) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (double) in[i];
}
__global__ void k_fill(float *__restrict__ out, float v, size_t n) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = v;
}

}  // namespace fk
