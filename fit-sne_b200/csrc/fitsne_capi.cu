// fitsne_capi.cu -- context, per-iteration launch sequence, CUDA-graph cache and the C ABI
// (include/fitsne_b200.h) of libfitsne_b200.so.  Kernels live in fitsne_kernels.cuh, fitsne_conv.cuh and fitsne_fft.cuh
// (no library call is left on the path: the FFTs are our own); NCCL (loaded with dlopen, only for sharded runs) carries
// the grid all-reduce and the Y all-gather.  There is no CPU fallback anywhere in this file.
#include "../../include/fitsne_b200.h"
#include "fitsne_kernels.cuh"
#include "fitsne_fft.cuh"
#include "fitsne_conv.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <vector>

using namespace fk;

// ------------------------------------------------------------------------------------------ NCCL (dlopen) --
namespace {
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string &err) {
        if (handle) return true;
        const char *env = getenv("FITSNE_NCCL_LIB");
        const char *names[] = {env, "libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            if (!n) continue;
            handle = dlopen(n, RTLD_NOW | RTLD_LOCAL);
            if (handle) break;
        }
        if (!handle) { err = std::string("cannot dlopen libnccl: ") + dlerror(); return false; }
#define LOADSYM(field, name) \
        *(void **) (&field) = dlsym(handle, name); \
        if (!field) { err = std::string("libnccl lacks ") + name; return false; }
        LOADSYM(GetUniqueId, "ncclGetUniqueId");
        LOADSYM(CommInitRank, "ncclCommInitRank");
        LOADSYM(CommDestroy, "ncclCommDestroy");
        LOADSYM(AllReduce, "ncclAllReduce");
        LOADSYM(AllGather, "ncclAllGather");
        LOADSYM(GetErrorString, "ncclGetErrorString");
#undef LOADSYM
        return true;
    }
};
NcclApi g_nccl;
std::string g_create_error;

struct Plans {          // per FFT length: radix plans + twiddle table (no library plans, nothing to JIT)
    FftPlan plan{};      // Stockham (natural order): row passes, 1-D lines
    ColPlan cplan{};     // in place (digit-reversed spectra): column passes of the 2-D convolution
    float2 *W = nullptr;
    size_t smem_row2 = 0, smem_row1 = 0, smem_col = 0, smem_line = 0;
    CUtensorMap tmS;     // 2-D: TMA view of S as [rows = M/2][(M/2+1) * 8 floats], box = 128 rows x 8 floats
    const void *tm_base = nullptr;   // the S allocation the map was encoded for
};

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

// The launch sequence depends on the FFT length M only (n_boxes and all grid geometry are read from the
// device-resident GridParams), so one captured graph serves every iteration whose 2G maps to the same M.
struct GraphKey {
    int M, kind;   // kind 0: gradient only, 1: full step
    bool operator<(const GraphKey &o) const {
        if (M != o.M) return M < o.M;
        return kind < o.kind;
    }
};
}  // namespace

struct fitsne_ctx {
    fitsne_config cfg{};
    int N = 0, D = 0;
    int rank = 0, world = 1, per = 0, row_begin = 0, row_end = 0, nloc = 0;
    int device = 0;
    bool df_is_one = true;
    int n_fwd = 0, n_kern = 0, n_inv = 0;
    cudaStream_t stream = nullptr;    // everything runs here
    ncclComm_t comm = nullptr;
    // peer-memory fabric (sharded contexts; see fitsne_kernels.cuh): the iteration's exchanges run over mapped peer memory
    bool p2p = false;
    bool dist_conv = false;           // 2-D: convolution distributed over the ranks (else replicated, with the grid sum fused into its loads)
    PeerComm pc{};
    uint32_t *peer_flags = nullptr;
    double *peer_zs = nullptr;        // per-rank sum_Q partials (peers write their slot)
    unsigned int *comm_seq = nullptr;
    float2 *grid1d = nullptr;         // 1-D: this rank's partial charge lines (peers read them; the sum goes to planes)
    std::vector<void *> ipc_opened;
    cudaStream_t stream_c = nullptr;  // copy-engine pushes of the Y slice to the peers (DMA only)
    cudaEvent_t ev_cfork = nullptr, ev_cjoin = nullptr;
    bool use_pdl = true, in_chain = false, prev_is_kernel = false;     // programmatic dependent launch bookkeeping (launch_k)
    int chunk = CHUNK;                // sorted points per spread thread (FITSNE_CHUNK overrides: tests/tools/chunk_sweep.py)
    cudaStream_t stream_k = nullptr;  // sharded runs: the kernel spectra, beside the sort and the spread
    cudaEvent_t ev_kfork = nullptr, ev_kjoin = nullptr;
    // sharded runs: per-rank reduction records (all-gathered, 128 B each) and whether c->Y currently holds every rank's slice
    ShardStats *shard_stats = nullptr;
    double *shard_sum_partial = nullptr;
    float4 *shard_mm_partial = nullptr;
    bool y_whole = true;

    // state (fp32)
    float *Y = nullptr, *Yb = nullptr, *uY = nullptr, *gains = nullptr, *frep = nullptr, *dC = nullptr, *attr = nullptr;
    size_t y_elems = 0;   // allocated elements per Y buffer (padded to per*world*D)
    // CSR
    uint32_t *row_P = nullptr;
    uint2 *edges = nullptr;           // (column, fp32 weight bits) per edge of this rank's rows
    uint32_t edge_base = 0;
    size_t E = 0;
    int lpr = 32;
    // locality re-ordering + tiled attractive term (single-GPU contexts)
    uint32_t *orig_of = nullptr, *orig_tmp = nullptr, *pos_of = nullptr, *rank_map = nullptr;   // u32[N]
    uint32_t *row_P2 = nullptr;       // second CSR buffer (re-labelled copy is built here, then swapped)
    uint2 *edges2 = nullptr;
    uint32_t *tile_cnt = nullptr, *tile_start = nullptr, *tile_cur = nullptr, *tile_pack = nullptr, *nonempty = nullptr;
    float *tile_val = nullptr;
    GridParams *gp_reorder = nullptr;
    TileGeom tg{};
    size_t ntiles = 0;
    bool reordered = false, use_tiles = false;
    uint64_t steps_total = 0;         // optimiser steps since creation (fitsne_reset_stats does not touch it: re-order scheduling)
    uint64_t last_reorder_iter = 0, reorder_interval = 50, reorders = 0;
    uint32_t nonempty_tiles = 0;
    float tile_fix32 = 1.0f;
    uint64_t kernel_launches_reorder = 0;
    // sort / bins
    uint32_t *keys[2] = {nullptr, nullptr}, *perm[2] = {nullptr, nullptr};
    float *sorted_u = nullptr;
    uint2 *box_range = nullptr;       // [first, end) sorted positions of every non-empty box (by-product of the spread walk)
    uint32_t *hist = nullptr, *sort_totals = nullptr, *sort_bases = nullptr, *sweep_state = nullptr;
    uint32_t *work = nullptr;         // spread work list: [0] = count, [1..] = boxes that span several chunks
    size_t box_cap = 0, hist_cap = 0;
    float4 *slots = nullptr;          // spread partials of boxes that cross a CTA boundary: [CTA][2][nodes]
    float4 *gpart = nullptr;          // run-time-nterms fallback only: per-chunk partials [chunk][2][nodes]
    // grids.  2-D: chg (spread result, float4 per node of the G x G grid), S (x-spectra per row, in place the convolved
    // half-spectra), KR / KS (kernel spectra after the row / column pass), pot (v1, Bx, By per node); 1-D: four packed lines
    float4 *chg = nullptr, *pot = nullptr, *KR = nullptr, *KS = nullptr;
    float2 *S = nullptr, *planes = nullptr;
    int grid_cap_M = 0;                              // FFT length the grid buffers are sized for
    // small stuff
    double *colsum_partial = nullptr, *zpartial = nullptr, *kl_partial = nullptr;
    float2 *bounds_partial = nullptr;
    GridParams *gp = nullptr;
    StepParams *sp = nullptr;
    Scalars *sc = nullptr;
    int *mismatch = nullptr;
    unsigned int *tickets = nullptr;   // last-block-done counters: [0] conv columns, [1] centre/bounds, [2] update, [3] shard stats, [4..6] sort, [7] combine, [8] rows fwd, [9] rows inv
    float *host_bounds = nullptr, *host_bounds_dev = nullptr;   // mapped pinned
    int *host_B = nullptr, *host_B_dev = nullptr;               // pinned word + its device-memory copy: the host's n_boxes for this iteration (0 = the device decides)
    Scalars *host_sc = nullptr;                                 // pinned staging for scalar read-back
    StepParams sp_host{};
    bool sp_valid = false;
    double *staging = nullptr;   // device fp64 staging for Y / dC transfers
    size_t staging_elems = 0;

    std::map<int, Plans> plans;
    std::set<int> nccl_warm;          // sharded: FFT lengths whose iteration has run eagerly once (before its graph is captured)
    struct GraphEntry { cudaGraphExec_t exec; uint64_t launches; };
    std::map<GraphKey, GraphEntry> graphs;
    int cur_B = -1, cur_M = -1;
    bool bounds_valid = false, have_grad = false;
    double last_run_ms = 0;
    fitsne_stats stats{};
    cudaEvent_t ev[FITSNE_PHASE_COUNT + 1] = {};
    bool timing_this_iter = false;
    // FITSNE_KTIMES=1 + FITSNE_FLAG_TIMERS: one CUDA event after every kernel of the iteration -> warm per-kernel times
    // (ncu's per-launch times are cold-cache); read back as text with fitsne_debug_copy(ctx, "ktimes", ...)
    bool ktimes_on = false;
    std::vector<cudaEvent_t> kt_ev;
    std::vector<const char *> kt_name;
    size_t kt_n = 0;
    std::map<std::string, std::pair<double, uint64_t>> kt_acc;
    std::string kt_text;
    FILE *src_col = nullptr, *src_val = nullptr;     // fitsne_create_from_files: the edges are streamed from P_col.dat / P_val.dat
    std::string err;
};

// ------------------------------------------------------------------------------------------------ errors --
static int fail(fitsne_ctx *c, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return fail(c, e_ == cudaErrorMemoryAllocation ? FITSNE_ENOMEM : FITSNE_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)
#define CKNCCL(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) \
    return fail(c, FITSNE_ENCCL, "%s failed: %s (%s:%d)", #call, g_nccl.GetErrorString(r_), __FILE__, __LINE__); } while (0)
#define CKRC(call) do { int rc_ = (call); if (rc_ != 0) return rc_; } while (0)
#define LAUNCH_CHECK() CK(cudaGetLastError())

static const bool g_trace = getenv("FITSNE_TRACE") != nullptr;
static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
#define TRACE(...) do { if (g_trace) { fprintf(stderr, "[fitsne %.3f] ", now_ms()); fprintf(stderr, __VA_ARGS__); fputc('\n', stderr); } } while (0)

static inline int cdiv(long long a, long long b) { return (int) ((a + b - 1) / b); }

static inline int max_fft_len(int D) { return D == 2 ? 4096 : 8192; }   // 1-D: two buffers + twiddles must fit in shared memory

// (Tried, to shorten fitsne_run_host's create (19 ms) and destroy (16 ms) at N = 1M, both measured on B200 with
// tests/tools/e2e_trace.py: (1) the stream-ordered pool, cudaMallocAsync / cudaFreeAsync with the release threshold raised --
// the first call of a process took 790 ms instead of 146 ms (the pool's first growth), later calls were no faster;
// (2) carving all ~60 buffers out of two or three slabs -- 115-137 ms per call against 118-156 ms, inside the box's noise: the
// cudaMalloc / cudaFree round trips are not where the time goes.  Neither is kept.)
template <typename T>
static int dev_alloc(fitsne_ctx *c, T **p, size_t count) {
    if (*p) { cudaFree(*p); *p = nullptr; }
    CK(cudaMalloc((void **) p, count * sizeof(T) + 256));
    return 0;
}

// --------------------------------------------------------------------------------------- capacity / plans --
// Largest grid a graph for FFT length M may be asked to handle: G <= M/2, hence B <= M/(2p).
static inline size_t max_boxes_for(const fitsne_ctx *c, int M) {
    const size_t Bmax = (size_t) std::max(1, M / (2 * c->cfg.nterms));
    return c->D == 2 ? Bmax * Bmax : Bmax;
}
static void drop_graphs(fitsne_ctx *c) {
    for (auto &g : c->graphs) cudaGraphExecDestroy(g.second.exec);
    c->graphs.clear();
}

static int ensure_grid_capacity(fitsne_ctx *c, int M) {
    const int D = c->D;
    const size_t nb = max_boxes_for(c, M);
    bool moved = false;
    if (nb + 2 > c->box_cap) {
        const size_t cap = nb + nb / 2 + 1024;
        CKRC(dev_alloc(c, &c->box_range, cap));
        c->box_cap = cap;
        moved = true;
    }
    if (M > c->grid_cap_M) {
        const size_t Mc = (size_t) M + M / 4;           // head room: the grid grows through a ladder of lengths
        const size_t Gc = Mc / 2 + 1, H = Mc / 2 + 2;
        if (D == 2) {
            CKRC(dev_alloc(c, &c->chg, Gc * Gc)); CKRC(dev_alloc(c, &c->pot, Gc * Gc));
            CKRC(dev_alloc(c, &c->S, Gc * H * COL_SLOTS));
            CKRC(dev_alloc(c, &c->KR, Gc * H)); CKRC(dev_alloc(c, &c->KS, H * Mc));
            // sharded runs all-reduce (M/2)^2 nodes of chg whatever G is: the tail must hold finite numbers
            CK(cudaMemsetAsync(c->chg, 0, Gc * Gc * sizeof(float4), c->stream));
        } else {
            CKRC(dev_alloc(c, &c->planes, Mc * 4));
        }
        c->grid_cap_M = (int) Mc;
        moved = true;
    }
    if (moved) drop_graphs(c);   // captured graphs point at the old buffers
    return 0;
}

static int get_plans(fitsne_ctx *c, int M, Plans **out) {
    auto it = c->plans.find(M);
    if (it == c->plans.end()) {
        Plans pl;
        // column plan: radices up to 16 where that saves passes over the tile (1152: 16 8 9 instead of 8 8 2 3 3 -- convolution
        // 0.109 -> 0.093 ms, 2 508 -> 2 597 it/s on B200); at equal stage counts the narrow radices win (320: 8 8 5 beats
        // 16 4 5 by 3 us).  FITSNE_COL_WIDE=0 / 1 forces one family.
        static const int col_wide_env = getenv("FITSNE_COL_WIDE") ? atoi(getenv("FITSNE_COL_WIDE")) : -1;
        ColPlan narrow, wide;
        const bool okn = col_make_plan(M, &narrow, false), okw = col_make_plan(M, &wide, true);
        const bool use_wide = okw && (col_wide_env >= 0 ? col_wide_env != 0 : (!okn || wide.nstages < narrow.nstages));
        pl.cplan = use_wide ? wide : narrow;
        if (!fft_make_plan(M, &pl.plan) || !(okn || okw))
            return fail(c, FITSNE_EINVAL, "FFT length %d is not of the form 2^a 3^b 5^c", M);
        CK(cudaMalloc((void **) &pl.W, (size_t) M * sizeof(float2)));
        k_fft_twiddles<<<cdiv(M, 256), 256, 0, c->stream>>>(pl.W, M);
        LAUNCH_CHECK();
        CK(cudaStreamSynchronize(c->stream));
        pl.smem_row2 = (size_t) 4 * fft_buf_len(M, 2) * sizeof(float2);          // two sequences, ping-pong
        pl.smem_row1 = (size_t) 2 * fft_buf_len(M, 1) * sizeof(float2);
        pl.smem_line = ((size_t) 2 * fft_buf_len(M, 1) + M) * sizeof(float2);    // 1-D: + twiddles
        pl.smem_col = (size_t) M * COL_SLOTS * sizeof(float2);
        if ((c->D == 2 ? std::max(pl.smem_row2, pl.smem_col) : pl.smem_line) > (size_t) 220 * 1024)
            return fail(c, FITSNE_EINVAL, "FFT length %d does not fit in shared memory", M);
        it = c->plans.emplace(M, pl).first;
    }
    Plans &pl = it->second;
    if (c->D == 2 && pl.tm_base != (const void *) c->S) {
        // S viewed as a 2-D fp32 tensor: inner = (M/2+1) frequencies x 4 slots x (re, im), outer = M/2 rows (>= G);
        // one box = one kx (8 floats = 32 bytes) x 128 rows.  Out-of-range rows read as zero / are not written.
        EncodeTiledFn enc = get_encode_tiled();
        if (!enc) return fail(c, FITSNE_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
        const cuuint64_t H = (cuuint64_t) M / 2 + 1;
        const cuuint64_t dims[2] = {H * 2 * COL_SLOTS, (cuuint64_t) M / 2};
        const cuuint64_t strides[1] = {H * COL_SLOTS * sizeof(float2)};
        const cuuint32_t box[2] = {2 * COL_SLOTS, COL_BOX_ROWS};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = enc(&pl.tmS, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *) c->S, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(c, FITSNE_ECUDA, "cuTensorMapEncodeTiled failed (%d) for M=%d", (int) r, M);
        pl.tm_base = c->S;
    }
    *out = &pl;
    return 0;
}

// ------------------------------------------------------------------------------------- launch sequences --
static inline void phase_mark(fitsne_ctx *c, int phase) {
    if (c->timing_this_iter) cudaEventRecord(c->ev[phase], c->stream);
}

static inline void kt(fitsne_ctx *c, const char *name) {
    if (!c->ktimes_on || !c->timing_this_iter) return;
    if (c->kt_n == c->kt_ev.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        c->kt_ev.push_back(e); c->kt_name.push_back(name);
    }
    c->kt_name[c->kt_n] = name;
    cudaEventRecord(c->kt_ev[c->kt_n], c->stream);
    c->kt_n++;
}

// Every kernel of the iteration chain goes through launch_k: a plain launch, or -- when the operation before it on the same
// stream was a kernel of the chain too (FITSNE_PDL=0 turns it off) -- one with programmatic stream serialisation (captured as a
// programmatic graph edge): the kernel is set up while its predecessor drains and waits for it in its first instruction
// (pdl_prologue): +1.2 % at N = 1M, +1.4 % at N = 10k on B200.  Anything else enqueued on the main stream
// (memset, event wait, collective) calls pdl_break().
template <typename... KArgs, typename... Args>
static inline void launch_k(fitsne_ctx *c, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at;
    memset(&at, 0, sizeof at);
    at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at.val.programmaticStreamSerializationAllowed = 1;
    const bool on = c->use_pdl && c->in_chain && c->prev_is_kernel && st == c->stream && !c->timing_this_iter;
    cfg.attrs = &at; cfg.numAttrs = on ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
    if (st == c->stream) c->prev_is_kernel = true;
}
#define LK(c, kern, grid, block, smem, st, ...) launch_k((c), kern, dim3(grid), dim3(block), (size_t) (smem), (st), __VA_ARGS__)
static inline void pdl_break(fitsne_ctx *c) { c->prev_is_kernel = false; }

template <int D>
static int launch_bounds_only(fitsne_ctx *c, const float *Yin, float *Yout, int do_center) {
    // do_center == 1: closing kernel of an optimiser step (skipped, like the rest, when the grid check failed); the means
    // are in Scalars::mean (k_update).  The last block to finish combines the per-block bounds and publishes them.
    const GridParams *gate = do_center ? c->gp : nullptr;
    LK(c, k_center_bounds<D>, RED_BLOCKS, 256, 0, c->stream, Yin, Yout, c->N, do_center, c->bounds_partial, c->sc,
                                                          c->reordered ? c->orig_of : nullptr, c->reordered ? c->pos_of : nullptr, gate,
                                                          c->host_bounds_dev, c->tickets + 1);
    LAUNCH_CHECK();
    c->stats.kernel_launches += 1;
    return 0;
}

template <int D, int P>
static int launch_spread_gather_variant(fitsne_ctx *c, bool gather, const uint32_t *skeys, const uint32_t *sperm) {
    void *grid = D == 2 ? (void *) c->chg : (c->p2p ? (void *) c->grid1d : (void *) c->planes);      // spread target
    if (!gather) {
        const int nchunks = cdiv(c->nloc, c->chunk);
        LK(c, (k_spread_chunks<D, P>), cdiv(nchunks, SP2_THREADS), SP2_THREADS, (spread_smem_bytes<D, P>()), c->stream,
            c->sorted_u, skeys, c->nloc, c->gp, c->slots, c->gpart, grid, c->box_range, c->work, c->chunk);
    } else {
        LK(c, (k_gather<D, P>), cdiv(c->nloc, 256), 256, 0, c->stream, c->sorted_u, skeys, sperm, c->nloc, c->gp, c->sc,
                                                                   D == 2 ? (const void *) c->pot : (const void *) c->planes, c->frep);
    }
    LAUNCH_CHECK();
    c->stats.kernel_launches += 1;
    return 0;
}

// compile-time node counts for nterms 2..4 (2-D) / 2..5 (1-D): accumulators in registers; anything else: run-time loops
template <int D>
static int launch_spread_gather(fitsne_ctx *c, bool gather, const uint32_t *skeys, const uint32_t *sperm) {
    switch (c->cfg.nterms) {
        case 2: return launch_spread_gather_variant<D, 2>(c, gather, skeys, sperm);
        case 3: return launch_spread_gather_variant<D, 3>(c, gather, skeys, sperm);
        case 4: return launch_spread_gather_variant<D, 4>(c, gather, skeys, sperm);
        case 5: if constexpr (D == 1) return launch_spread_gather_variant<D, 5>(c, gather, skeys, sperm);
        // fall through
        default: return launch_spread_gather_variant<D, 0>(c, gather, skeys, sperm);
    }
}

static inline size_t tiles_smem_bytes(int D) {
    const size_t yt = D == 2 ? sizeof(float2) : sizeof(float);
    return (size_t) TILE_ROWS * yt + (size_t) TILE_ROWS * D * sizeof(long long) + (size_t) TILE_COLS * yt;
}

template <int D>
static int launch_attract(fitsne_ctx *c, cudaStream_t st) {
    const int rows = c->row_end - c->row_begin;
    const float inv_df = (float) (1.0 / c->cfg.df);
    if (c->use_tiles) {
        // accumulation: 32-bit fixed point scaled by the largest row sum of P -- native shared-memory integer atomics,
        // order-independent => bitwise repeatable
        LK(c, k_attract_tiles<D>, c->tg.nchunks, 1024, tiles_smem_bytes(D), st, c->Y, c->N, c->tg, c->tile_start, c->tile_pack, c->tile_val,
                                                                              inv_df, c->tile_fix32, c->attr);
        LAUNCH_CHECK();
        c->stats.kernel_launches += 1;
        return 0;
    }
    // persistent grid: `per_sm` CTAs of 256 threads per SM (never more than the row groups there are)
    constexpr int per_sm = 8;          // B200 sweep (2..100000 CTAs per SM): a flat optimum from 6 up
    static const int lpr_env = getenv("FITSNE_LPR") ? atoi(getenv("FITSNE_LPR")) : 0;
#define ATT(L) LK(c, (k_attract<D, L>), std::min(cdiv((long long) rows * L, 256), 148 * per_sm), 256, 0, st,  \
        c->row_P, c->edges, c->edge_base, c->Y, c->row_begin, c->row_end, inv_df, c->attr)
    switch (lpr_env ? lpr_env : c->lpr) {
        case 4: ATT(4); break;
        case 8: ATT(8); break;
        case 16: ATT(16); break;
        default: ATT(32); break;
    }
#undef ATT
    LAUNCH_CHECK();
    c->stats.kernel_launches += 1;
    return 0;
}

// Everything from "bounds are known" to either dC (update=false) or the centred new Y and its bounds
// (update=true).  B is only passed to k_setup_grid (which verifies it against the device's own bounds); every
// launch shape below depends on M, the shard size and nterms only.  Pure work on ONE stream: capturable, sharded or not.
// (Round 1 forked the SpMV onto a second stream and round 2 first did the same with the kernel spectra.  Measured on
// B200: the SpMV saturates every SM's load/store pipe, so kernels running beside it crawl and the iteration takes the
// SUM of the kernel times either way -- 0.4446 ms with three streams, 0.4518 ms with one at N = 1M -- while in sharded
// runs the extra streams starved NCCL: 4.08 ms per iteration against 2.32 ms on one stream at N = 10M on two GPUs.)
template <int D>
static int enqueue_iteration(fitsne_ctx *c, const int *B_dev_arg, int M, bool update) {
    const int p = c->cfg.nterms, nloc = c->nloc;
    cudaStream_t st = c->stream;
    Plans *pl;
    CKRC(get_plans(c, M, &pl));

    c->in_chain = true;
    pdl_break(c);
    phase_mark(c, FITSNE_PHASE_BOUNDS);
    kt(c, "(start)");
    LK(c, k_setup_grid, 1, 256, 0, st, c->gp, c->sc, B_dev_arg, M, p, D, c->cfg.intervals_per_integer, c->cfg.min_num_intervals,
                                    c->mismatch, c->sort_totals, c->work, c->tickets + 5, c->p2p ? c->comm_seq : nullptr);
    c->stats.kernel_launches += 1;
    kt(c, "k_setup_grid");
    // Sharded, peer fabric: push my slice of Y (centred by the previous step) into every peer's Y with the copy engines on
    // a side stream -- DMA over NVLink, no SM involved -- while this stream sorts, spreads and convolves; the SpMV is the
    // only consumer of foreign rows and waits for the peers' flags.  (Timers mode: same operations, in line.)
    const bool push_Y = c->world > 1 && c->p2p;
    if (push_Y) {
        cudaStream_t cs = c->timing_this_iter ? st : c->stream_c;
        if (cs != st) { CK(cudaEventRecord(c->ev_cfork, st)); CK(cudaStreamWaitEvent(cs, c->ev_cfork, 0)); }
        const size_t off = (size_t) c->row_begin * D, bytes = (size_t) c->nloc * D * sizeof(float);
        for (int k = 1; k < c->world; k++) {
            const int r = (c->rank + k) % c->world;            // staggered targets: no two ranks hit the same peer first
            CK(cudaMemcpyAsync(c->pc.Y[r] + off, c->Y + off, bytes, cudaMemcpyDeviceToDevice, cs));
        }
        LK(c, k_peer_signal, 1, 32, 0, cs, c->pc, FLAG_Y);
        if (cs != st) CK(cudaEventRecord(c->ev_cjoin, cs));
        c->stats.kernel_launches += 1;
    }
    // column CTAs: 256 threads (~4 CTAs share an SM on one GPU).  FITSNE_COL_THREADS=512 selects the wide instantiation for
    // experiments with the distributed convolution, where a rank's few columns have an SM each: on 2 x B200 (289 columns
    // per rank) it shortened the serialised convolution phase 0.160 -> 0.141 ms but not the graph-replayed iteration
    // (0.348 -> 0.356 ms), so 256 stays the default.
    static const int col_env = getenv("FITSNE_COL_THREADS") ? atoi(getenv("FITSNE_COL_THREADS")) : 0;
    const int col_threads = col_env >= 64 && col_env <= COL_THREADS_MAX ? col_env / 32 * 32 : COL_THREADS;
    auto launch_kernel_side = [&](cudaStream_t ks) -> int {
        const int Gc = M / 2, H = M / 2 + 1;
        LK(c, k_kspec_rows, Gc, ROW_THREADS, pl->smem_row1, ks, c->KR, pl->plan, pl->W, c->gp, c->cfg.df);
#define COL_LAUNCH(KERN, GRID, STREAM, ...) do { \
        if (col_threads <= COL_THREADS) LK(c, KERN<COL_THREADS>, GRID, col_threads, pl->smem_col, STREAM, __VA_ARGS__); \
        else LK(c, KERN<COL_THREADS_MAX>, GRID, col_threads, pl->smem_col, STREAM, __VA_ARGS__); } while (0)
        COL_LAUNCH(k_kspec_cols, (H + 1) / 2, ks, c->KR, c->KS, pl->cplan, pl->W, c->gp, c->rank, c->p2p && c->dist_conv ? c->world : 1);
        LAUNCH_CHECK();
        c->stats.kernel_launches += 2;
        return 0;
    };
    const bool kside = D == 2 && c->world > 1 && c->p2p && !c->timing_this_iter && c->stream_k != nullptr;
    if (kside) {
        CK(cudaEventRecord(c->ev_kfork, st));
        CK(cudaStreamWaitEvent(c->stream_k, c->ev_kfork, 0));
        CKRC(launch_kernel_side(c->stream_k));
        CK(cudaEventRecord(c->ev_kjoin, c->stream_k));
    }
    // ---- bin + stable two-pass LSD radix sort by box (three launches; see fitsne_kernels.cuh)
    phase_mark(c, FITSNE_PHASE_SORT);
    const int tiles = cdiv(nloc, SORT_TILE);
    const int max_bins = 1 << SORT_MAX_BITS;
    const size_t sweep_smem = (size_t) max_bins * 4 + (size_t) (SWEEP_THREADS / 32) * max_bins * 2;
    // in-box coordinates travel with the keys: k_bin -> frep (free until the gather) -> sweep#0 -> dC (free until the
    // update) -> sweep#1 -> sorted_u.  One-pass layout: k_bin writes keys[1] and the coordinates straight to dC, sweep#0
    // returns at once.
    float *u0 = c->frep, *u1 = c->dC;
    LK(c, k_bin<D>, tiles, BIN_THREADS, 0, st, c->Y, c->row_begin, nloc, c->gp, c->keys[0], c->keys[1], u0, u1, c->sort_totals, c->sort_bases,
                                           c->sweep_state, tiles, c->tickets + 4);
    kt(c, "k_bin");
    LK(c, k_radix_sweep, tiles, SWEEP_THREADS, sweep_smem, st, c->keys[0], nullptr, c->keys[1], c->perm[1], nloc, 0, c->sort_bases, c->sweep_state,
                                                           tiles, c->tickets + 5, (uint32_t) c->row_begin, c->gp, u0, u1, D);
    kt(c, "k_radix_sweep#0");
    LK(c, k_radix_sweep, tiles, SWEEP_THREADS, sweep_smem, st, c->keys[1], c->perm[1], c->keys[0], c->perm[0], nloc, 1, c->sort_bases, c->sweep_state,
                                                           tiles, c->tickets + 6, (uint32_t) c->row_begin, c->gp, u1, c->sorted_u, D);
    const uint32_t *skeys = c->keys[0], *sperm = c->perm[0];
    kt(c, "k_radix_sweep#1");
    c->stats.kernel_launches += 3;
    LAUNCH_CHECK();

    // ---- spread: the grid is cleared, chunks write finished boxes + partial slots, the listed multi-chunk boxes are combined
    phase_mark(c, FITSNE_PHASE_SPREAD);
    const int Gc = M / 2;
    const size_t cplane = D == 2 ? (size_t) Gc * Gc : (size_t) M;            // grid elements a length-M FFT can hold
    void *spread_grid = D == 2 ? (void *) c->chg : (c->p2p ? (void *) c->grid1d : (void *) c->planes);
    if (D == 2) CK(cudaMemsetAsync(c->chg, 0, cplane * sizeof(float4), st));
    else CK(cudaMemsetAsync(spread_grid, 0, (size_t) 2 * M * sizeof(float2), st));
    pdl_break(c);
    CKRC(launch_spread_gather<D>(c, false, skeys, sperm));
    kt(c, "k_spread_chunks");
    // (sharded, peer fabric: the CTA of the combine that finishes last announces "my partial grid is complete" to the peers)
    LK(c, k_spread_combine<D>, 148 * 8, 256, 0, st, c->slots, c->box_range, c->gp, c->work, spread_grid, c->tickets + 7, c->pc,
                                                 c->world > 1 && c->p2p ? 1 : 0, c->chunk);
    c->stats.kernel_launches += 1;
    if (c->world > 1) {
        // every rank spread its own points: sum the partial grids (fp32).  2-D: the dense (M/2)^2 float4 region that holds
        // the G x G grid; 1-D: the two packed charge lines.  The element count depends on M only, like every launch shape.
        phase_mark(c, FITSNE_PHASE_COLLECTIVES);
        if (c->p2p) {
            // peer fabric: the sum over ranks happens inside k_conv_rows_fwd's loads (2-D) or in one small kernel (1-D)
            if (D == 1) {
                LK(c, k_grid_sum_1d, cdiv(2 * M, 256), 256, 0, st, c->pc, c->planes, 2 * M, &c->gp->ok);
                c->stats.kernel_launches += 1;
            }
        } else if (D == 2) { CKNCCL(g_nccl.AllReduce(c->chg, c->chg, cplane * 4, ncclFloat, ncclSum, c->comm, st)); pdl_break(c); }
        else { CKNCCL(g_nccl.AllReduce(c->planes, c->planes, (size_t) M * 4, ncclFloat, ncclSum, c->comm, st)); pdl_break(c); }
    }
    kt(c, "k_spread_combine(+collective)");

    // ---- convolution (+ sum_Q)
    phase_mark(c, FITSNE_PHASE_KERNEL_SPECTRUM);
    const int *gok = &c->gp->ok;
    if (D == 2) {
        // the kernel spectra depend on the grid geometry only.  Single GPU: in line (beside the SpMV-saturated kernels a second
        // stream buys nothing, see above).  Sharded runs leave most of every GPU idle between exchanges: there the two kernels
        // run on a side stream beside the sort and the spread (forked right after k_setup_grid, joined here).
        if (kside) { CK(cudaStreamWaitEvent(st, c->ev_kjoin, 0)); pdl_break(c); }
        else CKRC(launch_kernel_side(st));
        kt(c, "k_kspec_rows + k_kspec_cols");
        phase_mark(c, FITSNE_PHASE_FFT);
        const int H = M / 2 + 1;
        LK(c, k_conv_rows_fwd, Gc, ROW_THREADS, pl->smem_row2, st, c->chg, c->S, pl->plan, pl->W, c->gp, c->pc, c->p2p ? (c->dist_conv ? 2 : 1) : 0,
                                                                        c->tickets + 8);
        kt(c, "k_conv_rows_fwd");
        const int dist = c->p2p && c->dist_conv ? 1 : 0, p2p = dist ? 2 : 0;
        // sharded, distributed convolution: like a 2-D FFT -- rows and spectrum columns dealt out in blocks; every
        // "transpose" is the producing kernel's stores on peer memory; the CTA that finishes last raises the stage's flag at
        // the peers, the consuming kernel's CTAs wait for it themselves -- no launch of its own for any exchange
        COL_LAUNCH(k_conv_cols, H, st, pl->tmS, c->KS, pl->cplan, pl->W, c->gp, c->df_is_one ? 1 : 0, c->zpartial, c->N, c->sc, c->tickets + 0,
                   c->pc, p2p);
        kt(c, "k_conv_cols");
        LK(c, k_conv_rows_inv, Gc, ROW_THREADS, pl->smem_row2, st, c->S, c->pot, pl->plan, pl->W, c->gp, c->pc, p2p, c->N, c->sc, c->tickets + 9);
        kt(c, "k_conv_rows_inv");
        c->stats.kernel_launches += 3;
    } else {
        LK(c, k_gen_kernels_1d, cdiv(M, 256), 256, 0, st, c->gp, c->cfg.df, c->planes);
        kt(c, "k_gen_kernels_1d");
        phase_mark(c, FITSNE_PHASE_FFT);
        // lines 0,1 = charges (zero beyond G: substituted while loading), lines 2,3 = kernels
        LK(c, k_fft_line, 4, FFT_THREADS, pl->smem_line, st, c->planes, pl->plan, pl->W, 0, 0x0u, &c->gp->G, gok);
        LK(c, k_hadamard_1d, Z_BLOCKS_1D, 256, 0, st, c->planes, c->gp, c->df_is_one ? 1 : 0, c->zpartial, c->N, c->sc, c->tickets + 0);
        LK(c, k_fft_line, 1, FFT_THREADS, pl->smem_line, st, c->planes, pl->plan, pl->W, 1, 0u, &c->gp->G, gok);
        kt(c, "1-D convolution");
        c->stats.kernel_launches += 4;
    }
    LAUNCH_CHECK();

    // ---- gather (+ 1/Z)
    phase_mark(c, FITSNE_PHASE_GATHER);
    const bool dist_conv_on = D == 2 && c->p2p && c->dist_conv;
    if (dist_conv_on) {
        // distributed convolution: the potential rows come from every rank -- one wait covers them AND the Y slices the
        // SpMV needs further down (pushed at the start of the iteration: long since there)
        if (push_Y && !c->timing_this_iter) { CK(cudaStreamWaitEvent(st, c->ev_cjoin, 0)); pdl_break(c); }
        LK(c, k_peer_wait, 1, 32, 0, st, c->pc, FLAG_POT, push_Y ? FLAG_Y : -1);
        c->stats.kernel_launches += 1;
    }
    CKRC(launch_spread_gather<D>(c, true, skeys, sperm));
    kt(c, "k_gather");

    // ---- attractive term + optimiser step.  Sharded: after an optimiser step every rank only holds ITS slice of the new Y;
    // the SpMV is the only consumer of foreign rows inside the iteration (bin / sort / spread / gather / update read the
    // local slice), so the all-gather that completes Y sits right in front of it.
    if (c->world > 1) {
        phase_mark(c, FITSNE_PHASE_ALLGATHER);
        if (push_Y) {
            if (!dist_conv_on) {
                if (!c->timing_this_iter) { CK(cudaStreamWaitEvent(st, c->ev_cjoin, 0)); pdl_break(c); }     // my own pushes are out ...
                LK(c, k_peer_wait, 1, 32, 0, st, c->pc, FLAG_Y, -1);                           // ... and everybody's have landed here
                c->stats.kernel_launches += 1;
            }
        } else { CKNCCL(g_nccl.AllGather(c->Y + (size_t) c->rank * c->per * D, c->Y, (size_t) c->per * D, ncclFloat, c->comm, st)); pdl_break(c); }
    }
    c->y_whole = true;
    phase_mark(c, FITSNE_PHASE_ATTRACT_UPDATE);
    CKRC(launch_attract<D>(c, st));
    kt(c, "k_attract");
    const int rows = c->row_end - c->row_begin;
    const int ublocks = std::min(RED_BLOCKS, cdiv(rows, 256));
    if (!update) {
        LK(c, (k_update<D, false>), ublocks, 256, 0, st, c->Y, c->attr, c->frep, c->row_begin, c->row_end, c->sp, c->gp, c->dC,
                                                   c->uY, c->gains, c->Yb, nullptr, c->N, c->sc, c->tickets + 2);
        c->stats.kernel_launches += 1;
        phase_mark(c, FITSNE_PHASE_CENTER);
        if (c->p2p) {
            // no statistics exchange closes a gradient-only pass: meet the peers explicitly, so that nobody clears its
            // partial grid for the next pass while a slower rank still reads it
            LK(c, k_peer_signal, 1, 32, 0, st, c->pc, FLAG_STATS);
            LK(c, k_peer_wait, 1, 32, 0, st, c->pc, FLAG_STATS, -1);
            c->stats.kernel_launches += 2;
        }
    } else if (c->world == 1) {
        // single GPU: k_update also produces the column means of the new positions (per-CTA register sums, last-block
        // reduction), the centring kernel subtracts them, finds the bounds and publishes them -- two launches for the tail
        LK(c, (k_update<D, true>), ublocks, 256, 0, st, c->Y, c->attr, c->frep, c->row_begin, c->row_end, c->sp, c->gp, c->dC,
                                                  c->uY, c->gains, c->Yb, c->colsum_partial, c->N, c->sc, c->tickets + 2);
        c->stats.kernel_launches += 1;
        phase_mark(c, FITSNE_PHASE_CENTER);
        CKRC(launch_bounds_only<D>(c, c->Yb, c->Y, 1));
    } else {
        // sharded tail: local update with its sums / bounds riding along -> 128-byte records exchanged -> every rank centres
        // its own slice with the global mean and publishes the (identical) global bounds.  Y itself is gathered next iteration.
        LK(c, k_update_shard<D>, ublocks, 256, 0, st, c->Y, c->attr, c->frep, c->row_begin, c->row_end, c->rank, c->sp, c->gp, c->dC, c->uY,
                                                   c->gains, c->Yb, c->shard_sum_partial, c->shard_mm_partial, c->shard_stats + c->rank,
                                                   c->tickets + 3, c->pc, c->p2p ? 1 : 0, c->reordered ? c->orig_of : nullptr,
                                                   c->reordered ? c->pos_of : nullptr);
        phase_mark(c, FITSNE_PHASE_CENTER);
        if (!c->p2p) { CKNCCL(g_nccl.AllGather(c->shard_stats + c->rank, c->shard_stats, sizeof(ShardStats), ncclChar, c->comm, st)); pdl_break(c); }
        LK(c, k_center_shard<D>, cdiv(rows, 256), 256, 0, st, c->Yb, c->Y, c->row_begin, c->row_end, c->N, c->shard_stats, c->world,
                                                          c->gp, c->sc, c->host_bounds_dev, c->pc, c->p2p ? 1 : 0);
        c->stats.kernel_launches += 2;
        c->y_whole = false;
    }
    phase_mark(c, FITSNE_PHASE_COUNT);
    c->in_chain = false;
    LAUNCH_CHECK();
    kt(c, "k_update + zero-mean/bounds tail");
    return 0;
}

static int enqueue_iteration_d(fitsne_ctx *c, int M, bool update) {
    // the host's n_boxes choice travels through a mapped pinned word so the captured graph stays valid
    return c->D == 2 ? enqueue_iteration<2>(c, c->host_B_dev, M, update) : enqueue_iteration<1>(c, c->host_B_dev, M, update);
}

// Sharded: make c->Y hold every rank's slice again (after a step only the local slice is current; the next iteration's
// own exchange (peer pushes / all-gather) does this, anything else that reads foreign rows -- KL, downloads, a bounds scan --
// calls this first).  Collective: every rank reaches it at the same point of the call sequence.
static int ensure_whole_Y(fitsne_ctx *c) {
    if (c->world == 1 || c->y_whole) return 0;
    CKNCCL(g_nccl.AllGather(c->Y + (size_t) c->rank * c->per * c->D, c->Y, (size_t) c->per * c->D, ncclFloat, c->comm, c->stream));
    c->y_whole = true;
    return 0;
}

static int refresh_bounds(fitsne_ctx *c) {
    if (c->bounds_valid) return 0;
    CKRC(ensure_whole_Y(c));
    if (c->D == 2) CKRC(launch_bounds_only<2>(c, c->Y, c->Y, 0)); else CKRC(launch_bounds_only<1>(c, c->Y, c->Y, 0));
    c->bounds_valid = true;
    return 0;
}

static int push_step_params(fitsne_ctx *c, const StepParams &sp) {
    if (c->sp_valid && memcmp(&sp, &c->sp_host, sizeof sp) == 0) return 0;
    c->sp_host = sp;
    CK(cudaMemcpyAsync(c->sp, &c->sp_host, sizeof sp, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));   // sp_host is reused; rare (only when the schedule changes)
    c->sp_valid = true;
    return 0;
}

// ------------------------------------------------------------------------------ locality re-ordering --
// Physically re-order the points along a Morton curve of the current embedding, re-label the CSR and regroup
// its edges into (row chunk x column block) tiles for k_attract_tiles.  Rare (iterations 0, 50, 150, 350, ...),
// so it favours simplicity: a handful of streaming kernels, ~E*40 bytes of traffic.
static int reorder_points(fitsne_ctx *c) {
    const int N = c->N, D = c->D;
    cudaStream_t st = c->stream;
    const size_t E = c->E;
    if (!c->orig_of) {
        CKRC(dev_alloc(c, &c->orig_of, (size_t) N)); CKRC(dev_alloc(c, &c->orig_tmp, (size_t) N));
        CKRC(dev_alloc(c, &c->pos_of, (size_t) N)); CKRC(dev_alloc(c, &c->rank_map, (size_t) N));
        CKRC(dev_alloc(c, &c->row_P2, (size_t) N + 1)); CKRC(dev_alloc(c, &c->edges2, E + 1));
        CKRC(dev_alloc(c, &c->tile_pack, E + 1)); CKRC(dev_alloc(c, &c->tile_val, E + 1));
        const int w = std::max(1, cdiv(N, 148 * TILE_ROWS));            // waves of 148 row chunks
        c->tg.rows_per_chunk = std::min(TILE_ROWS, std::max(64, cdiv(N, 148 * w)));
        c->tg.nchunks = cdiv(N, c->tg.rows_per_chunk);
        c->tg.ncb = cdiv(N, TILE_COLS);
        c->ntiles = (size_t) c->tg.nchunks * c->tg.ncb;
        CKRC(dev_alloc(c, &c->tile_cnt, c->ntiles + 1)); CKRC(dev_alloc(c, &c->tile_start, c->ntiles + 1));
        CKRC(dev_alloc(c, &c->tile_cur, c->ntiles + 1)); CKRC(dev_alloc(c, &c->nonempty, (size_t) 1));
        CKRC(dev_alloc(c, &c->gp_reorder, (size_t) 1));
        GridParams g;
        memset(&g, 0, sizeof g);
        g.ok = 1; g.sort_bits = 11; g.sort_passes = 2;      // 22-bit locality keys: two passes of 11 bits
        CK(cudaMemcpyAsync(c->gp_reorder, &g, sizeof g, cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));
        {   // fixed-point scale of the tiled kernel: 2^30 / max row sum
            const int blocks = 1024;
            k_row_sum_max<<<blocks, 256, 0, st>>>(c->row_P, c->edges, c->edge_base, c->row_begin, c->row_end, c->kl_partial);
            std::vector<double> h(blocks);
            CK(cudaMemcpyAsync(h.data(), c->kl_partial, blocks * sizeof(double), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            double mx = 0;
            for (double v : h) mx = std::max(mx, v);
            // |q dx| / p <= sqrt(df) / 2 for the kernel (1 + d^2/df)^-1, so a row sum stays below rowsum * max(1, sqrt(df)) / 2
            c->tile_fix32 = (float) (1073741824.0 / (std::max(mx, 1e-300) * std::max(1.0, std::sqrt(c->cfg.df))));
        }
        if (D == 2) CK(cudaFuncSetAttribute(k_attract_tiles<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) tiles_smem_bytes(2)));
        else CK(cudaFuncSetAttribute(k_attract_tiles<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) tiles_smem_bytes(1)));
    }
    CKRC(refresh_bounds(c));     // sc->bmin / bmax of the current Y
    // 1. locality keys, stable 2 x 11-bit LSD sort (same kernels as the per-iteration box sort)
    const int tiles = cdiv(N, SORT_TILE);
    const int max_bins = 1 << SORT_MAX_BITS;
    const size_t scatter_smem = (size_t) max_bins * 4 + (size_t) (SORT_THREADS / 32) * max_bins * 2;
    if (D == 2) k_morton_keys<2><<<cdiv(N, 256), 256, 0, st>>>(c->Y, N, c->sc, c->keys[0]);
    else k_morton_keys<1><<<cdiv(N, 256), 256, 0, st>>>(c->Y, N, c->sc, c->keys[0]);
    CK(cudaMemsetAsync(c->sort_totals, 0, sizeof(uint32_t) * 2 * max_bins, st));
    k_radix_hist<<<tiles, SORT_THREADS, 0, st>>>(c->keys[0], N, 0, c->hist, tiles, c->sort_totals, c->gp_reorder);
    k_radix_offsets<<<max_bins, 256, 0, st>>>(c->hist, tiles, c->sort_totals, 0, c->gp_reorder);
    k_radix_scatter<<<tiles, SORT_THREADS, scatter_smem, st>>>(c->keys[0], nullptr, c->keys[1], c->perm[1], N, 0, c->hist, tiles, 0u, c->gp_reorder);
    k_radix_hist<<<tiles, SORT_THREADS, 0, st>>>(c->keys[1], N, 1, c->hist, tiles, c->sort_totals + max_bins, c->gp_reorder);
    k_radix_offsets<<<max_bins, 256, 0, st>>>(c->hist, tiles, c->sort_totals + max_bins, 1, c->gp_reorder);
    k_radix_scatter<<<tiles, SORT_THREADS, scatter_smem, st>>>(c->keys[1], c->perm[1], c->keys[0], c->perm[0], N, 1, c->hist, tiles, 0u, c->gp_reorder);
    LAUNCH_CHECK();
    // 2. maps
    k_reorder_maps<<<cdiv(N, 256), 256, 0, st>>>(c->perm[0], N, c->rank_map, c->reordered ? c->orig_of : nullptr, c->orig_tmp, c->pos_of);
    CK(cudaMemcpyAsync(c->orig_of, c->orig_tmp, (size_t) N * 4, cudaMemcpyDeviceToDevice, st));
    // 3. per-point state
    const size_t bytes = (size_t) N * D * sizeof(float);
    float *state[3] = {c->Y, c->uY, c->gains};
    for (float *buf : state) {
        if (D == 2) k_permute_rows<2><<<cdiv(N, 256), 256, 0, st>>>(buf, c->Yb, c->perm[0], N);
        else k_permute_rows<1><<<cdiv(N, 256), 256, 0, st>>>(buf, c->Yb, c->perm[0], N);
        CK(cudaMemcpyAsync(buf, c->Yb, bytes, cudaMemcpyDeviceToDevice, st));
    }
    // 4. CSR relabel + tiles (keys[1] is free again: row lengths)
    uint32_t *new_len = c->keys[1];
    CK(cudaMemsetAsync(c->tile_cnt, 0, (c->ntiles + 1) * 4, st));
    CK(cudaMemsetAsync(c->tile_cur, 0, (c->ntiles + 1) * 4, st));
    CK(cudaMemsetAsync(c->nonempty, 0, 4, st));
    k_relabel_csr<<<cdiv((long long) N * 8, 256), 256, 0, st>>>(0, c->row_P, c->edges, c->rank_map, N, c->tg, new_len, c->tile_cnt,
                                                                 nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    k_scan_excl<<<1, 1024, 0, st>>>(new_len, c->row_P2, N);
    k_scan_excl<<<1, 1024, 0, st>>>(c->tile_cnt, c->tile_start, (int) c->ntiles);
    k_count_nonempty<<<cdiv(c->ntiles, 256), 256, 0, st>>>(c->tile_cnt, c->ntiles, c->nonempty);
    LAUNCH_CHECK();
    CK(cudaMemcpyAsync(&c->nonempty_tiles, c->nonempty, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    // 5. which attractive kernel: modelled cost of the tiled path (column-block fills from L2 + the edge stream) vs
    //    the CSR gather path (~2 cycles per divergent gather per SM; measured 6.9 ps/edge chip-wide on B200)
    // (calibrated on B200, N=1M/E=30M: 4393 non-empty tiles -> 340 us without accumulation => ~78 ns per column-block fill)
    const double est_tiles_us = (double) c->nonempty_tiles * 0.078 + (double) E * 8.0 / 5.0e6 + 5.0;
    const double est_csr_us = (double) E * 6.9e-6 + 5.0;
    const bool was_reordered = c->reordered, old_tiles = c->use_tiles;
    c->use_tiles = est_tiles_us < est_csr_us;
    if (c->cfg.flags & FITSNE_FLAG_FORCE_TILES) c->use_tiles = true;
    if (c->cfg.flags & FITSNE_FLAG_NO_TILES) c->use_tiles = false;
    c->reorders++;
    TRACE("reorder #%llu: %u of %zu tiles non-empty, est tiles %.0f us vs csr %.0f us -> %s", (unsigned long long) c->reorders,
          c->nonempty_tiles, c->ntiles, est_tiles_us, est_csr_us, c->use_tiles ? "tiles" : "csr");
    // second pass: the relabelled CSR, and -- only if the tiled kernel will run -- the edges scattered into their tiles
    // (an atomic and 8 scattered bytes per edge that the CSR path never reads)
    k_relabel_csr<<<cdiv((long long) N * 8, 256), 256, 0, st>>>(1, c->row_P, c->edges, c->rank_map, N, c->tg, new_len, c->tile_cnt,
                                                                 c->row_P2, c->edges2, c->tile_start, c->tile_cur,
                                                                 c->use_tiles ? c->tile_pack : nullptr, c->tile_val);
    LAUNCH_CHECK();
    if (was_reordered && c->use_tiles == old_tiles) {
        // a later re-ordering: the new CSR goes back into the buffers the captured graphs point at -- a 250 MB device copy
        // (~80 us at N = 1M) instead of capturing and instantiating every graph again
        CK(cudaMemcpyAsync(c->row_P, c->row_P2, ((size_t) N + 1) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(c->edges, c->edges2, E * sizeof(uint2), cudaMemcpyDeviceToDevice, st));
        CK(cudaStreamSynchronize(st));
    } else {
        CK(cudaStreamSynchronize(st));
        std::swap(c->row_P, c->row_P2); std::swap(c->edges, c->edges2);
        drop_graphs(c);          // CSR pointers (first re-ordering: also the order maps) or the attractive kernel changed
    }
    c->reordered = true;
    c->kernel_launches_reorder += 20;
    return 0;
}

static int ensure_whole_Y(fitsne_ctx *c);

// Sharded contexts: every rank Morton-orders the points of ITS OWN slice (block-diagonal permutation: row ownership does not
// change, so no CSR row moves between ranks), the maps are all-gathered, the local CSR rows are re-labelled through the
// global map.  Gives the sort / spread / gather their locality back and turns the SpMV's neighbour gathers into `world`
// Morton-ordered windows instead of uniformly random reads.  Collective: every rank calls it at the same step count.
static int reorder_points_sharded(fitsne_ctx *c) {
    const int N = c->N, D = c->D, nloc = c->nloc, b = c->row_begin;
    cudaStream_t st = c->stream;
    const size_t E = c->E, padded = (size_t) c->per * c->world;
    if (!c->orig_of) {
        CKRC(dev_alloc(c, &c->orig_of, padded)); CKRC(dev_alloc(c, &c->orig_tmp, padded));
        CKRC(dev_alloc(c, &c->pos_of, padded)); CKRC(dev_alloc(c, &c->rank_map, padded));
        CKRC(dev_alloc(c, &c->row_P2, (size_t) N + 1)); CKRC(dev_alloc(c, &c->edges2, E + 1));
        CKRC(dev_alloc(c, &c->gp_reorder, (size_t) 1));
        GridParams g;
        memset(&g, 0, sizeof g);
        g.ok = 1; g.sort_bits = 11; g.sort_passes = 2;
        CK(cudaMemcpyAsync(c->gp_reorder, &g, sizeof g, cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));
    }
    CKRC(refresh_bounds(c));     // whole Y, identical bounds on every rank
    const int tiles = cdiv(nloc, SORT_TILE);
    const int max_bins = 1 << SORT_MAX_BITS;
    const size_t scatter_smem = (size_t) max_bins * 4 + (size_t) (SORT_THREADS / 32) * max_bins * 2;
    if (D == 2) k_morton_keys<2><<<cdiv(nloc, 256), 256, 0, st>>>(c->Y + (size_t) b * D, nloc, c->sc, c->keys[0]);
    else k_morton_keys<1><<<cdiv(nloc, 256), 256, 0, st>>>(c->Y + (size_t) b * D, nloc, c->sc, c->keys[0]);
    CK(cudaMemsetAsync(c->sort_totals, 0, sizeof(uint32_t) * 2 * max_bins, st));
    k_radix_hist<<<tiles, SORT_THREADS, 0, st>>>(c->keys[0], nloc, 0, c->hist, tiles, c->sort_totals, c->gp_reorder);
    k_radix_offsets<<<max_bins, 256, 0, st>>>(c->hist, tiles, c->sort_totals, 0, c->gp_reorder);
    k_radix_scatter<<<tiles, SORT_THREADS, scatter_smem, st>>>(c->keys[0], nullptr, c->keys[1], c->perm[1], nloc, 0, c->hist, tiles, 0u, c->gp_reorder);
    k_radix_hist<<<tiles, SORT_THREADS, 0, st>>>(c->keys[1], nloc, 1, c->hist, tiles, c->sort_totals + max_bins, c->gp_reorder);
    k_radix_offsets<<<max_bins, 256, 0, st>>>(c->hist, tiles, c->sort_totals + max_bins, 1, c->gp_reorder);
    k_radix_scatter<<<tiles, SORT_THREADS, scatter_smem, st>>>(c->keys[1], c->perm[1], c->keys[0], c->perm[0], nloc, 1, c->hist, tiles, 0u, c->gp_reorder);
    // maps of my slice, then everybody's
    k_reorder_maps_local<<<cdiv(nloc, 256), 256, 0, st>>>(c->perm[0], nloc, (uint32_t) b, c->rank_map, c->reordered ? c->orig_of : nullptr,
                                                          c->orig_tmp, c->pos_of);
    LAUNCH_CHECK();
    uint32_t *maps[3] = {c->rank_map, c->orig_tmp, c->pos_of};
    for (uint32_t *m : maps) CKNCCL(g_nccl.AllGather(m + (size_t) c->rank * c->per, m, (size_t) c->per, ncclUint32, c->comm, st));
    CK(cudaMemcpyAsync(c->orig_of, c->orig_tmp, padded * 4, cudaMemcpyDeviceToDevice, st));
    // per-point state of my rows; Y is completed again below
    const size_t bytes = (size_t) nloc * D * sizeof(float);
    float *state[3] = {c->Y, c->uY, c->gains};
    for (float *buf : state) {
        if (D == 2) k_permute_rows<2><<<cdiv(nloc, 256), 256, 0, st>>>(buf + (size_t) b * D, c->Yb, c->perm[0], nloc);
        else k_permute_rows<1><<<cdiv(nloc, 256), 256, 0, st>>>(buf + (size_t) b * D, c->Yb, c->perm[0], nloc);
        CK(cudaMemcpyAsync(buf + (size_t) b * D, c->Yb, bytes, cudaMemcpyDeviceToDevice, st));
    }
    // my CSR rows
    uint32_t *new_len = c->keys[1], *row_new = c->keys[0];          // [nloc] / [nloc + 1] (keys are allocated with head room)
    k_relabel_csr_local<<<cdiv((long long) nloc * 8, 256), 256, 0, st>>>(0, c->row_P, c->edge_base, c->edges, c->rank_map, b, nloc, new_len, nullptr, nullptr);
    k_scan_excl<<<1, 1024, 0, st>>>(new_len, row_new, nloc);
    k_relabel_csr_local<<<cdiv((long long) nloc * 8, 256), 256, 0, st>>>(1, c->row_P, c->edge_base, c->edges, c->rank_map, b, nloc, new_len, row_new, c->edges2);
    CK(cudaMemcpyAsync(c->row_P2, c->row_P, ((size_t) N + 1) * 4, cudaMemcpyDeviceToDevice, st));
    k_local_row_offsets<<<cdiv(nloc + 1, 256), 256, 0, st>>>(row_new, nloc, b, c->edge_base, c->row_P2);
    LAUNCH_CHECK();
    if (c->reordered) {          // a later re-ordering: back into the buffers the captured graphs point at (see reorder_points)
        CK(cudaMemcpyAsync(c->row_P, c->row_P2, ((size_t) N + 1) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(c->edges, c->edges2, c->E * sizeof(uint2), cudaMemcpyDeviceToDevice, st));
        CK(cudaStreamSynchronize(st));
    } else {
        CK(cudaStreamSynchronize(st));
        std::swap(c->row_P, c->row_P2); std::swap(c->edges, c->edges2);
        drop_graphs(c);
    }
    c->reordered = true;
    c->reorders++;
    c->y_whole = false;          // the other ranks re-ordered their slices too
    CKRC(ensure_whole_Y(c));
    c->bounds_valid = true;      // a permutation does not move the bounds
    c->kernel_launches_reorder += 16;
    return 0;
}

static int maybe_reorder(fitsne_ctx *c) {
    if ((c->cfg.flags & FITSNE_FLAG_NO_REORDER) || c->E == 0) return 0;
    if (c->world > 1) {
        static const bool no_shard_reorder = getenv("FITSNE_NO_SHARD_REORDER") && atoi(getenv("FITSNE_NO_SHARD_REORDER")) != 0;
        if (no_shard_reorder) return 0;
        if (c->reordered && c->steps_total - c->last_reorder_iter < c->reorder_interval) return 0;
        if (c->reordered) c->reorder_interval = std::min<uint64_t>(c->reorder_interval * 2, 400);
        c->last_reorder_iter = c->steps_total;
        return reorder_points_sharded(c);
    }
    if (c->reordered && c->steps_total - c->last_reorder_iter < c->reorder_interval) return 0;
    if (c->reordered) c->reorder_interval = std::min<uint64_t>(c->reorder_interval * 2, 400);
    c->last_reorder_iter = c->steps_total;
    return reorder_points(c);
}

// The captured CUDA graph of one iteration for FFT length M (kind: gradient only / full step); created on first use.
static int get_graph(fitsne_ctx *c, int M, bool update, fitsne_ctx::GraphEntry **out) {
    const GraphKey key{M, update ? 1 : 0};
    auto it = c->graphs.find(key);
    if (it == c->graphs.end()) {
        Plans *pl;
        TRACE("new graph for M=%d kind=%d", M, update ? 1 : 0);
        CKRC(get_plans(c, M, &pl));   // twiddle tables are computed outside the capture
        const uint64_t launches_before = c->stats.kernel_launches;
        CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        const int rc = enqueue_iteration_d(c, M, update);
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
        if (rc != 0) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (e != cudaSuccess) return fail(c, FITSNE_ECUDA, "graph capture failed: %s", cudaGetErrorString(e));
        cudaGraphExec_t exec;
        CK(cudaGraphInstantiate(&exec, graph, 0));
        cudaGraphDestroy(graph);
        c->graphs[key] = fitsne_ctx::GraphEntry{exec, c->stats.kernel_launches - launches_before};
        it = c->graphs.find(key);
        c->stats.kernel_launches = launches_before;
    }
    *out = &it->second;
    return 0;
}

// Host-side grid choice from the last published bounds (the stream must be idle).
static int choose_grid(fitsne_ctx *c, int *B_out, int *M_out) {
    const double mn = (double) c->host_bounds[0], mx = (double) c->host_bounds[1];
    if (!(mx > mn)) return fail(c, FITSNE_EINVAL, "degenerate embedding: max_coord (%g) <= min_coord (%g)", mx, mn);
    const int B = choose_n_boxes(mn, mx, c->cfg.intervals_per_integer, c->cfg.min_num_intervals, c->D);
    const int G = B * c->cfg.nterms;
    if (sort_bits_for(B, c->D) > SORT_MAX_BITS || (long long) G * 2 > 65536)   // two passes of <= 11 bits
        return fail(c, FITSNE_EINVAL, "grid too large: n_boxes=%d", B);
    const int M = nice_fft_size(2 * G);
    if (M > max_fft_len(c->D)) return fail(c, FITSNE_EINVAL, "grid too large: n_boxes=%d needs FFT length %d > %d", B, M, max_fft_len(c->D));
    CKRC(ensure_grid_capacity(c, M));
    if (B != c->cur_B || M != c->cur_M) { c->cur_B = B; c->cur_M = M; c->stats.regrids++; }
    c->stats.n_boxes = B; c->stats.grid_side = G; c->stats.fft_side = M;
    c->stats.min_coord = mn; c->stats.max_coord = mx;
    *B_out = B; *M_out = M;
    return 0;
}

static int run_iteration(fitsne_ctx *c, bool update);

// Speculative batch: enqueue up to `n` full optimiser steps back to back with NO host round trip in between.  Each
// step sizes its own grid on the device (k_setup_grid); a step whose grid no longer maps to this graph's FFT length
// turns itself -- and, since nothing changed, every later step of the batch -- into a no-op.  One synchronisation at
// the end tells how many steps really ran (Scalars::iter_done, mirrored in mapped host memory).
static int run_batch(fitsne_ctx *c, int n, int *done) {
    CKRC(maybe_reorder(c));
    CKRC(refresh_bounds(c));
    CK(cudaStreamSynchronize(c->stream));
    int B, M;
    CKRC(choose_grid(c, &B, &M));
    *c->host_B = 0;                                           // let the device choose n_boxes
    CK(cudaMemcpyAsync(c->host_B_dev, c->host_B, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    volatile unsigned long long *host_iter = reinterpret_cast<volatile unsigned long long *>(c->host_bounds + 4);
    const unsigned long long before = *host_iter;
    int ran = 0;
    // Sharded contexts replay graphs too (one stream, NCCL collectives captured as graph nodes); the first iteration at a new
    // FFT length runs eagerly so that NCCL has seen every collective shape before it is captured.  Every rank sees the same
    // bounds (derived from the same all-gathered bytes), hence takes the same decisions: collectives stay matched.
    static const bool sharded_no_graph = getenv("FITSNE_SHARDED_NO_GRAPH") && atoi(getenv("FITSNE_SHARDED_NO_GRAPH")) != 0;
    if (c->world == 1 || !sharded_no_graph) {
        int eager = 0;
        if (c->world > 1 && !c->nccl_warm.count(M)) {
            c->timing_this_iter = false;
            const uint64_t l0 = c->stats.kernel_launches;
            CKRC(enqueue_iteration_d(c, M, true));
            CK(cudaStreamSynchronize(c->stream));
            c->stats.kernel_launches = l0;          // counted below with the replayed ones
            c->nccl_warm.insert(M);
            eager = 1;
        }
        fitsne_ctx::GraphEntry *ge;
        CKRC(get_graph(c, M, true, &ge));
        for (int i = eager; i < n; i++) CK(cudaGraphLaunch(ge->exec, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        ran = (int) (*host_iter - before);
        c->stats.graph_launches += n - eager;
        c->stats.kernel_launches += ge->launches * (uint64_t) ran;   // no-op launches are not counted as work
    } else {
        // diagnostics (FITSNE_SHARDED_NO_GRAPH=1): plain stream launches, the host running ahead of the device by the batch
        const uint64_t l0 = c->stats.kernel_launches;
        uint64_t per_iter = 0;
        c->timing_this_iter = false;
        for (int i = 0; i < n; i++) {
            CKRC(enqueue_iteration_d(c, M, true));
            if (i == 0) per_iter = c->stats.kernel_launches - l0;
        }
        CK(cudaStreamSynchronize(c->stream));
        ran = (int) (*host_iter - before);
        c->stats.kernel_launches = l0 + per_iter * (uint64_t) ran;
    }
    c->stats.iterations += ran;
    c->steps_total += ran;
    if (c->world > 1 && ran > 0) c->y_whole = false;      // graph replays do not pass through enqueue_iteration's bookkeeping
    c->have_grad = c->have_grad || ran > 0;
    c->bounds_valid = true;
    TRACE("batch of %d at M=%d: %d ran", n, M, ran);
    if (ran < 0 || ran > n) return fail(c, FITSNE_ESTATE, "executed-iterations counter out of range (%d of %d)", ran, n);
    *done = ran;
    if (ran == 0) {
        // even the first step refused: the published bounds must map to a different M than the one just chosen --
        // impossible unless something is inconsistent; fall back to one host-sized step
        CKRC(run_iteration(c, true));
        *done = 1;
    }
    return 0;
}

// Decide the grid from the (host-visible) bounds and run one iteration, through a cached CUDA graph unless
// disabled.  The one host<->device handshake per iteration is the 8-byte bounds read.
static int run_iteration(fitsne_ctx *c, bool update) {
    CKRC(maybe_reorder(c));
    CKRC(refresh_bounds(c));
    TRACE("iteration: waiting for bounds");
    CK(cudaStreamSynchronize(c->stream));
    int B, M;
    CKRC(choose_grid(c, &B, &M));
    *c->host_B = B;
    CK(cudaMemcpyAsync(c->host_B_dev, c->host_B, sizeof(int), cudaMemcpyHostToDevice, c->stream));

    const bool timers = (c->cfg.flags & FITSNE_FLAG_TIMERS) != 0;
    const bool use_graph = !(c->cfg.flags & FITSNE_FLAG_NO_GRAPH) && !timers && c->world == 1;   // single steps of sharded contexts: plain launches
    c->timing_this_iter = timers;
    if (!use_graph) {
        CKRC(enqueue_iteration_d(c, M, update));
    } else {
        fitsne_ctx::GraphEntry *ge;
        CKRC(get_graph(c, M, update, &ge));
        c->stats.kernel_launches += ge->launches;
        CK(cudaGraphLaunch(ge->exec, c->stream));
        c->stats.graph_launches++;
    }
    if (timers) {
        CK(cudaStreamSynchronize(c->stream));
        int order[] = {FITSNE_PHASE_BOUNDS, FITSNE_PHASE_SORT, FITSNE_PHASE_SPREAD, FITSNE_PHASE_COLLECTIVES,
                       FITSNE_PHASE_KERNEL_SPECTRUM, FITSNE_PHASE_FFT, FITSNE_PHASE_GATHER, FITSNE_PHASE_ALLGATHER,
                       FITSNE_PHASE_ATTRACT_UPDATE, FITSNE_PHASE_CENTER, FITSNE_PHASE_COUNT};
        std::vector<int> seq;
        for (int ph : order) {
            if ((ph == FITSNE_PHASE_COLLECTIVES || ph == FITSNE_PHASE_ALLGATHER) && c->world == 1) continue;
            seq.push_back(ph);
        }
        for (size_t i = 0; i + 1 < seq.size(); i++) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, c->ev[seq[i]], c->ev[seq[i + 1]]) == cudaSuccess) c->stats.phase_ms[seq[i]] += ms;
        }
        for (size_t i = 1; i < c->kt_n; i++) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, c->kt_ev[i - 1], c->kt_ev[i]) == cudaSuccess) {
                auto &a = c->kt_acc[c->kt_name[i]];
                a.first += ms; a.second += 1;
            }
        }
        c->kt_n = 0;
        c->timing_this_iter = false;
    }
    c->have_grad = true;
    if (update) { c->stats.iterations++; c->steps_total++; c->bounds_valid = true; }
    return 0;
}

// ------------------------------------------------------------------------------------------ transfers --
// Host arrays are always in the caller's ORIGINAL point order; the device may have re-ordered the points.
static int upload_as_float(fitsne_ctx *c, const double *host, float *dev, size_t n, bool per_point = true) {
    if (n > c->staging_elems) { CKRC(dev_alloc(c, &c->staging, n)); c->staging_elems = n; }
    CK(cudaMemcpyAsync(c->staging, host, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (per_point && c->reordered) {
        if (c->D == 2) k_d2f_ordered<2><<<cdiv(c->N, 256), 256, 0, c->stream>>>(c->staging, dev, c->orig_of, c->N);
        else k_d2f_ordered<1><<<cdiv(c->N, 256), 256, 0, c->stream>>>(c->staging, dev, c->orig_of, c->N);
    } else {
        k_d2f<<<cdiv(n, 256), 256, 0, c->stream>>>(c->staging, dev, n);
    }
    LAUNCH_CHECK();
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}
static int download_as_double(fitsne_ctx *c, const float *dev, double *host, size_t n, bool per_point = true) {
    if (n > c->staging_elems) { CKRC(dev_alloc(c, &c->staging, n)); c->staging_elems = n; }
    if (per_point && c->reordered) {
        if (c->D == 2) k_f2d_ordered<2><<<cdiv(c->N, 256), 256, 0, c->stream>>>(dev, c->staging, c->orig_of, c->N);
        else k_f2d_ordered<1><<<cdiv(c->N, 256), 256, 0, c->stream>>>(dev, c->staging, c->orig_of, c->N);
    } else {
        k_f2d<<<cdiv(n, 256), 256, 0, c->stream>>>(dev, c->staging, n);
    }
    LAUNCH_CHECK();
    CK(cudaMemcpyAsync(host, c->staging, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

static int read_scalars(fitsne_ctx *c) {
    CK(cudaMemcpyAsync(c->host_sc, c->sc, sizeof(Scalars), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

// ------------------------------------------------------------------------------------------- lifetime --
// Sharded contexts: map every peer's exchange buffers (CUDA IPC; same node, NVLink / NVSwitch) so that the iteration's three
// exchanges are plain loads / stores / DMA on peer memory (PeerComm, fitsne_kernels.cuh).  The handles travel through
// one NCCL all-gather at creation; NCCL stays in use for the rare scalar reductions (KL, automatic exaggeration) and
// for completing Y outside the loop.  Any failure leaves the context on the NCCL collectives (FITSNE_NO_P2P=1 forces that).
struct IpcHandles { cudaIpcMemHandle_t Y, grid, stats, flags, S, pot, zs; };
static int setup_peer_fabric(fitsne_ctx *c) {
    static const bool no_p2p = getenv("FITSNE_NO_P2P") && atoi(getenv("FITSNE_NO_P2P")) != 0;
    const int world = c->world;
    if (no_p2p || world > MAX_RANKS) return 0;
    // fixed-size grids for the lifetime of the context: the peers hold mappings of them
    CKRC(ensure_grid_capacity(c, max_fft_len(c->D)));
    CKRC(dev_alloc(c, &c->peer_flags, (size_t) FLAG_KINDS * world));
    CKRC(dev_alloc(c, &c->peer_zs, (size_t) MAX_RANKS));
    CK(cudaMemsetAsync(c->peer_zs, 0, sizeof(double) * MAX_RANKS, c->stream));
    CKRC(dev_alloc(c, &c->comm_seq, (size_t) 1));
    CK(cudaMemsetAsync(c->peer_flags, 0, sizeof(uint32_t) * FLAG_KINDS * world, c->stream));
    CK(cudaMemsetAsync(c->comm_seq, 0, sizeof(unsigned int), c->stream));
    if (c->D == 1) {
        CKRC(dev_alloc(c, &c->grid1d, (size_t) 2 * max_fft_len(1)));
        CK(cudaMemsetAsync(c->grid1d, 0, sizeof(float2) * 2 * max_fft_len(1), c->stream));
    }
    IpcHandles mine;
    void *grid = c->D == 2 ? (void *) c->chg : (void *) c->grid1d;
    memset(&mine, 0, sizeof mine);
    int okl = cudaIpcGetMemHandle(&mine.Y, c->Y) == cudaSuccess && cudaIpcGetMemHandle(&mine.grid, grid) == cudaSuccess &&
              cudaIpcGetMemHandle(&mine.stats, c->shard_stats) == cudaSuccess && cudaIpcGetMemHandle(&mine.flags, c->peer_flags) == cudaSuccess &&
              cudaIpcGetMemHandle(&mine.zs, c->peer_zs) == cudaSuccess;
    if (c->D == 2) okl = okl && cudaIpcGetMemHandle(&mine.S, c->S) == cudaSuccess && cudaIpcGetMemHandle(&mine.pot, c->pot) == cudaSuccess;
    cudaGetLastError();
    // all-gather the handles (+ a per-rank "ok" word in front) through NCCL
    const size_t rec = sizeof(int) * 4 + sizeof(IpcHandles);
    unsigned char *dbuf = nullptr;
    CKRC(dev_alloc(c, &dbuf, rec * world));
    std::vector<unsigned char> host(rec * world, 0);
    memcpy(host.data() + rec * c->rank, &okl, sizeof(int));
    memcpy(host.data() + rec * c->rank + sizeof(int) * 4, &mine, sizeof mine);
    CK(cudaMemcpyAsync(dbuf + rec * c->rank, host.data() + rec * c->rank, rec, cudaMemcpyHostToDevice, c->stream));
    CKNCCL(g_nccl.AllGather(dbuf + rec * c->rank, dbuf, rec, ncclChar, c->comm, c->stream));
    CK(cudaMemcpyAsync(host.data(), dbuf, rec * world, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    cudaFree(dbuf);
    bool all_ok = true;
    for (int r = 0; r < world; r++) { int o; memcpy(&o, host.data() + rec * r, sizeof(int)); all_ok = all_ok && o; }
    PeerComm pc;
    memset(&pc, 0, sizeof pc);
    pc.rank = c->rank; pc.world = world; pc.seq = c->comm_seq;
    int opened_ok = all_ok ? 1 : 0;
    for (int r = 0; r < world && opened_ok; r++) {
        if (r == c->rank) {
            pc.Y[r] = c->Y; pc.grid[r] = grid; pc.stats[r] = c->shard_stats; pc.flags[r] = c->peer_flags;
            pc.S[r] = c->S; pc.pot[r] = c->pot; pc.zs[r] = c->peer_zs;
            continue;
        }
        IpcHandles h;
        memcpy(&h, host.data() + rec * r + sizeof(int) * 4, sizeof h);
        void *p[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        const cudaIpcMemHandle_t *hs[7] = {&h.Y, &h.grid, &h.stats, &h.flags, &h.zs, &h.S, &h.pot};
        const int nh = c->D == 2 ? 7 : 5;
        for (int k = 0; k < nh; k++) {
            if (cudaIpcOpenMemHandle(&p[k], *hs[k], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); opened_ok = 0; break; }
            c->ipc_opened.push_back(p[k]);
        }
        pc.Y[r] = (float *) p[0]; pc.grid[r] = p[1]; pc.stats[r] = p[2]; pc.flags[r] = (uint32_t *) p[3];
        pc.zs[r] = (double *) p[4]; pc.S[r] = (float2 *) p[5]; pc.pot[r] = (float4 *) p[6];
    }
    // everybody must agree (a rank that failed to map a peer cannot be waited for)
    int *agree = nullptr;
    CKRC(dev_alloc(c, &agree, (size_t) 1));
    CK(cudaMemcpyAsync(agree, &opened_ok, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CKNCCL(g_nccl.AllReduce(agree, agree, 1, ncclInt, ncclMin, c->comm, c->stream));
    CK(cudaMemcpyAsync(&opened_ok, agree, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    cudaFree(agree);
    if (!opened_ok) {
        TRACE("peer fabric unavailable: staying on NCCL collectives");
        for (void *p : c->ipc_opened) cudaIpcCloseMemHandle(p);
        c->ipc_opened.clear();
        return 0;
    }
    c->pc = pc;
    c->p2p = true;
    // distributing the convolution trades 7/8 of its work for three more flag stages: worth it from four ranks up
    // (measured on 2 x B200: 0.147 ms distributed vs 0.125 ms replicated at M = 1152); FITSNE_DIST_CONV=0/1 overrides
    c->dist_conv = c->D == 2 && (getenv("FITSNE_DIST_CONV") ? atoi(getenv("FITSNE_DIST_CONV")) != 0 : world >= 4);
    int prio_lo = 0, prio_hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    CK(cudaStreamCreateWithPriority(&c->stream_c, cudaStreamNonBlocking, prio_hi));
    CK(cudaEventCreateWithFlags(&c->ev_cfork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_cjoin, cudaEventDisableTiming));
    if (!getenv("FITSNE_NO_KSIDE")) {
        CK(cudaStreamCreateWithFlags(&c->stream_k, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&c->ev_kfork, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c->ev_kjoin, cudaEventDisableTiming));
    }
    TRACE("peer fabric up: %d ranks", world);
    return 0;
}

static int create_impl(fitsne_ctx *c, const fitsne_config *cfg, int N, int no_dims, const unsigned int *row_P,
                       const unsigned int *col_P, const double *val_P, const double *Y0, int rank, int world,
                       int row_begin, int row_end, const void *nccl_id) {
    if (!cfg || !row_P || N < 2) return fail(c, FITSNE_EINVAL, "bad arguments");
    if (no_dims != 1 && no_dims != 2)
        return fail(c, FITSNE_EINVAL, "FFT interpolation scheme supports only 1 or 2 output dimensions, not %d", no_dims);
    if (cfg->nterms < 1 || cfg->nterms > PMAX) return fail(c, FITSNE_EINVAL, "nterms must be in 1..%d", PMAX);
    if (!(cfg->df > 0) || !(cfg->intervals_per_integer > 0) || cfg->min_num_intervals < 1)
        return fail(c, FITSNE_EINVAL, "df, intervals_per_integer and min_num_intervals must be positive");
    if (world < 1 || world > MAX_RANKS || rank < 0 || rank >= world)
        return fail(c, FITSNE_EINVAL, "world size must be in 1..%d (the GPUs of one node) and 0 <= rank < world, got rank %d of %d", MAX_RANKS, rank, world);
    c->cfg = *cfg;
    c->N = N; c->D = no_dims;
    c->rank = rank; c->world = world;
    c->per = cdiv(N, world);
    if (row_begin != rank * c->per || row_end != std::min(N, (rank + 1) * c->per) || row_end <= row_begin)
        return fail(c, FITSNE_EINVAL, "rank %d must own rows [%d,%d) (contiguous ceil(N/world) slices)", rank,
                    rank * c->per, std::min(N, (rank + 1) * c->per));
    c->row_begin = row_begin; c->row_end = row_end; c->nloc = row_end - row_begin;
    c->df_is_one = cfg->df == 1.0;
    c->n_fwd = 1 + no_dims + (c->df_is_one ? 1 : 0);
    c->n_kern = no_dims + 2;
    c->n_inv = 1 + no_dims;

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(c, FITSNE_ENODEV, "no CUDA device: libfitsne_b200 has no CPU fallback");
    }
    if (cfg->device >= 0) { CK(cudaSetDevice(cfg->device)); }
    CK(cudaGetDevice(&c->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, c->device));
    if (prop.major < 10) return fail(c, FITSNE_ENODEV, "device %d is sm_%d%d; this library is built for sm_100a (B200)", c->device, prop.major, prop.minor);
    int prio_lo = 0, prio_hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    CK(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_hi));
    c->use_pdl = !(getenv("FITSNE_PDL") && atoi(getenv("FITSNE_PDL")) == 0);
    c->ktimes_on = getenv("FITSNE_KTIMES") && atoi(getenv("FITSNE_KTIMES")) != 0;
    for (auto &e : c->ev) CK(cudaEventCreate(&e));
    CK(cudaFuncSetAttribute(k_radix_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    CK(cudaFuncSetAttribute(k_radix_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    CK(cudaFuncSetAttribute((k_spread_chunks<2, 4>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int) spread_smem_bytes<2, 4>()));
    CK(cudaFuncSetAttribute(k_fft_line, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    CK(cudaFuncSetAttribute(k_conv_rows_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    CK(cudaFuncSetAttribute(k_conv_rows_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    CK(cudaFuncSetAttribute(k_kspec_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    CK(cudaFuncSetAttribute(k_kspec_cols<COL_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    CK(cudaFuncSetAttribute(k_kspec_cols<COL_THREADS_MAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    CK(cudaFuncSetAttribute(k_conv_cols<COL_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    CK(cudaFuncSetAttribute(k_conv_cols<COL_THREADS_MAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));

    const size_t yel = (size_t) c->per * world * no_dims;
    c->y_elems = yel;
    CKRC(dev_alloc(c, &c->Y, yel)); CKRC(dev_alloc(c, &c->Yb, yel));
    CKRC(dev_alloc(c, &c->uY, yel)); CKRC(dev_alloc(c, &c->gains, yel));
    CKRC(dev_alloc(c, &c->frep, yel)); CKRC(dev_alloc(c, &c->dC, yel)); CKRC(dev_alloc(c, &c->attr, yel));
    CK(cudaMemsetAsync(c->Y, 0, yel * 4, c->stream)); CK(cudaMemsetAsync(c->Yb, 0, yel * 4, c->stream));
    CK(cudaMemsetAsync(c->uY, 0, yel * 4, c->stream)); CK(cudaMemsetAsync(c->frep, 0, yel * 4, c->stream));
    CK(cudaMemsetAsync(c->dC, 0, yel * 4, c->stream)); CK(cudaMemsetAsync(c->attr, 0, yel * 4, c->stream));
    k_fill<<<cdiv(yel, 256), 256, 0, c->stream>>>(c->gains, 1.0f, yel);
    LAUNCH_CHECK();

    // CSR: full row offsets, this rank's edge slice
    c->edge_base = row_P[row_begin];
    c->E = (size_t) row_P[row_end] - (size_t) row_P[row_begin];
    CKRC(dev_alloc(c, &c->row_P, (size_t) N + 1));
    CK(cudaMemcpyAsync(c->row_P, row_P, ((size_t) N + 1) * 4, cudaMemcpyHostToDevice, c->stream));
    CKRC(dev_alloc(c, &c->edges, c->E + 1));
    if (c->E) {
        const bool from_files = c->src_col && c->src_val;
        if (!from_files && (!col_P || !val_P)) return fail(c, FITSNE_EINVAL, "col_P/val_P are NULL but the graph has edges");
        // (col u32, val f64) -> one 8-byte edge word (col, fp32 weight), in bounded chunks through two staging buffers.
        // From files: the chunks are read into pinned memory and go straight to the device -- no host copy of the edges.
        const size_t chunk = 1u << 24;
        CKRC(dev_alloc(c, &c->staging, chunk)); c->staging_elems = chunk;
        uint32_t *col_stage = nullptr, *hcol = nullptr;
        double *hval = nullptr;
        CKRC(dev_alloc(c, &col_stage, chunk));
        if (from_files) {
            CK(cudaHostAlloc((void **) &hcol, chunk * 4, cudaHostAllocDefault));
            CK(cudaHostAlloc((void **) &hval, chunk * 8, cudaHostAllocDefault));
            if (fseeko(c->src_col, (off_t) c->edge_base * 4, SEEK_SET) != 0 || fseeko(c->src_val, (off_t) c->edge_base * 8, SEEK_SET) != 0)
                return fail(c, FITSNE_EINVAL, "cannot seek in the affinity files");
        }
        int rc_files = 0;
        for (size_t off = 0; off < c->E && rc_files == 0; off += chunk) {
            const size_t n = std::min(chunk, c->E - off);
            const uint32_t *cs = col_P ? col_P + off : nullptr;
            const double *vs = val_P ? val_P + off : nullptr;
            if (from_files) {
                if (fread(hcol, 4, n, c->src_col) != n || fread(hval, 8, n, c->src_val) != n) { rc_files = 1; break; }
                cs = hcol; vs = hval;
            }
            CK(cudaMemcpyAsync(col_stage, cs, n * 4, cudaMemcpyHostToDevice, c->stream));
            CK(cudaMemcpyAsync(c->staging, vs, n * 8, cudaMemcpyHostToDevice, c->stream));
            k_pack_edges<<<cdiv(n, 256), 256, 0, c->stream>>>(col_stage, c->staging, c->edges + off, n);
            CK(cudaStreamSynchronize(c->stream));
        }
        cudaFree(col_stage);
        if (hcol) cudaFreeHost(hcol);
        if (hval) cudaFreeHost(hval);
        if (rc_files) return fail(c, FITSNE_EINVAL, "affinity files are shorter than P_row.dat says (%zu edges)", c->E);
    }
    const double avg = (double) c->E / (double) std::max(1, c->nloc);
    c->lpr = avg > 96 ? 32 : avg > 40 ? 16 : avg > 6 ? 8 : 4;   // lanes per CSR row (B200 sweep at 30 nnz/row: 8 lanes best)

    CKRC(dev_alloc(c, &c->keys[0], (size_t) c->nloc)); CKRC(dev_alloc(c, &c->keys[1], (size_t) c->nloc));
    CKRC(dev_alloc(c, &c->perm[0], (size_t) c->nloc)); CKRC(dev_alloc(c, &c->perm[1], (size_t) c->nloc));
    CKRC(dev_alloc(c, &c->sorted_u, (size_t) c->nloc * no_dims));
    c->hist_cap = (size_t) cdiv(c->nloc, SORT_TILE) * (1 << SORT_MAX_BITS);
    CKRC(dev_alloc(c, &c->hist, c->hist_cap)); CKRC(dev_alloc(c, &c->sweep_state, 2 * c->hist_cap));
    CKRC(dev_alloc(c, &c->sort_bases, (size_t) 2 * (1 << SORT_MAX_BITS)));
    c->chunk = CHUNK;
    if (getenv("FITSNE_CHUNK") && atoi(getenv("FITSNE_CHUNK")) > 0) c->chunk = std::min(64, atoi(getenv("FITSNE_CHUNK")));
    const int sp_points = SP2_THREADS * c->chunk;            // sorted points per spread CTA
    CKRC(dev_alloc(c, &c->work, (size_t) cdiv(c->nloc, c->chunk) + 2));
    CKRC(dev_alloc(c, &c->sort_totals, (size_t) 2 * (1 << SORT_MAX_BITS)));
    {
        const size_t nodes = no_dims == 2 ? (size_t) cfg->nterms * cfg->nterms : (size_t) cfg->nterms;
        CKRC(dev_alloc(c, &c->slots, (size_t) cdiv(c->nloc, sp_points) * 2 * nodes));
        const bool generic = cfg->nterms < 2 || cfg->nterms > (no_dims == 2 ? 4 : 5);      // launch_spread_gather's fallback
        if (generic) CKRC(dev_alloc(c, &c->gpart, (size_t) cdiv(c->nloc, sp_points) * SP2_THREADS * 2 * nodes));
    }
    CKRC(dev_alloc(c, &c->colsum_partial, (size_t) RED_BLOCKS * 2));
    CKRC(dev_alloc(c, &c->bounds_partial, (size_t) RED_BLOCKS));
    CKRC(dev_alloc(c, &c->zpartial, (size_t) 4096));     // per-column (2-D, <= M/2+1) or per-block (1-D) Parseval partials
    CKRC(dev_alloc(c, &c->kl_partial, (size_t) 4096));
    CKRC(dev_alloc(c, &c->gp, (size_t) 1)); CKRC(dev_alloc(c, &c->sp, (size_t) 1)); CKRC(dev_alloc(c, &c->sc, (size_t) 1));
    CKRC(dev_alloc(c, &c->mismatch, (size_t) 1));
    CKRC(dev_alloc(c, &c->tickets, (size_t) 16));
    if (world > 1) {
        CKRC(dev_alloc(c, &c->shard_stats, (size_t) world));
        CKRC(dev_alloc(c, &c->shard_sum_partial, (size_t) RED_BLOCKS * 2));
        CKRC(dev_alloc(c, &c->shard_mm_partial, (size_t) RED_BLOCKS));
        CK(cudaMemsetAsync(c->shard_stats, 0, sizeof(ShardStats) * world, c->stream));
    }
    CK(cudaMemsetAsync(c->tickets, 0, 16 * sizeof(unsigned int), c->stream));
    CK(cudaMemsetAsync(c->gp, 0, sizeof(GridParams), c->stream));
    CK(cudaMemsetAsync(c->sc, 0, sizeof(Scalars), c->stream));
    CK(cudaMemsetAsync(c->mismatch, 0, sizeof(int), c->stream));
    CK(cudaHostAlloc((void **) &c->host_bounds, 64, cudaHostAllocMapped));
    // pinned allocations are recycled by the driver WITHOUT being cleared: a later context of the same process would
    // otherwise start from the previous context's bounds / executed-iterations word (run_batch reads the latter before
    // its first launch) -- found as a wrong iteration count + out-of-range costs[] write in the third context of a process
    memset(c->host_bounds, 0, 64);
    CK(cudaHostGetDevicePointer((void **) &c->host_bounds_dev, c->host_bounds, 0));
    c->host_B = reinterpret_cast<int *>(c->host_bounds + 8);
    CKRC(dev_alloc(c, &c->host_B_dev, (size_t) 1));      // read by k_setup_grid: device memory, not a PCIe round trip per iteration
    CK(cudaMemsetAsync(c->host_B_dev, 0, sizeof(int), c->stream));
    CK(cudaHostAlloc((void **) &c->host_sc, sizeof(Scalars), cudaHostAllocDefault));
    memset(c->host_sc, 0, sizeof(Scalars));
    if (Y0) CKRC(upload_as_float(c, Y0, c->Y, (size_t) N * no_dims));

    if (world > 1) {
        if (!nccl_id) return fail(c, FITSNE_EINVAL, "sharded context needs an ncclUniqueId");
        if (!g_nccl.load(c->err)) return FITSNE_ENCCL;
        ncclUniqueId id;
        memcpy(&id, nccl_id, sizeof id);
        CKNCCL(g_nccl.CommInitRank(&c->comm, world, id, rank));
        CKRC(setup_peer_fabric(c));
    }
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" {

const char *fitsne_version(void) { return "fitsne_b200 0.1 (sm_100a; protocol-compatible with FIt-SNE 1.2.1)"; }

int fitsne_nccl_unique_id(void *out) {
    std::string err;
    if (!g_nccl.load(err)) { g_create_error = err; return FITSNE_ENCCL; }
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) { g_create_error = "ncclGetUniqueId failed"; return FITSNE_ENCCL; }
    memcpy(out, &id, sizeof id);
    return 0;
}

int fitsne_destroy(fitsne_ctx *c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    drop_graphs(c);
    for (auto &p : c->plans) if (p.second.W) cudaFree(p.second.W);
    if (c->p2p) {
        // nobody frees a buffer a peer may still have mapped: close my mappings, then meet the others
        if (c->stream_c) cudaStreamSynchronize(c->stream_c);
        if (c->stream_k) cudaStreamSynchronize(c->stream_k);
        for (void *p : c->ipc_opened) cudaIpcCloseMemHandle(p);
        if (c->comm && c->comm_seq) {
            g_nccl.AllReduce(c->comm_seq, c->comm_seq, 1, ncclInt, ncclMax, c->comm, c->stream);
            cudaStreamSynchronize(c->stream);
        }
        if (c->ev_cfork) cudaEventDestroy(c->ev_cfork);
        if (c->ev_cjoin) cudaEventDestroy(c->ev_cjoin);
        if (c->stream_c) cudaStreamDestroy(c->stream_c);
        if (c->ev_kfork) cudaEventDestroy(c->ev_kfork);
        if (c->ev_kjoin) cudaEventDestroy(c->ev_kjoin);
        if (c->stream_k) cudaStreamDestroy(c->stream_k);
    }
    if (c->comm) g_nccl.CommDestroy(c->comm);
    void *bufs[] = {c->Y, c->Yb, c->uY, c->gains, c->frep, c->dC, c->row_P, c->edges, c->keys[0], c->keys[1],
                    c->perm[0], c->perm[1], c->sorted_u, c->box_range, c->gpart, c->hist, c->sweep_state, c->sort_bases, c->work, c->sort_totals, c->slots, c->attr, c->planes,
                    c->chg, c->pot, c->S, c->KR, c->KS, c->colsum_partial, c->zpartial, c->kl_partial, c->bounds_partial,
                    c->gp, c->sp, c->sc, c->mismatch, c->tickets, c->host_B_dev, c->staging, c->orig_of, c->orig_tmp, c->pos_of, c->rank_map,
                    c->row_P2, c->edges2, c->tile_cnt, c->tile_start, c->tile_cur, c->tile_pack, c->tile_val,
                    c->nonempty, c->gp_reorder, c->peer_flags, c->peer_zs, c->comm_seq, c->grid1d, c->shard_stats, c->shard_sum_partial, c->shard_mm_partial};
    for (void *b : bufs) if (b) cudaFree(b);
    if (c->host_bounds) cudaFreeHost(c->host_bounds);
    if (c->host_sc) cudaFreeHost(c->host_sc);
    for (auto &e : c->ev) if (e) cudaEventDestroy(e);
    for (auto &e : c->kt_ev) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

int fitsne_create_sharded(const fitsne_config *cfg, int N, int no_dims, const unsigned int *row_P,
                          const unsigned int *col_P_local, const double *val_P_local, const double *Y0, int rank,
                          int world_size, int row_begin, int row_end, const void *nccl_unique_id, fitsne_ctx **out) {
    if (!out) return FITSNE_EINVAL;
    *out = nullptr;
    fitsne_ctx *c = new fitsne_ctx();
    int rc = create_impl(c, cfg, N, no_dims, row_P, col_P_local, val_P_local, Y0, rank, world_size, row_begin, row_end,
                         nccl_unique_id);
    if (rc != 0) {
        g_create_error = c->err;
        fitsne_destroy(c);
        return rc;
    }
    *out = c;
    return 0;
}

int fitsne_create(const fitsne_config *cfg, int N, int no_dims, const unsigned int *row_P, const unsigned int *col_P,
                  const double *val_P, const double *Y0, fitsne_ctx **out) {
    return fitsne_create_sharded(cfg, N, no_dims, row_P, col_P, val_P, Y0, 0, 1, 0, N, nullptr, out);
}

// P as files: the reference's load_affinities side files (tsne.cpp:236-281, :334-366) as a first-class input
static std::string affinity_path(const char *dir, const char *name) {
    const char *d = dir && *dir ? dir : getenv("FITSNE_AFFINITIES_DIR");
    return (d && *d ? std::string(d) + "/" : std::string()) + name;
}
int fitsne_create_from_files_sharded(const fitsne_config *cfg, const char *dir, int N, int no_dims, const double *Y0, int rank,
                                     int world_size, const void *nccl_unique_id, fitsne_ctx **out) {
    if (!out || N < 2 || world_size < 1) return FITSNE_EINVAL;
    *out = nullptr;
    FILE *fr = fopen(affinity_path(dir, "P_row.dat").c_str(), "rb");
    FILE *fc = fopen(affinity_path(dir, "P_col.dat").c_str(), "rb");
    FILE *fv = fopen(affinity_path(dir, "P_val.dat").c_str(), "rb");
    std::vector<unsigned int> row((size_t) N + 1);
    int rc = 0;
    if (!fr || !fc || !fv) { g_create_error = "cannot open P_row.dat / P_col.dat / P_val.dat in " + affinity_path(dir, ""); rc = FITSNE_EINVAL; }
    else if (fread(row.data(), sizeof(unsigned int), (size_t) N + 1, fr) != (size_t) N + 1) { g_create_error = "P_row.dat is shorter than N + 1 entries"; rc = FITSNE_EINVAL; }
    if (rc == 0) {
        fitsne_ctx *c = new fitsne_ctx();
        c->src_col = fc; c->src_val = fv;
        const int per = cdiv(N, world_size);
        rc = create_impl(c, cfg, N, no_dims, row.data(), nullptr, nullptr, Y0, rank, world_size, rank * per, std::min(N, (rank + 1) * per),
                         nccl_unique_id);
        c->src_col = c->src_val = nullptr;
        if (rc != 0) { g_create_error = c->err; fitsne_destroy(c); } else *out = c;
    }
    if (fr) fclose(fr);
    if (fc) fclose(fc);
    if (fv) fclose(fv);
    return rc;
}
int fitsne_create_from_files(const fitsne_config *cfg, const char *dir, int N, int no_dims, const double *Y0, fitsne_ctx **out) {
    return fitsne_create_from_files_sharded(cfg, dir, N, no_dims, Y0, 0, 1, nullptr, out);
}

const char *fitsne_last_error(const fitsne_ctx *c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int fitsne_synchronize(fitsne_ctx *c) {
    if (!c) return FITSNE_EINVAL;
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

int fitsne_set_Y(fitsne_ctx *c, const double *Y) {
    if (!c || !Y) return FITSNE_EINVAL;
    CK(cudaSetDevice(c->device));
    c->bounds_valid = false;
    c->y_whole = true;
    return upload_as_float(c, Y, c->Y, (size_t) c->N * c->D);
}

int fitsne_get_Y(fitsne_ctx *c, double *Y) {
    if (!c || !Y) return FITSNE_EINVAL;
    CK(cudaSetDevice(c->device));
    CKRC(ensure_whole_Y(c));
    return download_as_double(c, c->Y, Y, (size_t) c->N * c->D);
}

int fitsne_set_optimizer_state(fitsne_ctx *c, const double *uY, const double *gains) {
    if (!c) return FITSNE_EINVAL;
    CK(cudaSetDevice(c->device));
    if (uY) CKRC(upload_as_float(c, uY, c->uY, (size_t) c->N * c->D));
    if (gains) CKRC(upload_as_float(c, gains, c->gains, (size_t) c->N * c->D));
    return 0;
}

int fitsne_get_optimizer_state(fitsne_ctx *c, double *uY, double *gains) {
    if (!c) return FITSNE_EINVAL;
    CK(cudaSetDevice(c->device));
    // in a sharded run each rank only maintains its own rows of uY / gains
    if (uY) CKRC(download_as_double(c, c->uY, uY, (size_t) c->N * c->D));
    if (gains) CKRC(download_as_double(c, c->gains, gains, (size_t) c->N * c->D));
    return 0;
}

static StepParams make_sp(const fitsne_ctx *c, double alpha, double momentum, double lr, double msn, int mode) {
    StepParams sp;
    sp.alpha = (float) alpha; sp.momentum = (float) momentum; sp.lr = (float) lr; sp.max_step_norm = (float) msn;
    sp.mode = mode; sp.inv_df = (float) (1.0 / c->cfg.df);
    return sp;
}

int fitsne_gradient(fitsne_ctx *c, double exaggeration, double *dC_out, double *sum_Q_out) {
    if (!c) return FITSNE_EINVAL;
    CK(cudaSetDevice(c->device));
    CKRC(push_step_params(c, make_sp(c, exaggeration, 0, 0, 0, 0)));
    CKRC(run_iteration(c, false));
    if (dC_out) {
        // each rank computed its own rows; rows outside the shard read back as zero
        CKRC(download_as_double(c, c->dC, dC_out, (size_t) c->N * c->D));
    }
    if (sum_Q_out) {
        CKRC(read_scalars(c));
        *sum_Q_out = c->host_sc->Z;
    }
    return 0;
}

int fitsne_step(fitsne_ctx *c, const fitsne_step_params *p) {
    if (!c || !p) return FITSNE_EINVAL;
    if (p->mode < 0 || p->mode > 2) return fail(c, FITSNE_EINVAL, "bad step mode %d", p->mode);
    CK(cudaSetDevice(c->device));
    CKRC(push_step_params(c, make_sp(c, p->exaggeration, p->momentum, p->learning_rate, p->max_step_norm, p->mode)));
    return run_iteration(c, true);
}

static int kl_impl(fitsne_ctx *c, double exaggeration, double *C_out) {
    if (!c->have_grad) return fail(c, FITSNE_ESTATE, "fitsne_kl needs the sum_Q of a previous gradient/step");
    CKRC(ensure_whole_Y(c));
    const int blocks = std::min(4096, std::max(1, cdiv((long long) (c->row_end - c->row_begin) * 32, 256)));
    if (c->D == 2)
        k_kl<2><<<blocks, 256, 0, c->stream>>>(c->row_P, c->edges, c->edge_base, c->Y, c->row_begin, c->row_end,
                                              exaggeration, c->cfg.df, c->sc, c->kl_partial);
    else
        k_kl<1><<<blocks, 256, 0, c->stream>>>(c->row_P, c->edges, c->edge_base, c->Y, c->row_begin, c->row_end,
                                              exaggeration, c->cfg.df, c->sc, c->kl_partial);
    k_finalize_kl<<<1, 256, 0, c->stream>>>(c->kl_partial, blocks, c->sc);
    LAUNCH_CHECK();
    c->stats.kernel_launches += 2;
    if (c->world > 1) {
        double *klp = &c->sc->kl;
        CKNCCL(g_nccl.AllReduce(klp, klp, 1, ncclDouble, ncclSum, c->comm, c->stream));
    }
    CKRC(read_scalars(c));
    *C_out = c->host_sc->kl;
    return 0;
}

int fitsne_kl(fitsne_ctx *c, double exaggeration, double *C_out) {
    if (!c || !C_out) return FITSNE_EINVAL;
    CK(cudaSetDevice(c->device));
    return kl_impl(c, exaggeration, C_out);
}

static int auto_exaggeration(fitsne_ctx *c, double learning_rate, double *coeff) {
    const int blocks = 1024;
    k_row_sum_max<<<blocks, 256, 0, c->stream>>>(c->row_P, c->edges, c->edge_base, c->row_begin, c->row_end, c->kl_partial);
    LAUNCH_CHECK();
    std::vector<double> h(blocks);
    CK(cudaMemcpyAsync(h.data(), c->kl_partial, blocks * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    double mx = 0;
    for (double v : h) mx = std::max(mx, v);
    if (c->world > 1) {
        double *d = &c->sc->kl;
        CK(cudaMemcpyAsync(d, &mx, sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CKNCCL(g_nccl.AllReduce(d, d, 1, ncclDouble, ncclMax, c->comm, c->stream));
        CK(cudaMemcpyAsync(&mx, d, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    *coeff = 1.0 / (learning_rate * mx);
    return 0;
}

int fitsne_run(fitsne_ctx *c, const fitsne_schedule *s, double *costs, double *Y_out) {
    if (!c || !s) return FITSNE_EINVAL;
    CK(cudaSetDevice(c->device));
    double early = s->early_exag_coeff;
    if (early == 0) {   // tsne.cpp:392-402
        CKRC(auto_exaggeration(c, s->learning_rate, &early));
        if (s->verbose) printf("Max of the val_Ps is: %lf\n", 1.0 / (early * s->learning_rate));
    }
    if (s->verbose) printf("Exaggerating Ps by %f\n", early);
    double alpha = early, momentum = s->momentum;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, c->stream));
    auto t0 = std::chrono::steady_clock::now();
    // Sharded runs batch too, but with plain launches instead of graphs (run_batch); FITSNE_SHARDED_SYNC=1 restores one
    // host round trip per iteration there (diagnostics).
    const bool batched = c->world == 1 ? !(c->cfg.flags & (FITSNE_FLAG_NO_GRAPH | FITSNE_FLAG_TIMERS | FITSNE_FLAG_NO_SPECULATION))
                                       : !(c->cfg.flags & (FITSNE_FLAG_TIMERS | FITSNE_FLAG_NO_SPECULATION));
    int iter = 0;
    while (iter < s->max_iter) {
        int mode = FITSNE_STEP_MOMENTUM_CLIP;
        if (s->no_momentum_during_exag) mode = iter > s->stop_lying_iter ? FITSNE_STEP_MOMENTUM : FITSNE_STEP_PLAIN_GD;
        CKRC(push_step_params(c, make_sp(c, alpha, momentum, s->learning_rate, s->max_step_norm, mode)));
        // the batch [iter, end] shares one set of step parameters: it stops at the next schedule event, KL evaluation,
        // point re-ordering or after 64 steps, whichever comes first
        int end = std::min(s->max_iter - 1, iter + 63);
        auto clip = [&](long long ev) { if (ev >= iter && ev < end) end = (int) ev; };
        clip(s->stop_lying_iter); clip(s->start_late_exag_iter); clip(s->mom_switch_iter);
        clip((long long) (iter / 50 + 1) * 50 - 1);
        if (!(c->cfg.flags & FITSNE_FLAG_NO_REORDER) && c->reordered)
            clip((long long) iter + (long long) (c->last_reorder_iter + c->reorder_interval - c->steps_total) - 1);
        int ran = 1;
        if (batched) CKRC(run_batch(c, end - iter + 1, &ran));
        else { CKRC(run_iteration(c, true)); end = iter; }
        iter += ran;
        if (iter <= end) continue;          // a step refused its grid: re-plan from the fresh bounds
        const int last = end;
        // schedule changes take effect after the step of that iteration (tsne.cpp:534-544).  The reference
        // un-exaggerates by dividing and late-exaggerates by multiplying the stored P.
        if (last == s->stop_lying_iter) {
            if (s->verbose) printf("Unexaggerating Ps by %f\n", early);
            alpha /= early;
        }
        if (last == s->start_late_exag_iter) {
            if (s->verbose) printf("Exaggerating Ps by %f\n", s->late_exag_coeff);
            alpha *= s->late_exag_coeff;
        }
        if (last == s->mom_switch_iter) momentum = s->final_momentum;
        if ((last + 1) % 50 == 0 || last == s->max_iter - 1) {
            double C = 0;
            CKRC(kl_impl(c, alpha, &C));
            if (last < s->stop_lying_iter && s->stop_lying_iter != -1) C = C / early - log(early);
            if (last >= s->start_late_exag_iter && s->start_late_exag_iter != -1) C = C / s->late_exag_coeff - log(s->late_exag_coeff);
            if (costs) costs[last] = C;
            if (s->verbose) {
                auto t1 = std::chrono::steady_clock::now();
                printf("Iteration %d (50 iterations in %.2f seconds), cost %f\n", last + 1,
                       std::chrono::duration_cast<std::chrono::milliseconds>(t1 - t0).count() / (float) 1000.0, C);
                t0 = std::chrono::steady_clock::now();
            }
        }
    }
    CK(cudaEventRecord(e1, c->stream));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    c->last_run_ms = ms;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    CKRC(ensure_whole_Y(c));
    if (Y_out) CKRC(download_as_double(c, c->Y, Y_out, (size_t) c->N * c->D));
    return 0;
}

int fitsne_run_host(const fitsne_config *cfg, const fitsne_schedule *s, int N, int no_dims, const unsigned int *row_P,
                    const unsigned int *col_P, const double *val_P, double *Y, double *costs) {
    fitsne_ctx *c = nullptr;
    TRACE("run_host: create");
    int rc = fitsne_create(cfg, N, no_dims, row_P, col_P, val_P, Y, &c);
    if (rc != 0) return rc;
    TRACE("run_host: run");
    rc = fitsne_run(c, s, costs, Y);
    TRACE("run_host: destroy");
    if (rc != 0) g_create_error = c->err;
    fitsne_destroy(c);
    return rc;
}

int fitsne_run_files(const fitsne_config *cfg, const fitsne_schedule *s, const char *dir, int N, int no_dims, double *Y, double *costs) {
    fitsne_ctx *c = nullptr;
    int rc = fitsne_create_from_files(cfg, dir, N, no_dims, Y, &c);
    if (rc != 0) return rc;
    rc = fitsne_run(c, s, costs, Y);
    if (rc != 0) g_create_error = c->err;
    fitsne_destroy(c);
    return rc;
}

int fitsne_prewarm(fitsne_ctx *c, int n_boxes_lo, int n_boxes_hi) {
    if (!c || n_boxes_lo < 1 || n_boxes_hi < n_boxes_lo) return FITSNE_EINVAL;
    CK(cudaSetDevice(c->device));
    int last = -1;
    for (int B = n_boxes_lo; B <= n_boxes_hi; B++) {
        const int M = nice_fft_size(2 * B * c->cfg.nterms);
        if (M == last) continue;
        last = M;
        TRACE("prewarm: plans for M=%d", M);
        CKRC(ensure_grid_capacity(c, M));
        Plans *pl;
        CKRC(get_plans(c, M, &pl));
    }
    return 0;
}

int fitsne_last_run_ms(fitsne_ctx *c, double *ms) {
    if (!c || !ms) return FITSNE_EINVAL;
    *ms = c->last_run_ms;
    return 0;
}

int fitsne_get_stats(fitsne_ctx *c, fitsne_stats *out) {
    if (!c || !out) return FITSNE_EINVAL;
    CK(cudaStreamSynchronize(c->stream));
    c->stats.reorders = c->reorders;
    *out = c->stats;
    return 0;
}

int fitsne_reset_stats(fitsne_ctx *c) {
    if (!c) return FITSNE_EINVAL;
    const fitsne_stats old = c->stats;
    c->stats = fitsne_stats{};
    c->stats.n_boxes = old.n_boxes; c->stats.grid_side = old.grid_side; c->stats.fft_side = old.fft_side;
    c->stats.min_coord = old.min_coord; c->stats.max_coord = old.max_coord;
    return 0;
}

int fitsne_debug_copy(fitsne_ctx *c, const char *what, void *dst, size_t dst_bytes, size_t *needed) {
    if (!c || !what) return FITSNE_EINVAL;
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    const void *src = nullptr;
    size_t bytes = 0;
    const int sorted_buf = 0;   // two LSD passes: sorted data ends up back in buffer 0
    if (!strcmp(what, "frep")) {
        src = c->frep; bytes = (size_t) c->N * c->D * 4;
        if (c->reordered) {   // hand it back in the caller's point order
            if (c->D == 2) k_f2f_ordered<2><<<cdiv(c->N, 256), 256, 0, c->stream>>>(c->frep, c->dC, c->orig_of, c->N);
            else k_f2f_ordered<1><<<cdiv(c->N, 256), 256, 0, c->stream>>>(c->frep, c->dC, c->orig_of, c->N);
            CK(cudaStreamSynchronize(c->stream));
            src = c->dC;
        }
    }
    else if (!strcmp(what, "ktimes")) {      // text: "<kernel> <total ms> <count>" per line (FITSNE_KTIMES=1 + timers mode)
        c->kt_text.clear();
        for (auto &kv : c->kt_acc) {
            char line[160];
            snprintf(line, sizeof line, "%s\t%.6f\t%llu\n", kv.first.c_str(), kv.second.first, (unsigned long long) kv.second.second);
            c->kt_text += line;
        }
        if (needed) *needed = c->kt_text.size();
        if (!dst) return 0;
        if (dst_bytes < c->kt_text.size()) return fail(c, FITSNE_EINVAL, "buffer too small for 'ktimes'");
        memcpy(dst, c->kt_text.data(), c->kt_text.size());
        return 0;
    }
    else if (!strcmp(what, "perm")) { src = c->perm[sorted_buf]; bytes = (size_t) c->nloc * 4; }
    else if (!strcmp(what, "keys")) { src = c->keys[sorted_buf]; bytes = (size_t) c->nloc * 4; }
    else if (!strcmp(what, "box_range")) {      // (first, end) per box; entries of empty boxes are stale
        src = c->box_range; bytes = (c->D == 2 ? (size_t) c->cur_B * c->cur_B : (size_t) c->cur_B) * sizeof(uint2);
    }
    else if (!strcmp(what, "grid") && c->D == 2) { src = c->chg; bytes = (size_t) c->stats.grid_side * c->stats.grid_side * sizeof(float4); }
    else if (!strcmp(what, "pot") && c->D == 2) { src = c->pot; bytes = (size_t) c->stats.grid_side * c->stats.grid_side * sizeof(float4); }
    else return fail(c, FITSNE_EINVAL, "unknown debug array '%s'", what);
    if (needed) *needed = bytes;
    if (!dst) return 0;
    if (dst_bytes < bytes) return fail(c, FITSNE_EINVAL, "buffer too small for '%s': need %zu", what, bytes);
    CK(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return 0;
}

}  // extern "C"
