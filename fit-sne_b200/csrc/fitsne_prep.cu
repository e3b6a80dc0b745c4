// fitsne_prep.cu -- the step immediately before the loop, on the device (SURVEY.md section 8 f1 + f4):
//   fitsne_knn            exact Euclidean kNN by tiled brute force (replaces the Annoy / VP-tree searches,
//                         /root/reference/src/tsne.cpp:1535-1639, :1643-1726): fp32 distance tiles with a per-query
//                         threshold filter and sorted top-(K+8) lists in shared memory, then an fp64 re-ranking of the
//                         survivors, so the K neighbours and their distances are exact
//   fitsne_similarities   per-point bandwidth by bisection on the entropy (computeGaussianPerplexity, tsne.cpp:1394-1500,
//                         incl. the perplexity-list average and the fixed-sigma branch) and symmetrisation + normalisation
//                         (symmetrizeMatrix, tsne.cpp:1730-1828) -> CSR with ascending columns
// Same arithmetic as the host shell's TSNE::input_similarities (host/tsne_host.cpp), which the tests compare with.
#include "../../include/fitsne_b200.h"

#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

std::string g_prep_error;
int prep_fail(int code, const char *what, cudaError_t e) {
    char buf[256];
    snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
    g_prep_error = buf;
    return code;
}
#define PCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return prep_fail(e_ == cudaErrorMemoryAllocation ? FITSNE_ENOMEM : FITSNE_ECUDA, #call, e_); } while (0)

struct DevBuf {        // frees on scope exit
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    template <typename T> T *as() { return reinterpret_cast<T *>(p); }
};

// ------------------------------------------------------------------------------------------------ kNN --
constexpr int KQ = 64;          // queries per CTA
constexpr int KC = 64;          // candidates per step
constexpr int KD = 16;          // feature chunk staged in shared memory
constexpr int KNN_THREADS = 256;
constexpr int KQP = KQ + 4, KCP = KC + 4;   // padded slab pitch: staging stores 2-way instead of 16-way conflicted, float4 reads stay aligned
constexpr int KNN_MARGIN = 8;   // fp32 selection keeps K + margin candidates; the fp64 pass re-ranks them

__global__ void k_to_float_norms(const double *__restrict__ X, int N, int D, float *__restrict__ Xf, float *__restrict__ sq) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float s = 0.f;
    for (int d = 0; d < D; d++) { const float v = (float) X[(size_t) i * D + d]; Xf[(size_t) i * D + d] = v; s += v * v; }
    sq[i] = s;
}

// One CTA = KQ queries against all N candidates, KC at a time.  Each thread owns a 4 x 4 patch of the KQ x KC distance
// tile.  Per query: a sorted list of the best KP = K + margin (distance, index) pairs in shared memory and its threshold
// (the KP-th best so far).  Candidates below the threshold are appended to a per-query inbox during the tile step and
// merged by one warp per query after the barrier -- after the first few steps almost nothing passes the filter.
__global__ void __launch_bounds__(KNN_THREADS) k_knn_tiles(const float *__restrict__ Xf, const float *__restrict__ sq, int N, int D, int KP,
                                                           uint32_t *__restrict__ cand_idx /*[N][KP]*/) {
    extern __shared__ __align__(16) unsigned char knn_raw[];
    float *qs = reinterpret_cast<float *>(knn_raw);                    // [KD][KQ]
    float *cs = qs + KD * KQP;                                           // [KD][KC]
    float *tau = cs + KD * KCP;                                          // [KQ]
    int *inbox_n = reinterpret_cast<int *>(tau + KQ);                    // [KQ]
    float *inbox_d = reinterpret_cast<float *>(inbox_n + KQ);            // [KQ][KC]
    uint32_t *inbox_i = reinterpret_cast<uint32_t *>(inbox_d + KQ * KC); // [KQ][KC]
    float *best_d = reinterpret_cast<float *>(inbox_i + KQ * KC);        // [KQ][KP]
    uint32_t *best_i = reinterpret_cast<uint32_t *>(best_d + KQ * KP);   // [KQ][KP]
    int *best_n = reinterpret_cast<int *>(best_i + KQ * KP);             // [KQ]
    const int q0 = blockIdx.x * KQ;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;              // patch: queries ty*4.., candidates tx*4..
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < KQ; i += KNN_THREADS) { tau[i] = INFINITY; inbox_n[i] = 0; best_n[i] = 0; }
    __syncthreads();
    for (int c0 = 0; c0 < N; c0 += KC) {
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) acc[a][b] = 0.f;
        for (int d0 = 0; d0 < D; d0 += KD) {
            // stage a [KD x KQ] slab of the queries and a [KD x KC] slab of the candidates (feature-major: conflict-free reads)
            for (int e = threadIdx.x; e < KD * KQ; e += KNN_THREADS) {
                const int r = e / KD, d = e - r * KD;                    // consecutive threads -> consecutive features of one row
                const int q = q0 + r, c = c0 + r;
                qs[d * KQP + r] = (q < N && d0 + d < D) ? Xf[(size_t) q * D + d0 + d] : 0.f;
                cs[d * KCP + r] = (c < N && d0 + d < D) ? Xf[(size_t) c * D + d0 + d] : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int d = 0; d < KD; d++) {
                const float4 qa = *reinterpret_cast<const float4 *>(qs + d * KQP + ty * 4);
                const float4 cb = *reinterpret_cast<const float4 *>(cs + d * KCP + tx * 4);
                const float qv[4] = {qa.x, qa.y, qa.z, qa.w}, cv[4] = {cb.x, cb.y, cb.z, cb.w};
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int b = 0; b < 4; b++) acc[a][b] += qv[a] * cv[b];
            }
            __syncthreads();
        }
        // filter against the per-query thresholds
#pragma unroll
        for (int a = 0; a < 4; a++) {
            const int ql = ty * 4 + a, q = q0 + ql;
            if (q >= N) continue;
            const float sqq = sq[q], t = tau[ql];
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int c = c0 + tx * 4 + b;
                if (c >= N || c == q) continue;
                const float dist = fmaxf(sqq + sq[c] - 2.f * acc[a][b], 0.f);
                if (dist < t) {
                    const int slot = atomicAdd(&inbox_n[ql], 1);
                    inbox_d[ql * KC + slot] = dist; inbox_i[ql * KC + slot] = (uint32_t) c;
                }
            }
        }
        __syncthreads();
        // merge the inboxes: one warp per query, sorted insertion (ascending distance, then index)
        for (int ql = warp; ql < KQ; ql += KNN_THREADS / 32) {
            const int nin = inbox_n[ql];
            if (nin == 0) continue;
            float *bd = best_d + ql * KP;
            uint32_t *bi = best_i + ql * KP;
            int nb = best_n[ql];
            for (int e = 0; e < nin; e++) {
                const float dnew = inbox_d[ql * KC + e];
                const uint32_t inew = inbox_i[ql * KC + e];
                if (nb == KP && !(dnew < bd[KP - 1] || (dnew == bd[KP - 1] && inew < bi[KP - 1]))) continue;
                // position = number of entries that sort before the new one
                int before = 0;
                for (int j = lane; j < nb; j += 32) before += (bd[j] < dnew || (bd[j] == dnew && bi[j] < inew)) ? 1 : 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
                const int last = nb < KP ? nb : KP - 1;                  // entry `last` is created or overwritten by the shift
                // shift [before, last) up by one, back to front in chunks of 32 (read all, then write all)
                for (int hi = last; hi > before; hi -= 32) {
                    const int j = hi - lane;                             // destination index
                    float dv = 0.f; uint32_t iv = 0;
                    const bool act = j > before;
                    if (act) { dv = bd[j - 1]; iv = bi[j - 1]; }
                    __syncwarp();
                    if (act) { bd[j] = dv; bi[j] = iv; }
                    __syncwarp();
                }
                if (lane == 0) { bd[before] = dnew; bi[before] = inew; }
                if (nb < KP) nb++;
                __syncwarp();
            }
            if (lane == 0) { best_n[ql] = nb; inbox_n[ql] = 0; if (nb == KP) tau[ql] = bd[KP - 1]; }
        }
        __syncthreads();
    }
    for (int e = threadIdx.x; e < KQ * KP; e += KNN_THREADS) {
        const int ql = e / KP, j = e - ql * KP;
        if (q0 + ql < N) cand_idx[(size_t) (q0 + ql) * KP + j] = j < best_n[ql] ? best_i[ql * KP + j] : 0xffffffffu;
    }
}

// exact fp64 distances of the surviving candidates, re-ranked (distance, then index); the first K are the answer
__global__ void __launch_bounds__(128) k_knn_refine(const double *__restrict__ X, int N, int D, int K, int KP,
                                                    const uint32_t *__restrict__ cand_idx, uint32_t *__restrict__ nbr, double *__restrict__ dist) {
    extern __shared__ double rf_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * 4 + warp;
    double *dd = rf_raw + (size_t) warp * KP * 2;
    double *ii = dd + KP;                                  // indices kept as doubles next to the distances (exact up to 2^53)
    if (q >= N) return;
    for (int j = lane; j < KP; j += 32) {
        const uint32_t c = cand_idx[(size_t) q * KP + j];
        double s = INFINITY;
        if (c != 0xffffffffu) {
            s = 0;
            for (int d = 0; d < D; d++) { const double t = X[(size_t) q * D + d] - X[(size_t) c * D + d]; s += t * t; }
        }
        dd[j] = s; ii[j] = (double) c;
    }
    __syncwarp();
    // rank sort: position = number of candidates that sort before
    for (int j = lane; j < KP; j += 32) {
        int r = 0;
        for (int m = 0; m < KP; m++) r += (dd[m] < dd[j] || (dd[m] == dd[j] && ii[m] < ii[j])) ? 1 : 0;
        if (r < K) { nbr[(size_t) q * K + r] = (uint32_t) ii[j]; dist[(size_t) q * K + r] = sqrt(dd[j]); }
    }
}

// ----------------------------------------------------------------------------------- perplexity search --
// One warp per point.  Mirrors host/tsne_host.cpp:calibrate_row (itself tsne.cpp:1394-1469): bisection on beta, tolerance
// 1e-5, at most 200 steps, kernel exp(-beta * d) on the Euclidean distance (the reference's ifSquared quirk), the row is
// normalised with the sum of the last beta TESTED.
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ void calibrate_row_dev(const double *d, int K, double perplexity, double sigma, double *p, int lane) {
    double beta, sum = DBL_MIN;
    if (perplexity > 0) {
        double lo = -DBL_MAX, hi = DBL_MAX;
        const double target = log(perplexity), tol = 1e-5;
        beta = 1.0;
        for (int it = 0; it < 200; it++) {
            double s = 0, h = 0;
            for (int m = lane; m < K; m += 32) { const double pm = exp(-beta * d[m]); p[m] = pm; s += pm; h += beta * (d[m] * pm); }
            sum = DBL_MIN + warp_sum_d(s);
            h = warp_sum_d(h);
            const double diff = h / sum + log(sum) - target;
            if (diff < tol && -diff < tol) break;
            if (diff > 0) { lo = beta; beta = (hi == DBL_MAX || hi == -DBL_MAX) ? beta * 2.0 : (beta + hi) / 2.0; }
            else { hi = beta; beta = (lo == -DBL_MAX || lo == DBL_MAX) ? beta / 2.0 : (beta + lo) / 2.0; }
        }
    } else {
        beta = 1 / (2 * sigma * sigma);
        double s = 0;
        for (int m = lane; m < K; m += 32) { p[m] = exp(-beta * d[m]); s += p[m]; }
        sum = DBL_MIN + warp_sum_d(s);
    }
    __syncwarp();
    for (int m = lane; m < K; m += 32) p[m] /= sum;
    __syncwarp();
}

__global__ void __launch_bounds__(128) k_calibrate(const double *__restrict__ dist, int N, int K, double perplexity, double sigma, int list_len,
                                                   const double *__restrict__ list, double *__restrict__ cond, double *__restrict__ tmp) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 4 + warp;
    if (i >= N) return;
    const double *d = dist + (size_t) i * K;
    double *p = cond + (size_t) i * K;
    if (perplexity != 0) { calibrate_row_dev(d, K, perplexity, sigma, p, lane); return; }
    double *t = tmp + (size_t) i * K;                      // average over the perplexity list (tsne.cpp:1474-1500)
    calibrate_row_dev(d, K, list[0], sigma, p, lane);
    for (int l = 1; l < list_len; l++) {
        calibrate_row_dev(d, K, list[l], sigma, t, lane);
        for (int m = lane; m < K; m += 32) p[m] += t[m];
        __syncwarp();
    }
    for (int m = lane; m < K; m += 32) p[m] = p[m] / list_len;
}

// ------------------------------------------------------------------------------------- symmetrisation --
// P_sym = (P + P^T) / 2, normalised to sum 1, as CSR with ascending columns.  Row i holds its own K neighbours (value
// p_j|i, plus p_i|j when j lists i as well) and every j that lists i without being listed by i.
__global__ void __launch_bounds__(256) k_sym_count(const uint32_t *__restrict__ nbr, int N, int K, uint32_t *__restrict__ extra) {
    const size_t e = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t) N * K) return;
    const uint32_t i = (uint32_t) (e / K), j = nbr[e];
    const uint32_t *nj = nbr + (size_t) j * K;
    bool mutual = false;
    for (int m = 0; m < K; m++) mutual = mutual || nj[m] == i;
    if (!mutual) atomicAdd(&extra[j], 1u);
}
__global__ void __launch_bounds__(1024) k_sym_scan(const uint32_t *__restrict__ extra, int N, int K, uint32_t *__restrict__ row) {
    // single CTA exclusive scan of (K + extra[i]) -- preprocessing time only
    __shared__ unsigned long long carry;
    __shared__ uint32_t wsum[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < N; base += 1024) {
        const int i = base + threadIdx.x;
        const uint32_t x = i < N ? (uint32_t) K + extra[i] : 0u;
        uint32_t inc = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) wsum[w] = inc;
        __syncthreads();
        uint32_t wb = 0;
        for (int k = 0; k < w; k++) wb += wsum[k];
        if (i < N) row[i] = (uint32_t) carry + wb + inc - x;
        __syncthreads();
        if (threadIdx.x == 1023) carry += wb + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) row[N] = (uint32_t) carry;
}
__global__ void __launch_bounds__(256) k_sym_fill(const uint32_t *__restrict__ nbr, const double *__restrict__ cond, int N, int K,
                                                  const uint32_t *__restrict__ row, uint32_t *__restrict__ cursor, uint32_t *__restrict__ col,
                                                  double *__restrict__ val) {
    const size_t e = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t) N * K) return;
    const uint32_t i = (uint32_t) (e / K), j = nbr[e];
    const int m0 = (int) (e - (size_t) i * K);
    const uint32_t *nj = nbr + (size_t) j * K;
    double v = cond[e];
    bool mutual = false;
    for (int m = 0; m < K; m++) if (nj[m] == i) { mutual = true; v += cond[(size_t) j * K + m]; }
    col[row[i] + m0] = j;
    val[row[i] + m0] = v;
    if (!mutual) {
        const uint32_t slot = row[j] + (uint32_t) K + atomicAdd(&cursor[j], 1u);
        col[slot] = i;
        val[slot] = cond[e];
    }
}
// columns ascending within every row: rank sort by one CTA per row (columns are unique); values halved on the way
__global__ void __launch_bounds__(128) k_sym_sort_rows(const uint32_t *__restrict__ row, const uint32_t *__restrict__ col_in,
                                                       const double *__restrict__ val_in, uint32_t *__restrict__ col_out,
                                                       double *__restrict__ val_out) {
    const uint32_t b = row[blockIdx.x], e = row[blockIdx.x + 1];
    for (uint32_t k = b + threadIdx.x; k < e; k += blockDim.x) {
        const uint32_t c = col_in[k];
        uint32_t r = 0;
        for (uint32_t m = b; m < e; m++) r += col_in[m] < c ? 1u : 0u;
        col_out[b + r] = c;
        val_out[b + r] = val_in[k] / 2.0;
    }
}
__global__ void __launch_bounds__(256) k_sum_partial(const double *__restrict__ v, size_t n, double *__restrict__ partial) {
    __shared__ double sm[8];
    double s = 0;
    const size_t per = (n + gridDim.x - 1) / gridDim.x, b = blockIdx.x * per, e = b + per < n ? b + per : n;
    for (size_t k = b + threadIdx.x; k < e; k += blockDim.x) s += v[k];
    s = warp_sum_d(s);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0; for (int k = 0; k < 8; k++) t += sm[k]; partial[blockIdx.x] = t; }
}
__global__ void __launch_bounds__(256) k_scale(double *__restrict__ v, size_t n, double inv) {
    const size_t k = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) v[k] = v[k] * inv;
}

int select_device(int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); g_prep_error = "no CUDA device: libfitsne_b200 has no CPU fallback"; return FITSNE_ENODEV; }
    if (device >= 0 && cudaSetDevice(device) != cudaSuccess) { g_prep_error = "cudaSetDevice failed"; return FITSNE_ECUDA; }
    return 0;
}

}  // namespace

extern "C" {

const char *fitsne_prep_last_error(void) { return g_prep_error.c_str(); }

int fitsne_knn(const double *X, int N, int D, int K, int device, unsigned int *nbr, double *dist) {
    if (!X || !nbr || !dist || N < 2 || D < 1 || K < 1 || K >= N) { g_prep_error = "fitsne_knn: bad arguments (need 1 <= K < N)"; return FITSNE_EINVAL; }
    if (int rc = select_device(device)) return rc;
    const int KP = std::min(N - 1, K + KNN_MARGIN);
    const size_t smem = (size_t) (KD * KQP + KD * KCP + KQ) * 4 + (size_t) KQ * 4 + (size_t) KQ * KC * 8 + (size_t) KQ * KP * 8 + (size_t) KQ * 4;
    if (smem > 220 * 1024) { g_prep_error = "fitsne_knn: K too large for the shared-memory lists (K + 8 <= ~400)"; return FITSNE_EINVAL; }
    DevBuf dX, dXf, dsq, dcand, dnbr, ddist;
    PCK(cudaMalloc(&dX.p, (size_t) N * D * 8)); PCK(cudaMalloc(&dXf.p, (size_t) N * D * 4)); PCK(cudaMalloc(&dsq.p, (size_t) N * 4));
    PCK(cudaMalloc(&dcand.p, (size_t) N * KP * 4)); PCK(cudaMalloc(&dnbr.p, (size_t) N * K * 4)); PCK(cudaMalloc(&ddist.p, (size_t) N * K * 8));
    PCK(cudaMemcpy(dX.p, X, (size_t) N * D * 8, cudaMemcpyHostToDevice));
    k_to_float_norms<<<(N + 255) / 256, 256>>>(dX.as<double>(), N, D, dXf.as<float>(), dsq.as<float>());
    PCK(cudaFuncSetAttribute(k_knn_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    k_knn_tiles<<<(N + KQ - 1) / KQ, KNN_THREADS, smem>>>(dXf.as<float>(), dsq.as<float>(), N, D, KP, dcand.as<uint32_t>());
    PCK(cudaGetLastError());
    const size_t rsmem = (size_t) 4 * KP * 2 * sizeof(double);
    k_knn_refine<<<(N + 3) / 4, 128, rsmem>>>(dX.as<double>(), N, D, K, KP, dcand.as<uint32_t>(), dnbr.as<uint32_t>(), ddist.as<double>());
    PCK(cudaGetLastError());
    PCK(cudaMemcpy(nbr, dnbr.p, (size_t) N * K * 4, cudaMemcpyDeviceToHost));
    PCK(cudaMemcpy(dist, ddist.p, (size_t) N * K * 8, cudaMemcpyDeviceToHost));
    return 0;
}

int fitsne_similarities(const unsigned int *nbr, const double *dist, int N, int K, double perplexity, double sigma, int list_len,
                        const double *list, int device, unsigned int **row_P, unsigned int **col_P, double **val_P) {
    if (!nbr || !dist || !row_P || !col_P || !val_P || N < 2 || K < 1 || (perplexity == 0 && (list_len < 1 || !list)) ||
        (perplexity < 0 && !(sigma > 0))) { g_prep_error = "fitsne_similarities: bad arguments"; return FITSNE_EINVAL; }
    if (int rc = select_device(device)) return rc;
    const size_t NK = (size_t) N * K;
    DevBuf dnbr, ddist, dcond, dtmp, dlist, dextra, dcursor, drow, dcol, dval, dcol2, dval2, dpart;
    PCK(cudaMalloc(&dnbr.p, NK * 4)); PCK(cudaMalloc(&ddist.p, NK * 8)); PCK(cudaMalloc(&dcond.p, NK * 8));
    PCK(cudaMemcpy(dnbr.p, nbr, NK * 4, cudaMemcpyHostToDevice));
    PCK(cudaMemcpy(ddist.p, dist, NK * 8, cudaMemcpyHostToDevice));
    if (perplexity == 0) {
        PCK(cudaMalloc(&dtmp.p, NK * 8)); PCK(cudaMalloc(&dlist.p, (size_t) list_len * 8));
        PCK(cudaMemcpy(dlist.p, list, (size_t) list_len * 8, cudaMemcpyHostToDevice));
    }
    k_calibrate<<<(N + 3) / 4, 128>>>(ddist.as<double>(), N, K, perplexity, sigma, list_len, dlist.as<double>(), dcond.as<double>(), dtmp.as<double>());
    PCK(cudaGetLastError());
    PCK(cudaMalloc(&dextra.p, (size_t) N * 4)); PCK(cudaMalloc(&dcursor.p, (size_t) N * 4)); PCK(cudaMalloc(&drow.p, ((size_t) N + 1) * 4));
    PCK(cudaMemset(dextra.p, 0, (size_t) N * 4)); PCK(cudaMemset(dcursor.p, 0, (size_t) N * 4));
    const int eb = (int) ((NK + 255) / 256);
    k_sym_count<<<eb, 256>>>(dnbr.as<uint32_t>(), N, K, dextra.as<uint32_t>());
    k_sym_scan<<<1, 1024>>>(dextra.as<uint32_t>(), N, K, drow.as<uint32_t>());
    PCK(cudaGetLastError());
    std::vector<unsigned int> hrow((size_t) N + 1);
    PCK(cudaMemcpy(hrow.data(), drow.p, ((size_t) N + 1) * 4, cudaMemcpyDeviceToHost));
    const size_t E = hrow[N];
    PCK(cudaMalloc(&dcol.p, E * 4)); PCK(cudaMalloc(&dval.p, E * 8)); PCK(cudaMalloc(&dcol2.p, E * 4)); PCK(cudaMalloc(&dval2.p, E * 8));
    k_sym_fill<<<eb, 256>>>(dnbr.as<uint32_t>(), dcond.as<double>(), N, K, drow.as<uint32_t>(), dcursor.as<uint32_t>(), dcol.as<uint32_t>(), dval.as<double>());
    k_sym_sort_rows<<<N, 128>>>(drow.as<uint32_t>(), dcol.as<uint32_t>(), dval.as<double>(), dcol2.as<uint32_t>(), dval2.as<double>());
    PCK(cudaGetLastError());
    const int pb = 1024;
    PCK(cudaMalloc(&dpart.p, pb * 8));
    k_sum_partial<<<pb, 256>>>(dval2.as<double>(), E, dpart.as<double>());
    std::vector<double> hp(pb);
    PCK(cudaMemcpy(hp.data(), dpart.p, pb * 8, cudaMemcpyDeviceToHost));
    double total = 0;
    for (double v : hp) total += v;
    k_scale<<<(int) ((E + 255) / 256), 256>>>(dval2.as<double>(), E, 1.0 / total);
    PCK(cudaGetLastError());
    unsigned int *row = (unsigned int *) malloc(((size_t) N + 1) * sizeof(unsigned int));
    unsigned int *col = (unsigned int *) malloc(E * sizeof(unsigned int));
    double *val = (double *) malloc(E * sizeof(double));
    if (!row || !col || !val) { free(row); free(col); free(val); g_prep_error = "host allocation failed"; return FITSNE_ENOMEM; }
    memcpy(row, hrow.data(), ((size_t) N + 1) * 4);
    cudaError_t e1 = cudaMemcpy(col, dcol2.p, E * 4, cudaMemcpyDeviceToHost), e2 = cudaMemcpy(val, dval2.p, E * 8, cudaMemcpyDeviceToHost);
    if (e1 != cudaSuccess || e2 != cudaSuccess) { free(row); free(col); free(val); return prep_fail(FITSNE_ECUDA, "download of the CSR", e1 != cudaSuccess ? e1 : e2); }
    *row_P = row; *col_P = col; *val_P = val;
    return 0;
}

void fitsne_free(void *p) { free(p); }

}  // extern "C"
