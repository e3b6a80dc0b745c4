/*
 * fitsne_b200.h -- C ABI of libfitsne_b200.so: FIt-SNE's per-iteration gradient loop on one (or, sharded by
 * points, several) NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary.  Every entry point replaces a piece of the reference's TSNE::run iteration
 * body (reference citations are /root/reference/src/...); INTEGRATION.md shows the ~40-line patch a
 * maintainer applies to tsne.cpp to call it.  Plain pointers and sizes only; no C++/torch types; no
 * exceptions cross the boundary.  All functions return 0 on success or a negative FITSNE_E* code;
 * fitsne_last_error() gives the message.  One caller thread per context.  There is NO CPU fallback: without
 * a CUDA device fitsne_create fails with FITSNE_ENODEV.
 *
 * Host arrays stay caller-owned and use the reference's own types: CSR P as (unsigned int row_P[N+1],
 * unsigned int col_P[E], double val_P[E]) -- tsne.cpp:168-170 -- and Y as row-major double[N*no_dims].
 * The device keeps fp32 copies; Y values that are not fp32-representable are rounded once on upload.
 */
#ifndef FITSNE_B200_H
#define FITSNE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FITSNE_OK 0
#define FITSNE_EINVAL (-1)    /* bad argument (no_dims not 1/2, nterms out of range, ...)        */
#define FITSNE_ENODEV (-2)    /* no usable CUDA device / wrong architecture                      */
#define FITSNE_ECUDA (-3)     /* a CUDA call failed                                              */
#define FITSNE_ENOMEM (-4)    /* device or host allocation failed                                */
#define FITSNE_ENCCL (-5)     /* NCCL could not be loaded or a collective failed                 */
#define FITSNE_ESTATE (-6)    /* call sequence error (e.g. KL before any gradient)               */

typedef struct fitsne_ctx fitsne_ctx;

/* Interpolation / kernel parameters: the trailing arguments of computeFftGradient* (tsne.cpp:1027-1029,
 * :860-862, :758-760, :647-649) plus the device to run on. */
typedef struct fitsne_config {
    int nterms;                   /* n_interpolation_points per box and axis (reference default 3), 1..16 */
    double intervals_per_integer; /* reference default 1                                                  */
    int min_num_intervals;        /* reference default 50                                                 */
    double df;                    /* t-kernel degrees of freedom; 1.0 selects the Cauchy-kernel paths     */
    int device;                   /* CUDA device ordinal, -1 = current device                             */
    int flags;                    /* FITSNE_FLAG_* bit set                                                */
} fitsne_config;

#define FITSNE_FLAG_NO_GRAPH 1    /* launch kernels individually instead of replaying CUDA graphs         */
#define FITSNE_FLAG_TIMERS 2      /* record per-phase CUDA-event timers (adds synchronisation)            */
#define FITSNE_FLAG_NO_REORDER 4  /* never re-order the points on the device (keeps the plain CSR kernel)  */
#define FITSNE_FLAG_FORCE_TILES 8 /* always use the tiled attractive kernel after a re-ordering            */
#define FITSNE_FLAG_NO_TILES 16   /* re-order for locality but keep the CSR attractive kernel              */
#define FITSNE_FLAG_NO_SPECULATION 32 /* fitsne_run: one host round trip per iteration instead of batches  */

/* One optimiser step's parameters: the state TSNE::run carries across iterations (tsne.cpp:437-544). */
typedef struct fitsne_step_params {
    double exaggeration;  /* multiplier on P in force: early_exag_coeff, late_exag_coeff or 1 (tsne.cpp:404-412,534-543) */
    double momentum;      /* current momentum (tsne.cpp:495,544)                                           */
    double learning_rate; /* eta (tsne.cpp:495)                                                            */
    double max_step_norm; /* <=0: no clipping (tsne.cpp:498-511)                                           */
    int mode;             /* FITSNE_STEP_*                                                                 */
} fitsne_step_params;

#define FITSNE_STEP_MOMENTUM_CLIP 0 /* gains + momentum + optional clipping   (tsne.cpp:492-513) */
#define FITSNE_STEP_MOMENTUM 1      /* gains + momentum, never clipped        (tsne.cpp:481-485) */
#define FITSNE_STEP_PLAIN_GD 2      /* Y -= dY, no learning rate, no gains    (tsne.cpp:489)     */

/* The whole schedule of TSNE::run's loop (tsne.cpp:118-125 arguments that matter after preprocessing). */
typedef struct fitsne_schedule {
    int max_iter;
    int stop_lying_iter;
    int mom_switch_iter;
    int start_late_exag_iter;
    double momentum;
    double final_momentum;
    double learning_rate;
    double early_exag_coeff;      /* 0 = automatic: 1/(learning_rate * max row sum of P), tsne.cpp:392-402 */
    double late_exag_coeff;
    double max_step_norm;
    int no_momentum_during_exag;
    int verbose;                  /* print the reference's "Iteration k (50 iterations in ...)" lines      */
} fitsne_schedule;

/* Timers / counters filled by fitsne_get_stats (all times in milliseconds of device time). */
typedef struct fitsne_stats {
    uint64_t iterations;          /* optimiser steps executed                                   */
    uint64_t kernel_launches;     /* our kernels launched (graph nodes count)                    */
    uint64_t graph_launches;
    uint64_t regrids;             /* iterations whose grid (n_boxes) differed from the previous  */
    int n_boxes;                  /* last grid: boxes per dimension                              */
    int grid_side;                /* last grid: nterms * n_boxes                                 */
    int fft_side;                 /* last FFT length per dimension                               */
    double min_coord, max_coord;  /* last bounds used for the grid                               */
    double phase_ms[16];          /* FITSNE_PHASE_* accumulators (only with FITSNE_FLAG_TIMERS)  */
    uint64_t reorders;            /* device-side Morton re-orderings of the points                */
} fitsne_stats;

enum {
    FITSNE_PHASE_BOUNDS = 0, FITSNE_PHASE_SORT, FITSNE_PHASE_SPREAD, FITSNE_PHASE_KERNEL_SPECTRUM,
    FITSNE_PHASE_FFT, FITSNE_PHASE_GATHER, FITSNE_PHASE_ATTRACT_UPDATE, FITSNE_PHASE_CENTER, FITSNE_PHASE_KL,
    FITSNE_PHASE_COLLECTIVES /* grid all-reduce */, FITSNE_PHASE_ALLGATHER /* Y all-gather */, FITSNE_PHASE_COUNT
};

/* ---- lifetime ------------------------------------------------------------------------------------- */

/* Create a context holding P (converted to fp32 CSR) and the optimiser state (Y, uY=0, gains=1) on the
 * device.  Replaces the allocations at tsne.cpp:143-150.  Y0 may be NULL (zeros; set it later).
 * For a sharded run every rank passes the FULL row_P / Y0 but may pass only its own
 * rows' col_P/val_P slice -- see fitsne_create_sharded. */
int fitsne_create(const fitsne_config *cfg, int N, int no_dims, const unsigned int *row_P,
                  const unsigned int *col_P, const double *val_P, const double *Y0, fitsne_ctx **out);

/* Sharded variant: this rank owns points/rows [row_begin, row_end); col_P/val_P hold only the edges of
 * those rows (row_P is still the full N+1 offsets array).  nccl_unique_id is the 128-byte ncclUniqueId
 * produced by fitsne_nccl_unique_id on rank 0 and shipped by the launcher (torch.distributed, MPI, ...).
 * Rows are dealt out in contiguous blocks of ceil(N / world_size): rank r owns [r * per, min(N, (r + 1) * per)), which
 * must not be empty (FITSNE_EINVAL otherwise -- e.g. N = 9 on 4 ranks leaves rank 3 without rows: use fewer ranks);
 * 1 <= world_size <= 8, all ranks on one node (peer-memory exchanges over NVLink; NCCL collectives without peer access). */
int fitsne_create_sharded(const fitsne_config *cfg, int N, int no_dims, const unsigned int *row_P,
                          const unsigned int *col_P_local, const double *val_P_local, const double *Y0,
                          int rank, int world_size, int row_begin, int row_end,
                          const void *nccl_unique_id, fitsne_ctx **out);
int fitsne_nccl_unique_id(void *out_128_bytes);

/* P as files -- the reference's load_affinities side files (tsne.cpp:236-281 reads, :334-366 writes): <dir>/P_row.dat
 * (u32 x N+1), P_col.dat (u32 x E), P_val.dat (f64 x E).  dir == NULL or "": $FITSNE_AFFINITIES_DIR, else the current
 * directory (where the reference looks).  The edges are streamed to the device in bounded chunks through pinned
 * memory: no host copy of col/val is ever made; a sharded rank reads only its own rows' slice of the files. */
int fitsne_create_from_files(const fitsne_config *cfg, const char *dir, int N, int no_dims, const double *Y0, fitsne_ctx **out);
int fitsne_create_from_files_sharded(const fitsne_config *cfg, const char *dir, int N, int no_dims, const double *Y0, int rank,
                                     int world_size, const void *nccl_unique_id, fitsne_ctx **out);

int fitsne_destroy(fitsne_ctx *ctx);
const char *fitsne_last_error(const fitsne_ctx *ctx); /* ctx may be NULL: last create() failure */

/* ---- state transfer ------------------------------------------------------------------------------- */
int fitsne_set_Y(fitsne_ctx *ctx, const double *Y);           /* N*no_dims doubles, host */
int fitsne_get_Y(fitsne_ctx *ctx, double *Y);
int fitsne_set_optimizer_state(fitsne_ctx *ctx, const double *uY, const double *gains); /* NULL = keep */
int fitsne_get_optimizer_state(fitsne_ctx *ctx, double *uY, double *gains);             /* NULL = skip */

/* ---- the hot path --------------------------------------------------------------------------------- */

/* dC for the current Y, no state change.  Replaces computeFftGradient (tsne.cpp:1027-1169),
 * computeFftGradientVariableDf (:860-1024), computeFftGradientOneD (:758-856) and
 * computeFftGradientOneDVariableDf (:647-754) -- dispatch on no_dims/df as TSNE::run does (:446-464).
 * dC_out: N*no_dims doubles (host), may be NULL.  sum_Q_out: the reference's current_sum_Q, may be NULL.
 * `exaggeration` multiplies P (the reference pre-scales val_P in place, :410-411). */
int fitsne_gradient(fitsne_ctx *ctx, double exaggeration, double *dC_out, double *sum_Q_out);

/* One full iteration on device-resident state: gradient, gains/momentum/clip update, Y += uY, zero-mean
 * (tsne.cpp:446-531).  Asynchronous: returns after enqueueing; any getter synchronises. */
int fitsne_step(fitsne_ctx *ctx, const fitsne_step_params *p);

/* KL divergence with the sum_Q of the most recent gradient and the CURRENT Y (the reference evaluates it
 * after the update, tsne.cpp:547-555), including the exaggeration correction (:563-568 is applied by the
 * caller or by fitsne_run; this function returns the raw sum like evaluateErrorFft, :1329-1355). */
int fitsne_kl(fitsne_ctx *ctx, double exaggeration, double *C_out);

/* The loop of TSNE::run (tsne.cpp:389-577) on the context's P and current Y: exaggeration schedule,
 * momentum switch, KL every 50 iterations into costs[max_iter] (host, pre-zeroed by the caller like
 * tsne.cpp:2112; may be NULL).  Y_out (host, N*no_dims doubles) receives the final embedding; may be NULL. */
int fitsne_run(fitsne_ctx *ctx, const fitsne_schedule *s, double *costs, double *Y_out);

/* Convenience: the call tsne.cpp's TSNE::run makes after preprocessing when patched per INTEGRATION.md:
 * host CSR P + host Y in, host Y + costs out, context created and destroyed inside. */
int fitsne_run_host(const fitsne_config *cfg, const fitsne_schedule *s, int N, int no_dims,
                    const unsigned int *row_P, const unsigned int *col_P, const double *val_P, double *Y,
                    double *costs);

/* The same with P taken from files (fitsne_create_from_files): what the host shell calls for load_affinities == 1. */
int fitsne_run_files(const fitsne_config *cfg, const fitsne_schedule *s, const char *dir, int N, int no_dims, double *Y,
                     double *costs);

/* Prepare (twiddle tables, buffers) every FFT length a grid of n_boxes_lo..n_boxes_hi boxes per dimension can need,
 * so that no allocation happens inside the iteration loop.  Optional and cheap (the FFTs are our own kernels: there
 * are no library plans to create); lengths are otherwise prepared on first use. */
int fitsne_prewarm(fitsne_ctx *ctx, int n_boxes_lo, int n_boxes_hi);

/* ---- the step before the loop, on the device (optional; the host shell and the Python mirror use it) -- */

/* Exact Euclidean kNN of every row of X (row-major double[N*D], host) on the device: nbr[N*K] neighbour indices and
 * dist[N*K] distances (ascending, ties by index, the point itself excluded), both host arrays.  Replaces the reference's
 * Annoy / VP-tree searches (tsne.cpp:1535-1639, :1643-1726): the same neighbours as its exact VP-tree option. */
int fitsne_knn(const double *X, int N, int D, int K, int device, unsigned int *nbr, double *dist);

/* Conditional similarities by perplexity search (perplexity > 0), their average over a perplexity list (perplexity == 0)
 * or a fixed bandwidth (perplexity < 0, sigma), then symmetrisation and normalisation to sum 1 -- computeGaussianPerplexity
 * (tsne.cpp:1394-1500) + symmetrizeMatrix (:1730-1828) -- on the device.  Returns the CSR the loop consumes (columns
 * ascending) in malloc'ed host arrays; release them with fitsne_free. */
int fitsne_similarities(const unsigned int *nbr, const double *dist, int N, int K, double perplexity, double sigma,
                        int perplexity_list_length, const double *perplexity_list, int device, unsigned int **row_P,
                        unsigned int **col_P, double **val_P);
void fitsne_free(void *p);
const char *fitsne_prep_last_error(void);

/* ---- introspection -------------------------------------------------------------------------------- */
int fitsne_synchronize(fitsne_ctx *ctx);
int fitsne_get_stats(fitsne_ctx *ctx, fitsne_stats *out);
int fitsne_reset_stats(fitsne_ctx *ctx);
/* Device time of the last fitsne_run in ms (CUDA events around the loop, KL included). */
int fitsne_last_run_ms(fitsne_ctx *ctx, double *ms);
/* Copy internal device arrays to the host for tests: what = "frep" (N*no_dims floats, F_rep/Z of the last
 * gradient), "perm" (N u32, box-sorted order), "keys" (N u32), "box_range" (first, end per non-empty box), and in 2-D "grid" (G*G float4: the spread
 * result w1, delta_x, delta_y, wbb) and "pot" (G*G float4: v1, Bx, By, 0 at the nodes). */
int fitsne_debug_copy(fitsne_ctx *ctx, const char *what, void *dst, size_t dst_bytes, size_t *needed_bytes);
const char *fitsne_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FITSNE_B200_H */
