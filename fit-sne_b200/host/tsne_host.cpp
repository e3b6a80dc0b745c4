// tsne_host.cpp -- see tsne_host.h.  Written from scratch against the reference's observable behaviour:
//   file protocol   data.dat reader  == /root/reference/src/tsne.cpp:1915-1985 (writer: fast_tsne.py:259-297)
//                   result.dat writer == tsne.cpp:2024-2038 (reader: fast_tsne.py:311-328)
//                   P_row/P_col/P_val.dat (load_affinities 1 / 2) == tsne.cpp:236-281,334-366
//   prologue        perplexity check :128-131, zero-mean + max-abs normalisation :153-161, K = 3*perplexity
//                   :287-304, symmetrise + normalise :325-330, seeded Box-Muller init :369-387,1880-1890
//   similarities    Gaussian kernel with per-point bandwidth by bisection on the entropy, tol 1e-5, <=200 steps
//                   (:1394-1469), averaged over a perplexity list (:1474-1500), over the K nearest neighbours
// Neighbours are found exactly (blocked brute force over all pairs, multi-threaded) -- the same neighbours the
// reference's VP-tree option (knn_algo=2) returns; Annoy's approximate search (knn_algo=1) is not reproduced.
#include "tsne_host.h"

#include <algorithm>
#include <chrono>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <thread>
#include <vector>

#ifndef FITSNE_HOST_ONLY
#include "fitsne_b200.h"
#endif

namespace {

template <typename F>
void parallel_rows(unsigned nthreads, int n, F f) {
    if (nthreads <= 1 || n < 64) { f(0, n); return; }
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nthreads; t++) {
        const int b = (int) ((long long) n * t / nthreads), e = (int) ((long long) n * (t + 1) / nthreads);
        th.emplace_back([=] { f(b, e); });
    }
    for (auto &x : th) x.join();
}

// Row of conditional similarities over the K neighbour distances `d` (Euclidean).
// Reference quirk kept on purpose: the sparse path hands Euclidean (not squared) distances to
// distances2similarities with ifSquared=false (tsne.cpp:1607,1692), but the perplexity-list overload it goes through
// forwards them with ifSquared=TRUE (tsne.cpp:1479,1486,1489), so the kernel actually evaluated is exp(-beta*d),
// not exp(-beta*d^2).  The same holds for the fixed-sigma branch.
void calibrate_row(const double *d, int K, double perplexity, double sigma, double *p) {
    double beta, sum = DBL_MIN;
    if (perplexity > 0) {
        double lo = -DBL_MAX, hi = DBL_MAX;
        const double target = std::log(perplexity), tol = 1e-5;
        beta = 1.0;
        for (int it = 0; it < 200; it++) {
            sum = DBL_MIN;
            double h = 0;
            for (int m = 0; m < K; m++) { p[m] = std::exp(-beta * d[m]); sum += p[m]; }
            for (int m = 0; m < K; m++) h += beta * (d[m] * p[m]);
            const double diff = h / sum + std::log(sum) - target;
            if (diff < tol && -diff < tol) break;
            if (diff > 0) { lo = beta; beta = (hi == DBL_MAX || hi == -DBL_MAX) ? beta * 2.0 : (beta + hi) / 2.0; }
            else { hi = beta; beta = (lo == -DBL_MAX || lo == DBL_MAX) ? beta / 2.0 : (beta + lo) / 2.0; }
        }
        // p and sum belong to the last beta that was TESTED (tsne.cpp:1418-1454 normalises those, :1466)
    } else {
        beta = 1 / (2 * sigma * sigma);
        for (int m = 0; m < K; m++) { p[m] = std::exp(-beta * d[m]); sum += p[m]; }
    }
    for (int m = 0; m < K; m++) p[m] /= sum;
}

double randn_ref() {   // Marsaglia polar method on rand(), as tsne.cpp:1880-1890 (same stream for a given seed)
    double x, y, r;
    do {
        x = 2 * (rand() / ((double) RAND_MAX + 1)) - 1;
        y = 2 * (rand() / ((double) RAND_MAX + 1)) - 1;
        r = x * x + y * y;
    } while (r >= 1.0 || r == 0.0);
    return x * std::sqrt(-2 * std::log(r) / r);
}

bool read_exact(FILE *f, void *dst, size_t size, size_t count) { return fread(dst, size, count, f) == count; }

}  // namespace

void TSNE::zero_mean(double *X, int N, int D) {
    std::vector<double> mean(D, 0.0);
    for (int n = 0; n < N; n++) for (int d = 0; d < D; d++) mean[d] += X[(size_t) n * D + d];
    for (int d = 0; d < D; d++) mean[d] /= (double) N;
    for (int n = 0; n < N; n++) for (int d = 0; d < D; d++) X[(size_t) n * D + d] -= mean[d];
}

int TSNE::input_similarities(const double *X, int N, int D, double perplexity, int K, double sigma, int list_len,
                             const double *list, unsigned int nthreads, unsigned int **row_out, unsigned int **col_out,
                             double **val_out) {
    if (K >= N) { printf("K (%d) must be smaller than the number of points (%d)\n", K, N); return -1; }
    if (perplexity > K) printf("Perplexity should be lower than K!\n");
    printf("Exact kNN search (K=%d) on %u threads...\n", K, nthreads);
    std::vector<unsigned int> nbr((size_t) N * K);
    std::vector<double> cond((size_t) N * K);
    std::vector<double> sq(N);
    for (int i = 0; i < N; i++) { double s = 0; for (int d = 0; d < D; d++) s += X[(size_t) i * D + d] * X[(size_t) i * D + d]; sq[i] = s; }
    parallel_rows(nthreads, N, [&](int b, int e) {
        std::vector<std::pair<double, int>> cand(N);
        std::vector<double> dist(K), tmp(K), acc(K);
        for (int i = b; i < e; i++) {
            const double *xi = X + (size_t) i * D;
            for (int j = 0; j < N; j++) {
                const double *xj = X + (size_t) j * D;
                double dot = 0;
                for (int d = 0; d < D; d++) dot += xi[d] * xj[d];
                cand[j] = {std::max(0.0, sq[i] + sq[j] - 2 * dot), j};
            }
            cand[i].first = -1;   // self sorts first and is skipped, like the [1..K] slice of a K+1 query
            std::partial_sort(cand.begin(), cand.begin() + K + 1, cand.end());
            for (int m = 0; m < K; m++) {
                // exact distance for the kept neighbours (the expansion above loses digits for near-duplicates)
                const double *xj = X + (size_t) cand[m + 1].second * D;
                double s = 0;
                for (int d = 0; d < D; d++) s += (xi[d] - xj[d]) * (xi[d] - xj[d]);
                dist[m] = std::sqrt(s);
                nbr[(size_t) i * K + m] = (unsigned int) cand[m + 1].second;
            }
            double *p = &cond[(size_t) i * K];
            if (perplexity != 0) calibrate_row(dist.data(), K, perplexity, sigma, p);
            else {   // average over the perplexity list
                calibrate_row(dist.data(), K, list[0], sigma, acc.data());
                for (int l = 1; l < list_len; l++) {
                    calibrate_row(dist.data(), K, list[l], sigma, tmp.data());
                    for (int m = 0; m < K; m++) acc[m] += tmp[m];
                }
                for (int m = 0; m < K; m++) p[m] = acc[m] / list_len;
            }
        }
    });
    // symmetrise: P_sym = (P + P^T) / 2 as CSR, then normalise to sum 1
    printf("Symmetrizing...\n");
    struct Ent { unsigned int r, c; double v; };
    std::vector<Ent> ent;
    ent.reserve((size_t) 2 * N * K);
    for (int i = 0; i < N; i++)
        for (int m = 0; m < K; m++) {
            const unsigned int j = nbr[(size_t) i * K + m];
            const double v = cond[(size_t) i * K + m];
            ent.push_back({(unsigned) i, j, v});
            if (j != (unsigned) i) ent.push_back({j, (unsigned) i, v});
        }
    std::sort(ent.begin(), ent.end(), [](const Ent &a, const Ent &b) { return a.r != b.r ? a.r < b.r : a.c < b.c; });
    size_t w = 0;
    for (size_t k = 0; k < ent.size(); k++) {
        if (w > 0 && ent[w - 1].r == ent[k].r && ent[w - 1].c == ent[k].c) ent[w - 1].v += ent[k].v;
        else ent[w++] = ent[k];
    }
    ent.resize(w);
    unsigned int *row = (unsigned int *) calloc((size_t) N + 1, sizeof(unsigned int));
    unsigned int *col = (unsigned int *) malloc(w * sizeof(unsigned int));
    double *val = (double *) malloc(w * sizeof(double));
    if (!row || !col || !val) { printf("Memory allocation failed!\n"); exit(1); }
    double total = 0;
    for (size_t k = 0; k < w; k++) { row[ent[k].r + 1]++; col[k] = ent[k].c; val[k] = ent[k].v / 2.0; total += val[k]; }
    for (int i = 0; i < N; i++) row[i + 1] += row[i];
    for (size_t k = 0; k < w; k++) val[k] /= total;
    *row_out = row; *col_out = col; *val_out = val;
    return 0;
}

bool TSNE::load_data(const char *data_path, double **data, double **Y, int *n, int *d, int *no_dims, double *theta,
                     double *perplexity, int *rand_seed, int *max_iter, int *stop_lying_iter, int *mom_switch_iter,
                     double *momentum, double *final_momentum, double *learning_rate, int *K, double *sigma,
                     int *nbody_algo, int *knn_algo, double *early_exag_coeff, int *no_momentum_during_exag,
                     int *n_trees, int *search_k, int *start_late_exag_iter, double *late_exag_coeff, int *nterms,
                     double *intervals_per_integer, int *min_num_intervals, bool *skip_random_init,
                     int *load_affinities, int *perplexity_list_length, double **perplexity_list, double *df,
                     double *max_step_norm) {
    FILE *h = fopen(data_path, "rb");
    if (!h) { printf("Error: could not open data file.\n"); return false; }
    bool ok = read_exact(h, n, sizeof(int), 1) && read_exact(h, d, sizeof(int), 1) && read_exact(h, theta, sizeof(double), 1) &&
              read_exact(h, perplexity, sizeof(double), 1);
    *perplexity_list_length = 0;   // the reference leaves these uninitialised when perplexity != 0 (tsne.cpp:1928-1929)
    *perplexity_list = nullptr;
    if (ok && *perplexity == 0) {
        ok = read_exact(h, perplexity_list_length, sizeof(int), 1) && *perplexity_list_length > 0;
        if (ok) {
            *perplexity_list = (double *) malloc(*perplexity_list_length * sizeof(double));
            if (!*perplexity_list) { printf("Memory allocation failed!\n"); exit(1); }
            ok = read_exact(h, *perplexity_list, sizeof(double), *perplexity_list_length);
        }
    }
    ok = ok && read_exact(h, no_dims, sizeof(int), 1) && read_exact(h, max_iter, sizeof(int), 1) &&
         read_exact(h, stop_lying_iter, sizeof(int), 1) && read_exact(h, mom_switch_iter, sizeof(int), 1) &&
         read_exact(h, momentum, sizeof(double), 1) && read_exact(h, final_momentum, sizeof(double), 1) &&
         read_exact(h, learning_rate, sizeof(double), 1) && read_exact(h, max_step_norm, sizeof(double), 1) &&
         read_exact(h, K, sizeof(int), 1) && read_exact(h, sigma, sizeof(double), 1) &&
         read_exact(h, nbody_algo, sizeof(int), 1) && read_exact(h, knn_algo, sizeof(int), 1) &&
         read_exact(h, early_exag_coeff, sizeof(double), 1) && read_exact(h, no_momentum_during_exag, sizeof(int), 1) &&
         read_exact(h, n_trees, sizeof(int), 1) && read_exact(h, search_k, sizeof(int), 1) &&
         read_exact(h, start_late_exag_iter, sizeof(int), 1) && read_exact(h, late_exag_coeff, sizeof(double), 1) &&
         read_exact(h, nterms, sizeof(int), 1) && read_exact(h, intervals_per_integer, sizeof(double), 1) &&
         read_exact(h, min_num_intervals, sizeof(int), 1);
    if (!ok) { printf("Error: data file is truncated.\n"); fclose(h); return false; }
    if (*nbody_algo == 2 && *no_dims > 2) {
        printf("FFT interpolation scheme supports only 1 or 2 output dimensions, not %d\n", *no_dims);
        exit(1);
    }
    const size_t nd = (size_t) *n * (size_t) *d;
    *data = (double *) malloc(nd * sizeof(double));
    if (!*data) { printf("Memory allocation failed!\n"); exit(1); }
    if (!read_exact(h, *data, sizeof(double), nd)) { printf("Error: data file is truncated.\n"); fclose(h); return false; }
    // optional tail (tsne.cpp:1962-1985): seed, df, load_affinities, then an N x no_dims initialisation
    if (!read_exact(h, rand_seed, sizeof(int), 1)) { /* keep the caller's default */ }
    else if (!read_exact(h, df, sizeof(double), 1)) { /* keep default */ }
    else if (!read_exact(h, load_affinities, sizeof(int), 1)) { /* keep default */ }
    const size_t ny = (size_t) *n * (size_t) *no_dims;
    *Y = (double *) malloc(ny * sizeof(double));
    if (!*Y) { printf("Memory allocation failed!\n"); exit(1); }
    *skip_random_init = fread(*Y, sizeof(double), ny, h) == ny;
    fclose(h);
    printf("Read the following parameters:\n\t n %d by d %d dataset, theta %lf,\n\t perplexity %lf, no_dims %d, max_iter %d,\n"
           "\t stop_lying_iter %d, mom_switch_iter %d,\n\t momentum %lf, final_momentum %lf,\n\t learning_rate %lf, max_step_norm %lf,\n"
           "\t K %d, sigma %lf, nbody_algo %d,\n\t knn_algo %d, early_exag_coeff %lf,\n\t no_momentum_during_exag %d, n_trees %d, search_k %d,\n"
           "\t start_late_exag_iter %d, late_exag_coeff %lf\n\t nterms %d, interval_per_integer %lf, min_num_intervals %d, t-dist df %lf\n",
           *n, *d, *theta, *perplexity, *no_dims, *max_iter, *stop_lying_iter, *mom_switch_iter, *momentum, *final_momentum,
           *learning_rate, *max_step_norm, *K, *sigma, *nbody_algo, *knn_algo, *early_exag_coeff, *no_momentum_during_exag,
           *n_trees, *search_k, *start_late_exag_iter, *late_exag_coeff, *nterms, *intervals_per_integer, *min_num_intervals, *df);
    printf("Read the %i x %i data matrix successfully. X[0,0] = %lf\n", *n, *d, (*data)[0]);
    if (*perplexity == 0) {
        printf("Read the list of perplexities: ");
        for (int m = 0; m < *perplexity_list_length; m++) printf("%f ", (*perplexity_list)[m]);
        printf("\n");
    }
    if (*skip_random_init) printf("Read the initialization successfully.\n");
    return true;
}

void TSNE::save_data(const char *result_path, double *data, double *costs, int n, int d, int max_iter) {
    FILE *h = fopen(result_path, "wb");
    if (!h) { printf("Error: could not open data file.\n"); return; }
    fwrite(&n, sizeof(int), 1, h);
    fwrite(&d, sizeof(int), 1, h);
    fwrite(data, sizeof(double), (size_t) n * d, h);
    fwrite(&max_iter, sizeof(int), 1, h);
    fwrite(costs, sizeof(double), max_iter, h);
    fclose(h);
    printf("Wrote the %i x %i data matrix successfully.\n", n, d);
}

#ifndef FITSNE_HOST_ONLY
static bool write_file(const char *name, const void *src, size_t size, size_t count) {
    FILE *h = fopen(name, "wb");
    if (!h) { printf("Error: could not open data file.\n"); return false; }
    const bool ok = fwrite(src, size, count, h) == count;
    fclose(h);
    return ok;
}

int TSNE::run(double *X, int N, int D, double *Y, int no_dims, double perplexity, double theta, int rand_seed,
              bool skip_random_init, int max_iter, int stop_lying_iter, int mom_switch_iter, double momentum,
              double final_momentum, double learning_rate, int K, double sigma, int nbody_algorithm, int knn_algo,
              double early_exag_coeff, double *costs, bool no_momentum_during_exag, int start_late_exag_iter,
              double late_exag_coeff, int n_trees, int search_k, int nterms, double intervals_per_integer,
              int min_num_intervals, unsigned int nthreads, int load_affinities, int perplexity_list_length,
              double *perplexity_list, double df, double max_step_norm) {
    (void) n_trees; (void) search_k;
    if (N - 1 < 3 * perplexity) { printf("Perplexity too large for the number of data points!\n"); exit(1); }
    printf(no_momentum_during_exag ? "No momentum during the exaggeration phase.\n" : "Will use momentum during exaggeration phase\n");
    if (theta == .0 || nbody_algorithm != 2) {
        printf("Error: this build accelerates the FFT-interpolation path only (nbody_algo=2, theta>0); "
               "exact and Barnes-Hut modes are not part of it.\n");
        exit(2);
    }
    if (knn_algo != 1 && knn_algo != 2) { printf("Invalid knn_algo param\n"); exit(1); }
    const auto t_pre = std::chrono::steady_clock::now();
    printf("Computing input similarities...\n");
    zero_mean(X, N, D);
    if (perplexity > 0 || perplexity_list_length > 0) {
        printf("Using perplexity, so normalizing input data (to prevent numerical problems)\n");
        double mx = .0;
        for (size_t i = 0; i < (size_t) N * D; i++) mx = std::max(mx, std::fabs(X[i]));
        for (size_t i = 0; i < (size_t) N * D; i++) X[i] /= mx;
    } else printf("Not using perplexity, so data are left un-normalized.\n");

    unsigned int *row_P = nullptr, *col_P = nullptr;
    double *val_P = nullptr;
    const bool stream_files = load_affinities == 1;          // P_row/P_col/P_val.dat go from disk straight to the device
    if (stream_files) {
        printf("Loading approximate input similarities from files...\n");
    } else {
        int K_to_use;
        double sigma_to_use;
        if (perplexity < 0) {
            printf("Using manually set kernel width\n");
            K_to_use = K; sigma_to_use = sigma;
        } else {
            printf("Using perplexity, not the manually set kernel width.  K (number of nearest neighbors) and sigma (bandwidth) parameters are going to be ignored.\n");
            if (perplexity > 0) K_to_use = (int) 3 * perplexity;
            else {
                K_to_use = (int) 3 * perplexity_list[0];
                for (int pp = 1; pp < perplexity_list_length; pp++) K_to_use = std::max(K_to_use, (int) (3 * perplexity_list[pp]));
            }
            sigma_to_use = -1;
        }
        if (knn_algo == 1) printf("Note: Annoy's approximate search is replaced by an exact kNN search in this build.\n");
        // kNN, perplexity search and symmetrisation run on the device (fitsne_knn / fitsne_similarities: the same arithmetic
        // as input_similarities() above, which stays as the CPU statement of it -- FITSNE_HOST_PREP=1 selects it)
        int rc;
        if (getenv("FITSNE_HOST_PREP") && atoi(getenv("FITSNE_HOST_PREP")) != 0) {
            rc = input_similarities(X, N, D, perplexity, K_to_use, sigma_to_use, perplexity_list_length, perplexity_list, nthreads,
                                    &row_P, &col_P, &val_P);
        } else {
            if (K_to_use >= N) { printf("K (%d) must be smaller than the number of points (%d)\n", K_to_use, N); return -1; }
            if (perplexity > K_to_use) printf("Perplexity should be lower than K!\n");
            printf("Exact kNN search (K=%d) on the device...\n", K_to_use);
            std::vector<unsigned int> nbr((size_t) N * K_to_use);
            std::vector<double> dist((size_t) N * K_to_use);
            rc = fitsne_knn(X, N, D, K_to_use, -1, nbr.data(), dist.data());
            if (rc == 0) {
                printf("Perplexity search and symmetrisation on the device...\n");
                rc = fitsne_similarities(nbr.data(), dist.data(), N, K_to_use, perplexity, sigma_to_use, perplexity_list_length,
                                         perplexity_list, -1, &row_P, &col_P, &val_P);
            }
            if (rc != 0) { printf("Error: device preprocessing failed (%d): %s\n", rc, fitsne_prep_last_error()); return rc - 100; }
        }
        if (rc < 0) return rc;
    }
    if (load_affinities == 2) {
        printf("Saving approximate input similarities to files...\n");
        const size_t numel = row_P[N];
        if (!write_file("P_val.dat", val_P, sizeof(double), numel) || !write_file("P_col.dat", col_P, sizeof(unsigned int), numel) ||
            !write_file("P_row.dat", row_P, sizeof(unsigned int), (size_t) N + 1)) return -2;
    }
    if (!skip_random_init) {
        if (rand_seed >= 0) { printf("Using random seed: %d\n", rand_seed); srand((unsigned int) rand_seed); }
        else { printf("Using current time as random seed...\n"); srand(time(NULL)); }
        printf("Randomly initializing the solution.\n");
        for (size_t i = 0; i < (size_t) N * no_dims; i++) Y[i] = randn_ref() * .0001;
        printf("Y[0] = %lf\n", Y[0]);
    } else printf("Using the given initialization.\n");
    preprocessing_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_pre).count();
    if (row_P) printf("Input similarities computed (sparsity = %f)!\n", (double) row_P[N] / ((double) N * (double) N));
    printf("Learning embedding...\n");
    printf("Using FIt-SNE approximation (B200 build: %s).\n", fitsne_version());

    fitsne_config cfg;
    cfg.nterms = nterms; cfg.intervals_per_integer = intervals_per_integer; cfg.min_num_intervals = min_num_intervals;
    cfg.df = df; cfg.device = -1; cfg.flags = 0;
    fitsne_schedule s;
    s.max_iter = max_iter; s.stop_lying_iter = stop_lying_iter; s.mom_switch_iter = mom_switch_iter;
    s.start_late_exag_iter = start_late_exag_iter; s.momentum = momentum; s.final_momentum = final_momentum;
    s.learning_rate = learning_rate; s.early_exag_coeff = early_exag_coeff; s.late_exag_coeff = late_exag_coeff;
    s.max_step_norm = max_step_norm; s.no_momentum_during_exag = no_momentum_during_exag ? 1 : 0; s.verbose = 1;
    const auto t_loop = std::chrono::steady_clock::now();
    const int rc = stream_files ? fitsne_run_files(&cfg, &s, nullptr, N, no_dims, Y, costs)
                                : fitsne_run_host(&cfg, &s, N, no_dims, row_P, col_P, val_P, Y, costs);
    if (stream_files && rc == FITSNE_EINVAL) { printf("Error: could not open data file.\n"); return -2; }      // like tsne.cpp:186
    loop_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_loop).count();
    free(row_P); free(col_P); free(val_P);
    if (rc != 0) {
        printf("Error: the CUDA gradient loop failed (%d): %s\n", rc, fitsne_last_error(nullptr));
        return rc < 0 ? rc - 100 : -100;   // distinct from the reference's -1 / -2
    }
    printf("Preprocessing %.2f s, gradient loop %.2f s (B200, incl. transfers)\n", preprocessing_seconds, loop_seconds);
    return 0;
}
#endif  // FITSNE_HOST_ONLY

// ---- CPU-only C entry points for the protocol tests (libfitsne_host.so) -------------------------------------------------
extern "C" {
int fitsne_host_parse(const char *data_path, int *ints, double *dbls, double **X, double **Y, double **perplexity_list) {
    TSNE t;
    int n, d, no_dims, rand_seed = 0, max_iter, stop_lying_iter, mom_switch_iter, K, nbody_algo, knn_algo, nomom, n_trees,
        search_k, start_late, nterms, min_int, load_aff = 0, pll = 0;
    double theta, perplexity, momentum, final_momentum, lr, sigma, early, late, ipi, df = 1.0, msn;
    bool skip = false;
    if (!t.load_data(data_path, X, Y, &n, &d, &no_dims, &theta, &perplexity, &rand_seed, &max_iter, &stop_lying_iter,
                     &mom_switch_iter, &momentum, &final_momentum, &lr, &K, &sigma, &nbody_algo, &knn_algo, &early, &nomom,
                     &n_trees, &search_k, &start_late, &late, &nterms, &ipi, &min_int, &skip, &load_aff, &pll,
                     perplexity_list, &df, &msn)) return -1;
    const int iv[20] = {n, d, no_dims, max_iter, stop_lying_iter, mom_switch_iter, K, nbody_algo, knn_algo, nomom, n_trees,
                        search_k, start_late, nterms, min_int, rand_seed, load_aff, pll, skip ? 1 : 0, 0};
    const double dv[16] = {theta, perplexity, momentum, final_momentum, lr, msn, sigma, early, late, ipi, df, 0, 0, 0, 0, 0};
    memcpy(ints, iv, sizeof iv);
    memcpy(dbls, dv, sizeof dv);
    return 0;
}
int fitsne_host_write_result(const char *result_path, const double *Y, const double *costs, int n, int d, int max_iter) {
    TSNE t;
    t.save_data(result_path, const_cast<double *>(Y), const_cast<double *>(costs), n, d, max_iter);
    return 0;
}
int fitsne_host_similarities(const double *X, int N, int D, double perplexity, int K, double sigma, int list_len,
                             const double *list, unsigned int nthreads, unsigned int **row_P, unsigned int **col_P,
                             double **val_P) {
    return TSNE::input_similarities(X, N, D, perplexity, K, sigma, list_len, list, nthreads, row_P, col_P, val_P);
}
void fitsne_host_free(void *p) { free(p); }
}
