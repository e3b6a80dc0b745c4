// tsne_host.h -- host shell of the B200 build: the reference's TSNE class surface for the FFT path
// (/root/reference/src/tsne.h:40-58: run / load_data / save_data, same argument lists and meaning), re-hosted
// from scratch.  run() does the reference's prologue (tsne.cpp:118-436) on the host -- centring/normalising X,
// building or loading the CSR P, initialising Y -- and hands the whole iteration loop (tsne.cpp:437-577) to
// libfitsne_b200 through the C ABI (fitsne_run_host).  Barnes-Hut (nbody_algo=1) and exact (theta=0) modes are
// other algorithms and are not part of this build: run() reports that and exits non-zero.
#ifndef FITSNE_B200_TSNE_HOST_H
#define FITSNE_B200_TSNE_HOST_H

class TSNE {
public:
    int run(double *X, int N, int D, double *Y, int no_dims, double perplexity, double theta, int rand_seed,
            bool skip_random_init, int max_iter, int stop_lying_iter, int mom_switch_iter, double momentum,
            double final_momentum, double learning_rate, int K, double sigma, int nbody_algorithm, int knn_algo,
            double early_exag_coeff, double *costs, bool no_momentum_during_exag, int start_late_exag_iter,
            double late_exag_coeff, int n_trees, int search_k, int nterms, double intervals_per_integer,
            int min_num_intervals, unsigned int nthreads, int load_affinities, int perplexity_list_length,
            double *perplexity_list, double df, double max_step_norm);

    bool load_data(const char *data_path, double **data, double **Y, int *n, int *d, int *no_dims, double *theta,
                   double *perplexity, int *rand_seed, int *max_iter, int *stop_lying_iter, int *mom_switch_iter,
                   double *momentum, double *final_momentum, double *learning_rate, int *K, double *sigma,
                   int *nbody_algo, int *knn_algo, double *early_exag_coeff, int *no_momentum_during_exag,
                   int *n_trees, int *search_k, int *start_late_exag_iter, double *late_exag_coeff, int *nterms,
                   double *intervals_per_integer, int *min_num_intervals, bool *skip_random_init,
                   int *load_affinities, int *perplexity_list_length, double **perplexity_list, double *df,
                   double *max_step_norm);

    void save_data(const char *result_path, double *data, double *costs, int n, int d, int max_iter);

    // Host preprocessing (stays on the CPU, timed separately from the loop): exact kNN + perplexity calibration
    // + symmetrisation -> CSR.  Public so that tests can call it without a GPU.
    static int input_similarities(const double *X, int N, int D, double perplexity, int K, double sigma,
                                  int perplexity_list_length, const double *perplexity_list, unsigned int nthreads,
                                  unsigned int **row_P, unsigned int **col_P, double **val_P);
    static void zero_mean(double *X, int N, int D);

    double preprocessing_seconds = 0, loop_seconds = 0;
};

// C entry points of libfitsne_host.so (CPU only; used by the protocol tests)
extern "C" {
int fitsne_host_parse(const char *data_path, int *ints /*[20]*/, double *dbls /*[16]*/, double **X, double **Y,
                      double **perplexity_list);
int fitsne_host_write_result(const char *result_path, const double *Y, const double *costs, int n, int d, int max_iter);
int fitsne_host_similarities(const double *X, int N, int D, double perplexity, int K, double sigma, int list_len,
                             const double *list, unsigned int nthreads, unsigned int **row_P, unsigned int **col_P,
                             double **val_P);
void fitsne_host_free(void *p);
}
#endif
