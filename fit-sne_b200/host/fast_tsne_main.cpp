// fast_tsne_main.cpp -- `bin/fast_tsne <version> [data_path] [result_path] [nthreads]`: same command line, version
// handshake and exit codes as the reference's main (/root/reference/src/tsne.cpp:2041-2136), so the unmodified
// fast_tsne.py / fast_tsne.R / fast_tsne.m wrappers drive it unchanged.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <thread>

#include "tsne_host.h"

int main(int argc, char *argv[]) {
    const char version_number[] = "1.2.1";
    printf("=============== t-SNE v%s ===============\n", version_number);
    if (argc < 2) { std::cout << "Please pass version number as first argument." << std::endl; exit(-1); }
    if (strcmp(argv[1], version_number)) { std::cout << "Wrapper passed wrong version number: " << argv[1] << std::endl; exit(-1); }
    const char *data_path = argc >= 3 ? argv[2] : "data.dat";
    const char *result_path = argc >= 4 ? argv[3] : "result.dat";
    unsigned int nthreads = argc >= 5 ? (unsigned int) strtoul(argv[4], nullptr, 10) : 0;
    if (nthreads == 0) nthreads = std::thread::hardware_concurrency();
    std::cout << "fast_tsne data_path: " << data_path << std::endl;
    std::cout << "fast_tsne result_path: " << result_path << std::endl;
    std::cout << "fast_tsne nthreads: " << nthreads << std::endl;

    int N, D, no_dims, max_iter, stop_lying_iter, mom_switch_iter, K, nbody_algo, knn_algo, no_momentum_during_exag;
    int n_trees, search_k, start_late_exag_iter, nterms, min_num_intervals, rand_seed = 0, load_affinities = 0;
    int perplexity_list_length = 0;
    double momentum, final_momentum, learning_rate, max_step_norm, sigma, early_exag_coeff, late_exag_coeff, perplexity, theta;
    double intervals_per_integer, df = 1.0, *data = nullptr, *Y = nullptr, *perplexity_list = nullptr;
    bool skip_random_init = false;
    TSNE tsne;
    if (tsne.load_data(data_path, &data, &Y, &N, &D, &no_dims, &theta, &perplexity, &rand_seed, &max_iter, &stop_lying_iter,
                       &mom_switch_iter, &momentum, &final_momentum, &learning_rate, &K, &sigma, &nbody_algo, &knn_algo,
                       &early_exag_coeff, &no_momentum_during_exag, &n_trees, &search_k, &start_late_exag_iter,
                       &late_exag_coeff, &nterms, &intervals_per_integer, &min_num_intervals, &skip_random_init,
                       &load_affinities, &perplexity_list_length, &perplexity_list, &df, &max_step_norm)) {
        double *costs = (double *) calloc(max_iter, sizeof(double));
        if (!costs) { printf("Memory allocation failed!\n"); exit(1); }
        const int error_code = tsne.run(data, N, D, Y, no_dims, perplexity, theta, rand_seed, skip_random_init, max_iter,
                                        stop_lying_iter, mom_switch_iter, momentum, final_momentum, learning_rate, K, sigma,
                                        nbody_algo, knn_algo, early_exag_coeff, costs, no_momentum_during_exag != 0,
                                        start_late_exag_iter, late_exag_coeff, n_trees, search_k, nterms,
                                        intervals_per_integer, min_num_intervals, nthreads, load_affinities,
                                        perplexity_list_length, perplexity_list, df, max_step_norm);
        if (error_code < 0) exit(error_code);
        tsne.save_data(result_path, Y, costs, N, no_dims, max_iter);
        free(data); free(Y); free(costs); free(perplexity_list);
    }
    printf("Done.\n\n");
    return 0;
}
