// Host emulation of the column-sorted attractive term (k_sorted_count / k_sorted_fill / k_attract_sorted,
// fitsne_kernels.cuh): the kernels' own phase functions run for every lane / thread on the host; the result is compared
// with a direct fp64 evaluation of  attr_i = sum_j p_ij (y_i - y_j) / (1 + |y_i - y_j|^2 / df)  over the CSR.
// Also checks the layout itself: every edge placed exactly once, column blocks non-decreasing inside a row chunk.
// Test infrastructure; prints SPMV_EMUL_OK.
#include "../../fit-sne_b200/csrc/fitsne_kernels.cuh"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>
using namespace fk;

template <int D>
static bool run_case(const char *name, int n, int deg, int col_shift, double df, unsigned seed) {
    std::mt19937 rng(seed);
    std::uniform_real_distribution<float> U(0.f, 1.f);
    std::vector<uint32_t> row(n + 1, 0), col;
    std::vector<float> val;
    for (int i = 0; i < n; i++) {
        const int k = i % 7 == 0 ? 0 : 1 + (int) (U(rng) * 2 * deg);          // ragged rows, some empty
        for (int e = 0; e < k; e++) {
            int j = U(rng) < 0.8f ? i + (int) ((U(rng) - 0.5f) * 600) : (int) (U(rng) * n);   // mostly local, some far
            j = std::min(n - 1, std::max(0, j));
            col.push_back((uint32_t) j); val.push_back(U(rng) / (float) n);
        }
        row[i + 1] = (uint32_t) col.size();
    }
    const size_t E = col.size();
    std::vector<float> Y((size_t) n * D);
    for (auto &v : Y) v = (U(rng) - 0.5f) * 40.f;
    SortedGeom g; g.col_shift = col_shift; g.nchunks = (n + SRT_ROWS - 1) / SRT_ROWS; g.ncb = (n + (1 << col_shift) - 1) >> col_shift;
    const size_t nt = (size_t) g.nchunks * g.ncb;
    std::vector<uint32_t> cnt(nt + 1, 0), start(nt + 1, 0), cur(nt + 1, 0), pack(E + 1, 0xffffffffu);
    std::vector<float> val2(E + 1, NAN);
    std::vector<uint2> edges(E + 1);                                   // the device's edge words: (column, fp32 weight bits)
    for (size_t e = 0; e < E; e++) { uint32_t bits; memcpy(&bits, &val[e], 4); edges[e] = make_uint2(col[e], bits); }
    for (int r = 0; r < n; r++) for (int sub = 0; sub < 8; sub++) sorted_count_lane(r, sub, row.data(), edges.data(), g, cnt.data());
    uint32_t run = 0;
    for (size_t t = 0; t < nt; t++) { start[t] = run; run += cnt[t]; }
    start[nt] = run;
    bool ok = run == E;
    for (int r = 0; r < n; r++) for (int sub = 0; sub < 8; sub++) sorted_fill_lane(r, sub, row.data(), edges.data(), g, start.data(), cur.data(), pack.data(), val2.data());
    for (size_t t = 0; t < nt; t++) ok = ok && cur[t] == cnt[t];
    // layout: inside a chunk the column blocks never decrease; rows belong to the chunk
    for (int rc = 0; rc < g.nchunks; rc++) {
        uint32_t prev = 0;
        for (uint32_t e = start[(size_t) rc * g.ncb]; e < start[(size_t) (rc + 1) * g.ncb]; e++) {
            const uint32_t rl = pack[e] >> SRT_COL_BITS, c = pack[e] & ((1u << SRT_COL_BITS) - 1u);
            ok = ok && (c >> col_shift) >= prev && (int) (rc * SRT_ROWS + rl) < n && (int) c < n;
            prev = c >> col_shift;
        }
    }
    // max row sum -> fixed-point scale, as reorder_points does
    double mx = 0;
    for (int i = 0; i < n; i++) { double s = 0; for (uint32_t e = row[i]; e < row[i + 1]; e++) s += val[e]; mx = std::max(mx, s); }
    const float fix32 = (float) (1073741824.0 / std::max(mx, 1e-300));
    const float inv_df = (float) (1.0 / df);
    std::vector<float> attr((size_t) n * D, NAN);
    static SrtSmem<D> sm;
    for (int rc = 0; rc < g.nchunks; rc++) {
        memset(&sm, 0xff, sizeof sm);
        for (int t = 0; t < SRT_THREADS; t++) attract_sorted_load<D>(t, SRT_THREADS, rc, Y.data(), n, sm);
        for (int t = 0; t < SRT_THREADS; t++) attract_sorted_edges<D>(t, SRT_THREADS, rc, Y.data(), g, start.data(), pack.data(), val2.data(), inv_df, fix32, sm);
        for (int t = 0; t < SRT_THREADS; t++) attract_sorted_store<D>(t, SRT_THREADS, rc, n, fix32, sm, attr.data());
    }
    double num = 0, den = 0, worst = 0;
    for (int i = 0; i < n; i++) {
        double a[2] = {0, 0};
        for (uint32_t e = row[i]; e < row[i + 1]; e++) {
            double d[2], d2 = 0;
            for (int k = 0; k < D; k++) { d[k] = (double) Y[(size_t) i * D + k] - (double) Y[(size_t) col[e] * D + k]; d2 += d[k] * d[k]; }
            const double q = (double) val[e] / (1.0 + d2 / df);
            for (int k = 0; k < D; k++) a[k] += q * d[k];
        }
        const double budget = ((double) (row[i + 1] - row[i]) + 1.0) / (double) fix32;      // half an ulp of the fixed point per edge, + the final rounding
        for (int k = 0; k < D; k++) {
            const double err = std::fabs((double) attr[(size_t) i * D + k] - a[k]);
            worst = std::max(worst, err / (budget + 1e-6 * std::fabs(a[k])));
            num += err * err; den += a[k] * a[k];
        }
    }
    ok = ok && worst < 1.0;
    printf("%-28s n=%6d E=%8zu shift=%2d: %s  (rel-L2 %.2e, worst error / fixed-point budget %.2f)\n", name, n, E, col_shift, ok ? "ok" : "FAILED", std::sqrt(num / den), worst);
    return ok;
}

int main() {
    bool ok = true;
    ok &= run_case<2>("2-D three chunks", 10000, 12, 6, 1.0, 1);
    ok &= run_case<2>("2-D fine column blocks", 10000, 12, 3, 1.0, 2);
    ok &= run_case<2>("2-D coarse column blocks", 9000, 8, 12, 0.5, 3);
    ok &= run_case<2>("2-D exactly one chunk", SRT_ROWS, 10, 6, 1.0, 4);
    ok &= run_case<2>("2-D chunk + 1 row", SRT_ROWS + 1, 10, 6, 1.0, 5);
    ok &= run_case<1>("1-D", 12345, 15, 6, 0.5, 6);
    ok &= run_case<1>("1-D tiny", 37, 3, 5, 1.0, 7);
    if (ok) printf("SPMV_EMUL_OK\n");
    return ok ? 0 : 1;
}
