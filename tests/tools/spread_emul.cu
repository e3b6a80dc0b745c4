// Host emulation of the spread (tests/test_kernel_emulation.py builds and runs this with nvcc, no GPU needed).
// k_spread_chunks is written as two phase functions (fitsne_kernels.cuh: spread2_chunk / spread2_stitch); here every thread
// of a block runs the walk, then every thread runs the stitch -- exactly what the __syncthreads() in the kernel enforces --
// onto a pre-zeroed grid; then the work list is combined with k_spread_combine's own per-lane function and tree order.
// Checked: the grid against a direct fp64 spread of the same points; box_range[] (a by-product of the chunk walk) against
// plain boundary detection; the work list against "boxes that span more than one chunk"; compile-time and run-time node
// counts (P > 0 / P == 0) bit for bit against each other.
// Test infrastructure only; prints "SPREAD_EMUL_OK" when every configuration passes.
#include "../../fit-sne_b200/csrc/fitsne_kernels.cuh"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>
using namespace fk;

static GridParams make_gp(int D, int B, int p, int M) {
    GridParams gp;
    memset(&gp, 0, sizeof gp);
    gp.B = B; gp.p = p; gp.G = B * p; gp.M = M; gp.ok = 1;
    int xb = 0; while ((1 << xb) < B) xb++;
    gp.xbits = xb; gp.nb = D == 2 ? B * B : B;
    double s[PMAX]; const double hh = 1.0 / p; s[0] = hh / 2;
    for (int i = 1; i < p; i++) s[i] = s[i - 1] + hh;
    for (int i = 0; i < p; i++) {
        double den = 1; for (int j = 0; j < p; j++) if (i != j) den *= s[i] - s[j];
        gp.s[i] = (float) s[i]; gp.inv_den[i] = (float) (1.0 / den);
    }
    return gp;
}


static int g_chunk = CHUNK;      // chunk length under test (the device uses CHUNK; FITSNE_CHUNK overrides it for sweeps)

template <int D, int P>
static void run_spread(int n, const GridParams &gp, const std::vector<uint32_t> &skeys, const std::vector<float> &su,
                       std::vector<float> &grid, std::vector<uint2> &box_range, std::vector<uint32_t> &work, size_t gsize) {
    const int p = gp.p, nodes = D == 2 ? p * p : p;
    const int nchunks = (n + g_chunk - 1) / g_chunk;
    const int nblocks = (nchunks + SP2_THREADS - 1) / SP2_THREADS;
    std::vector<float4> cslots((size_t) nblocks * 2 * nodes, make_float4(NAN, NAN, NAN, NAN));
    std::vector<float4> part((size_t) SP2_THREADS * 2 * nodes);
    grid.assign(gsize, 0.f);                                             // the memset node of the iteration
    box_range.assign(gp.nb + 2, make_uint2(0xdeadbeefu, 0xdeadbeefu));
    work.assign((size_t) nblocks + 2, 0u);
    static Sp2Meta meta;
    for (int blk = 0; blk < nblocks; blk++) {
        std::fill(part.begin(), part.end(), make_float4(NAN, NAN, NAN, NAN));   // poison: only parked partials may be read
        for (int t = 0; t < SP2_THREADS; t++) spread2_chunk<D, P>(t, blk, su.data(), skeys.data(), n, gp, part.data(), meta, grid.data(), box_range.data(), g_chunk);
        for (int t = 0; t < SP2_THREADS; t++) spread2_stitch<D, P>(t, blk, n, gp, part.data(), meta, grid.data(), cslots.data(), work.data(), g_chunk);
    }
    // k_spread_combine: one lane per (box, node) in plain CTA order; COMBINE_COOP CTAs or more: 32 lanes + xor-shuffle tree
    for (uint32_t e = 0; e < work[0]; e++) {
        const int box = (int) work[1 + e];
        const int s = (int) box_range[box].x, en = (int) box_range[box].y;
        const int c0 = s / (SP2_THREADS * g_chunk), c1 = (en - 1) / (SP2_THREADS * g_chunk);
        for (int node = 0; node < nodes; node++) {
            if (c1 - c0 + 1 < 32) {
                store_node<D>(grid.data(), (size_t) gp.M, node_offset<D>(box, node, gp, p), combine_node_lane<D>(cslots.data(), c0, c1, nodes, node, 0, 1));
                continue;
            }
            float4 lane[32];
            for (int l = 0; l < 32; l++) lane[l] = combine_node_lane<D>(cslots.data(), c0, c1, nodes, node, l, 32);
            for (int o = 16; o > 0; o >>= 1) {
                float4 nxt[32];
                for (int l = 0; l < 32; l++) nxt[l] = make_float4(lane[l].x + lane[l ^ o].x, lane[l].y + lane[l ^ o].y, lane[l].z + lane[l ^ o].z, lane[l].w + lane[l ^ o].w);
                for (int l = 0; l < 32; l++) lane[l] = nxt[l];
            }
            store_node<D>(grid.data(), (size_t) gp.M, node_offset<D>(box, node, gp, p), lane[0]);
        }
    }
}

template <int D, int P>
static bool run_case(const char *name, int n, int B, int M, double heavy_frac, unsigned seed, int p_runtime = 0) {
    const int p = P > 0 ? P : p_runtime, nodes = D == 2 ? p * p : p;
    GridParams gp = make_gp(D, B, p, M);
    std::mt19937 rng(seed);
    std::uniform_real_distribution<float> U(0.f, 1.f);
    // points: a fraction sits in one heavy box (early-exaggeration-like), the rest uniformly in all boxes
    std::vector<uint32_t> keys(n); std::vector<float> u((size_t) n * D);
    const uint32_t heavy = D == 2 ? (((uint32_t) (B / 3)) << gp.xbits) | (uint32_t) (B / 2) : (uint32_t) (B / 2);
    for (int i = 0; i < n; i++) {
        if (U(rng) < heavy_frac) keys[i] = heavy;
        else {
            const uint32_t bx = std::min(B - 1, (int) (U(rng) * B)), by = std::min(B - 1, (int) (U(rng) * B));
            keys[i] = D == 2 ? (by << gp.xbits) | bx : bx;
        }
        for (int d = 0; d < D; d++) u[(size_t) i * D + d] = U(rng);
    }
    std::vector<int> order(n); for (int i = 0; i < n; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return keys[x] < keys[y]; });
    std::vector<uint32_t> skeys(n); std::vector<float> su((size_t) n * D);
    for (int k = 0; k < n; k++) { skeys[k] = keys[order[k]]; for (int d = 0; d < D; d++) su[(size_t) k * D + d] = u[(size_t) order[k] * D + d]; }
    std::vector<uint32_t> bs_ref(gp.nb + 2, 0);
    {   // plain boundary detection
        int prev = -1;
        for (int k = 0; k < n; k++) { const int box = key_to_box<D>(skeys[k], gp); for (int b = prev + 1; b <= box; b++) bs_ref[b] = k; prev = box; }
        for (int b = prev + 1; b <= gp.nb; b++) bs_ref[b] = n;
    }
    const size_t gsize = D == 2 ? (size_t) gp.G * gp.G * 4 : (size_t) M * 4;     // floats: G^2 float4 / two packed complex lines
    std::vector<float> grid, grid0;
    std::vector<uint2> box_range, bs0;
    std::vector<uint32_t> work, work0;
    run_spread<D, P>(n, gp, skeys, su, grid, box_range, work, gsize);
    bool ok = true;
    for (int b = 0; b < gp.nb; b++) if (bs_ref[b + 1] > bs_ref[b] && (box_range[b].x != bs_ref[b] || box_range[b].y != bs_ref[b + 1])) {
        printf("%s: box_range[%d] = [%u, %u), expected [%u, %u)\n", name, b, box_range[b].x, box_range[b].y, bs_ref[b], bs_ref[b + 1]); ok = false; break; }
    {   // work list == boxes crossing a CTA boundary (one CTA = SP2_THREADS chunks of sorted points)
        std::vector<uint32_t> want, got(work.begin() + 1, work.begin() + 1 + work[0]);
        for (int b = 0; b < gp.nb; b++) if (bs_ref[b + 1] > bs_ref[b] && bs_ref[b] / (SP2_THREADS * g_chunk) != (bs_ref[b + 1] - 1) / (SP2_THREADS * g_chunk)) want.push_back(b);
        std::sort(got.begin(), got.end());
        if (got != want) { printf("%s: work list has %zu boxes, expected %zu\n", name, got.size(), want.size()); ok = false; }
    }
    if (P > 0) {   // the run-time-p instantiation must give the same bits
        run_spread<D, 0>(n, gp, skeys, su, grid0, bs0, work0, gsize);
        if (memcmp(grid.data(), grid0.data(), grid.size() * sizeof(float)) != 0) { printf("%s: P=%d and run-time p differ\n", name, P); ok = false; }
    }
    // direct fp64 reference
    double max_err = 0, max_ref = 0;
    std::vector<double> ref((size_t) gp.nb * nodes * 4, 0.0);
    for (int k = 0; k < n; k++) {
        const int box = key_to_box<D>(skeys[k], gp);
        for (int node = 0; node < nodes; node++) {
            const int a = D == 2 ? node / p : node, b = D == 2 ? node % p : 0;
            double L, ox, oy = 0;
            auto lag = [&](int j, double x) { double v = 1; for (int q = 0; q < p; q++) if (q != j) v *= (x - (q + 0.5) / p) / ((j + 0.5) / p - (q + 0.5) / p); return v; };
            if (D == 2) { const double ux = su[(size_t) k * 2], uy = su[(size_t) k * 2 + 1]; L = lag(a, uy) * lag(b, ux); ox = ux - (b + 0.5) / p; oy = uy - (a + 0.5) / p; }
            else { const double ux = su[k]; L = lag(a, ux); ox = ux - (a + 0.5) / p; }
            double *r = &ref[((size_t) box * nodes + node) * 4];
            r[0] += L; r[1] += L * ox;
            if (D == 2) { r[2] += L * oy; r[3] += L * (ox * ox + oy * oy); } else r[2] += L * ox * ox;
        }
    }
    for (int box = 0; box < gp.nb; box++) for (int node = 0; node < nodes; node++) {
        const size_t off = node_offset<D>(box, node, gp, p);
        double got[4];
        if (D == 2) for (int q = 0; q < 4; q++) got[q] = grid[off * 4 + q];
        else { got[0] = grid[off * 2]; got[1] = grid[off * 2 + 1]; got[2] = grid[((size_t) M + off) * 2]; got[3] = grid[((size_t) M + off) * 2 + 1]; }
        const double *r = &ref[((size_t) box * nodes + node) * 4];
        for (int q = 0; q < 4; q++) { max_err = std::max(max_err, std::fabs(got[q] - r[q])); max_ref = std::max(max_ref, std::fabs(r[q])); }
    }
    const double rel = max_err / std::max(max_ref, 1e-30);
    if (!(rel < 2e-5)) { printf("%s: grid vs fp64 reference: max err %.3e of %.3e\n", name, max_err, max_ref); ok = false; }
    printf("%-34s n=%7d B=%4d M=%5d heavy=%.2f : %s (vs fp64 %.1e, %u boxes cross a CTA boundary)\n", name, n, B, M, heavy_frac, ok ? "ok" : "FAILED", rel, work[0]);
    return ok;
}

int main() {
    bool ok = true;
    for (int chunk : {8, 4, 2, 1}) {      // the kernel takes the chunk length at run time
    g_chunk = chunk;
    ok &= run_case<2, 3>("2-D p=3 late (small boxes)", 50000, 50, 320, 0.0, 1);
    ok &= run_case<2, 3>("2-D p=3 ragged tail", 50000 + 17, 36, 224, 0.0, 2);
    ok &= run_case<2, 3>("2-D p=3 heavy box", 40000, 25, 160, 0.6, 3);
    ok &= run_case<2, 3>("2-D p=3 very heavy box (coop)", 150000, 25, 160, 0.9, 14);
    ok &= run_case<2, 3>("2-D p=3 one box", 5000, 25, 160, 1.0, 4);
    ok &= run_case<2, 3>("2-D p=3 sparse (empty boxes)", 3011, 200, 1280, 0.1, 5);
    ok &= run_case<2, 2>("2-D p=2", 20000, 60, 256, 0.0, 6);
    ok &= run_case<2, 4>("2-D p=4", 20000, 30, 256, 0.2, 7);
    ok &= run_case<2, 0>("2-D p=5 (run-time p)", 9000, 30, 320, 0.2, 13, 5);
    ok &= run_case<1, 3>("1-D p=3", 60000, 138, 864, 0.0, 8);
    ok &= run_case<1, 5>("1-D p=5 heavy", 33333, 50, 512, 0.5, 9);
    ok &= run_case<1, 0>("1-D p=7 (run-time p)", 20000, 60, 864, 0.0, 10, 7);
    ok &= run_case<2, 3>("2-D p=3 tiny n", 7, 25, 160, 0.0, 11);
    ok &= run_case<2, 3>("2-D p=3 n = 1 block + 1", SP2_THREADS * g_chunk + 1, 50, 320, 0.0, 12);
    }
    if (ok) printf("SPREAD_EMUL_OK\n");
    return ok ? 0 : 1;
}
