// Host emulation of k_spread_chunks2 (tests/test_spread_emul.py builds and runs this with nvcc, no GPU needed).
// The kernel is written as two phase functions (fitsne_kernels.cuh: spread2_load / spread2_chunk); here every thread of a
// block runs phase 1, then every thread runs phase 2 -- exactly what the __syncthreads() in the kernel enforces -- and the
// results (slots + directly written grid nodes) are compared
//   (a) bit for bit with a host transcription of the per-(chunk, node) thread of k_spread_chunks (same segment rules,
//       same summation order), and
//   (b) after a transcription of k_spread_combine, with a direct fp64 spread of the same points.
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

// transcription of the body of k_spread_chunks for one (chunk, node) thread
template <int D, int P>
static void old_thread(int c, int node, const float *sorted_u, const uint32_t *skeys, const uint32_t *box_start, int n,
                       const GridParams &gp, float4 *slots, float2 *fft_in, float2 *compact) {
    const int p = P, nodes = D == 2 ? p * p : p;
    const int kb = c * CHUNK; if (kb >= n) return;
    const int ke = std::min(kb + CHUNK, n);
    const int a = D == 2 ? node / p : node, b = D == 2 ? node - a * p : 0;
    const float sa = gp.s[a], sb = gp.s[b];
    float2 *dst = compact ? compact : fft_in;
    const int Gc = gp.M / 2;
    const size_t stride = compact ? (D == 2 ? (size_t) Gc * Gc : (size_t) Gc) : (D == 2 ? (size_t) gp.M * gp.M : (size_t) gp.M);
    float4 *myslots = slots + (size_t) c * 2 * nodes;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int cur = key_to_box<D>(skeys[kb], gp);
    for (int k = kb; k < ke; k++) {
        const int box = key_to_box<D>(skeys[k], gp);
        if (box != cur) {
            if ((int) box_start[cur] >= kb) store_node(dst, stride, node_offset<D>(cur, node, gp, p, compact != nullptr), acc);
            else myslots[node] = acc;
            acc = make_float4(0.f, 0.f, 0.f, 0.f); cur = box;
        }
        if (D == 2) {
            const float2 u = reinterpret_cast<const float2 *>(sorted_u)[k];
            const float L = lagrange1<P>(gp, p, a, u.y) * lagrange1<P>(gp, p, b, u.x);
            const float ox = u.x - sb, oy = u.y - sa;
            acc.x += L; acc.y += L * ox; acc.z += L * oy; acc.w += L * (ox * ox + oy * oy);
        } else {
            const float u = sorted_u[k];
            const float L = lagrange1<P>(gp, p, a, u);
            const float o = u - sa;
            acc.x += L; acc.y += L * o; acc.z += L * o * o;
        }
    }
    const bool started_here = (int) box_start[cur] >= kb, ends_here = (int) box_start[cur + 1] <= ke;
    if (started_here && ends_here) store_node(dst, stride, node_offset<D>(cur, node, gp, p, compact != nullptr), acc);
    else myslots[(started_here ? 1 : 0) * nodes + node] = acc;
}

// transcription of k_spread_combine's arithmetic for one node of one box (lpn = 1: plain chunk order)
static float4 combine_node(const float4 *slots, int nodes, int node, int s, int e, bool *single) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    *single = false;
    if (e <= s) return acc;
    const int c0 = s / CHUNK, c1 = (e - 1) / CHUNK;
    if (c0 == c1) { *single = true; return acc; }
    for (int c = c0; c <= c1; c++) {
        const float4 v = slots[((size_t) c * 2 + (c == c0 ? 1 : 0)) * nodes + node];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    return acc;
}

template <int D, int P>
static bool run_case(const char *name, int n, int B, int M, double heavy_frac, bool use_compact, unsigned seed) {
    const int p = P, nodes = D == 2 ? p * p : p;
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
    std::vector<uint32_t> box_start(gp.nb + 2, 0);
    {   // boundary detection like k_post_sort
        int prev = -1;
        for (int k = 0; k < n; k++) { const int box = key_to_box<D>(skeys[k], gp); for (int b = prev + 1; b <= box; b++) box_start[b] = k; prev = box; }
        for (int b = prev + 1; b <= gp.nb; b++) box_start[b] = n;
    }
    const int nchunks = (n + CHUNK - 1) / CHUNK;
    const int Gc = M / 2;
    const size_t plane = use_compact ? (D == 2 ? (size_t) Gc * Gc : (size_t) Gc) : (D == 2 ? (size_t) M * M : (size_t) M);
    const float4 sentinel4 = make_float4(-7.f, -7.f, -7.f, -7.f);
    std::vector<float4> slotsA((size_t) nchunks * 2 * nodes, sentinel4), slotsB(slotsA);
    std::vector<float2> gridA(2 * plane, make_float2(-9.f, -9.f)), gridB(gridA);
    float2 *fa = use_compact ? nullptr : gridA.data(), *ca = use_compact ? gridA.data() : nullptr;
    float2 *fb = use_compact ? nullptr : gridB.data(), *cb = use_compact ? gridB.data() : nullptr;
    // A: the new kernel, block by block, phase by phase
    const int nblocks = (nchunks + SP2_THREADS - 1) / SP2_THREADS;
    static Sp2Smem<D> sm;
    for (int blk = 0; blk < nblocks; blk++) {
        memset(&sm, 0xff, sizeof sm);                                   // poison: nothing may be read before it is staged
        for (int t = 0; t < SP2_THREADS; t++) spread2_load<D>(t, blk, su.data(), skeys.data(), n, sm);
        for (int t = 0; t < SP2_THREADS; t++) spread2_chunk<D, P>(t, blk, sm, box_start.data(), n, gp, slotsA.data(), fa, ca);
    }
    // B: the old kernel's threads
    for (int c = 0; c < nchunks; c++) for (int node = 0; node < nodes; node++)
        old_thread<D, P>(c, node, su.data(), skeys.data(), box_start.data(), n, gp, slotsB.data(), fb, cb);
    bool ok = true;
    if (memcmp(slotsA.data(), slotsB.data(), slotsA.size() * sizeof(float4)) != 0) { printf("%s: slots differ\n", name); ok = false; }
    if (memcmp(gridA.data(), gridB.data(), gridA.size() * sizeof(float2)) != 0) { printf("%s: direct grid writes differ\n", name); ok = false; }
    // combine + direct fp64 reference
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
        bool single;
        float4 v = combine_node(slotsA.data(), nodes, node, (int) box_start[box], (int) box_start[box + 1], &single);
        if (single) {
            const size_t off = node_offset<D>(box, node, gp, p, use_compact);
            const float2 p0 = gridA[off], p1 = gridA[plane + off];
            v = make_float4(p0.x, p0.y, p1.x, p1.y);
        }
        const double *r = &ref[((size_t) box * nodes + node) * 4];
        const double got[4] = {v.x, v.y, v.z, v.w};
        for (int q = 0; q < 4; q++) { max_err = std::max(max_err, std::fabs(got[q] - r[q])); max_ref = std::max(max_ref, std::fabs(r[q])); }
    }
    const double rel = max_err / std::max(max_ref, 1e-30);
    if (!(rel < 2e-5)) { printf("%s: combined grid vs fp64 reference: max err %.3e of %.3e\n", name, max_err, max_ref); ok = false; }
    printf("%-34s n=%7d B=%4d M=%5d heavy=%.2f compact=%d : %s (vs fp64 %.1e)\n", name, n, B, M, heavy_frac, (int) use_compact, ok ? "ok" : "FAILED", rel);
    return ok;
}

int main() {
    bool ok = true;
    ok &= run_case<2, 3>("2-D p=3 late (small boxes)", 50000, 50, 320, 0.0, false, 1);
    ok &= run_case<2, 3>("2-D p=3 ragged tail", 50000 + 17, 36, 224, 0.0, false, 2);
    ok &= run_case<2, 3>("2-D p=3 heavy box", 40000, 25, 160, 0.6, false, 3);
    ok &= run_case<2, 3>("2-D p=3 one box", 5000, 25, 160, 1.0, false, 4);
    ok &= run_case<2, 3>("2-D p=3 compact (sharded)", 30011, 50, 320, 0.1, true, 5);
    ok &= run_case<2, 2>("2-D p=2", 20000, 60, 256, 0.0, false, 6);
    ok &= run_case<2, 4>("2-D p=4", 20000, 30, 256, 0.2, false, 7);
    ok &= run_case<1, 3>("1-D p=3", 60000, 138, 864, 0.0, false, 8);
    ok &= run_case<1, 5>("1-D p=5 heavy", 33333, 50, 512, 0.5, false, 9);
    ok &= run_case<1, 3>("1-D p=3 compact", 60000, 138, 864, 0.0, true, 10);
    ok &= run_case<2, 3>("2-D p=3 tiny n", 7, 25, 160, 0.0, false, 11);
    ok &= run_case<2, 3>("2-D p=3 n = 1 block + 1", SP2_POINTS + 1, 50, 320, 0.0, false, 12);
    if (ok) printf("SPREAD_EMUL_OK\n");
    return ok ? 0 : 1;
}
