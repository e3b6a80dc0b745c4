/*
 * fitsne_oracle.c -- CPU restatement (fp64, plain C) of FIt-SNE's per-iteration
 * gradient loop.  TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg may load this.  The product path
 * (fit-sne_b200/) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this file against
 * golden vectors produced by the unmodified, compiled reference
 * (oracle/_ref/libfitsne_ref.so, generator tests/golden/make_golden.py) -- the
 * reference itself ships no tests or golden vectors.
 *
 * What is restated (all citations are /root/reference/src/...):
 *   bounds / grid sizing        tsne.cpp:1039-1049,1065-1077 (2-D), :767-774 (1-D)
 *   box geometry, nodes         nbodyfft.cpp:15-46 (2-D), :257-280 (1-D)
 *   kernel sampling             nbodyfft.cpp:52-61 (2-D), :286-298 (1-D); kernels tsne.cpp:69-94
 *   point->box, in-box coords   nbodyfft.cpp:85-114 (2-D), :349-364 (1-D)
 *   Lagrange basis              nbodyfft.cpp:310-336
 *   spread / gather             nbodyfft.cpp:129-147,222-239 (2-D), :373-382,439-447 (1-D)
 *   circulant convolution       nbodyfft.cpp:170-209 (2-D), :387-430 (1-D)
 *   charges, sum_Q, combine     tsne.cpp:1052-1062,1101-1112,1148-1155 (2-D df=1)
 *                               tsne.cpp:901-955,986-994            (2-D df!=1)
 *                               tsne.cpp:777-818,846-850            (1-D df=1)
 *                               tsne.cpp:665-713,736-741            (1-D df!=1)
 *   attractive term             tsne.cpp:1121-1137 / :965-980 / :826-837 / :721-732
 *   KL                          tsne.cpp:1329-1355, correction :563-568
 *   optimiser step, zero-mean   tsne.cpp:479-531, :1851-1876, sign() tsne.h:37
 *   schedule                    tsne.cpp:404-412,534-547
 *
 * The one deliberate difference: the reference embeds the G-point grid in a
 * 2G-point circulant and calls FFTW; the convolution it evaluates is the plain
 * linear (Toeplitz) convolution  v[i] = sum_a K(|i-a|) w[a], which does not
 * depend on the FFT length.  Here it is evaluated with a self-contained
 * radix-2 FFT of length M = nextpow2(2G) (no FFT library needed), which agrees
 * with the reference to ~1e-15 relative.
 */
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

typedef struct { double re, im; } cplx;

/* ------------------------------------------- tiny fork/join helper (pthreads) -- */
typedef void (*range_fn)(int begin, int end, void *ctx);
typedef struct { range_fn fn; void *ctx; int b, e; } par_job;
static void *par_tramp(void *a) { par_job *j = (par_job *) a; j->fn(j->b, j->e, j->ctx); return NULL; }
static int g_threads = 0;
void fitsne_oracle_set_threads(int n) { g_threads = n; }
static void par_for(int n, range_fn fn, void *ctx) {
    int nt = g_threads > 0 ? g_threads : (int) sysconf(_SC_NPROCESSORS_ONLN);
    if (nt > 64) nt = 64;
    if (nt < 2 || n < 2 * nt) { fn(0, n, ctx); return; }
    pthread_t th[64]; par_job jobs[64];
    for (int t = 0; t < nt; t++) {
        jobs[t].fn = fn; jobs[t].ctx = ctx;
        jobs[t].b = (int) ((long long) n * t / nt); jobs[t].e = (int) ((long long) n * (t + 1) / nt);
        pthread_create(&th[t], NULL, par_tramp, &jobs[t]);
    }
    for (int t = 0; t < nt; t++) pthread_join(th[t], NULL);
}

/* ------------------------------------------------------------------ FFT -- */

static void fft_pow2(cplx *a, int n, int stride, int inverse, const cplx *tw /* n/2 forward twiddles */) {
    /* in-place iterative radix-2, elements a[k*stride] */
    for (int i = 1, j = 0; i < n; i++) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { cplx t = a[(size_t) i * stride]; a[(size_t) i * stride] = a[(size_t) j * stride]; a[(size_t) j * stride] = t; }
    }
    for (int len = 2; len <= n; len <<= 1) {
        int half = len >> 1, step = n / len;
        for (int i = 0; i < n; i += len) {
            for (int k = 0; k < half; k++) {
                cplx w = tw[k * step];
                if (inverse) w.im = -w.im;
                cplx *u = &a[(size_t) (i + k) * stride], *v = &a[(size_t) (i + k + half) * stride];
                cplx t = { v->re * w.re - v->im * w.im, v->re * w.im + v->im * w.re };
                v->re = u->re - t.re; v->im = u->im - t.im;
                u->re += t.re; u->im += t.im;
            }
        }
    }
}

static cplx *make_twiddles(int n) {
    cplx *tw = (cplx *) malloc(sizeof(cplx) * (size_t) (n / 2 + 1));
    for (int k = 0; k < n / 2; k++) {
        double ang = -2.0 * M_PI * (double) k / (double) n;
        tw[k].re = cos(ang); tw[k].im = sin(ang);
    }
    return tw;
}

typedef struct { cplx *a; int M, inverse, cols; const cplx *tw; } fft2_ctx;
static void fft2_range(int b, int e, void *vc) {
    fft2_ctx *c = (fft2_ctx *) vc;
    for (int k = b; k < e; k++) {
        if (c->cols) fft_pow2(c->a + k, c->M, c->M, c->inverse, c->tw);
        else fft_pow2(c->a + (size_t) k * c->M, c->M, 1, c->inverse, c->tw);
    }
}
static void fft2_pow2(cplx *a, int M, int inverse, const cplx *tw) {
    fft2_ctx c = { a, M, inverse, 0, tw };
    par_for(M, fft2_range, &c);
    c.cols = 1;
    par_for(M, fft2_range, &c);
}

static int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

/* ----------------------------------------------------- reference kernels -- */
/* tsne.cpp:69-94.  kind 0: (1+r2)^-2 ; 1: (1+r2/df)^-df ; 2: (1+r2/df)^-(df+1) */
static double kernel_r2(int kind, double r2, double df) {
    if (kind == 0) return pow(1.0 + r2, -2);
    if (kind == 1) return pow(1.0 + r2 / df, -(df));
    return pow(1.0 + r2 / df, -(df + 1.0));
}

/* nbodyfft.cpp:310-336 -- one point's Lagrange basis values on the in-box nodes */
static void lagrange(int p, const double *nodes, const double *denom, double u, double *out) {
    for (int j = 0; j < p; j++) {
        double v = 1;
        for (int k = 0; k < p; k++) if (k != j) v *= u - nodes[k];
        out[j] = v / denom[j];
    }
}

static void lagrange_setup(int p, double *nodes, double *denom) {
    double h = 1 / (double) p;                 /* nbodyfft.cpp:30-34 */
    nodes[0] = h / 2;
    for (int i = 1; i < p; i++) nodes[i] = nodes[i - 1] + h;
    for (int i = 0; i < p; i++) {              /* nbodyfft.cpp:313-321 */
        denom[i] = 1;
        for (int j = 0; j < p; j++) if (i != j) denom[i] *= nodes[i] - nodes[j];
    }
}

/* ------------------------------------------------------------ 2-D solver -- */
/* potentials[i*T+t] = sum_j K(y_i,y_j) q[j*T+t] through the interpolation grid,
 * restating precompute_2d + n_body_fft_2d for one kernel. */
static void nbody_2d(int N, int T, const double *xs, const double *ys, const double *q, int B, int p,
                     double cmin, double cmax, int kind, double df, double *pot) {
    const int G = p * B;
    const double bw = (cmax - cmin) / (double) B;                    /* nbodyfft.cpp:16 */
    /* n_body_fft_2d re-derives the width from the first box's bounds (nbodyfft.cpp:79-80) */
    const double coord_min = 0 * bw + cmin;
    const double bw2 = (1 * bw + cmin) - coord_min;
    double nodes[64], denom[64];
    lagrange_setup(p, nodes, denom);

    /* global node coordinates, accumulated as the reference does (nbodyfft.cpp:40-46) */
    double *tilde = (double *) malloc(sizeof(double) * (size_t) G);
    double h = (1 / (double) p) * bw;
    tilde[0] = cmin + h / 2;
    for (int i = 1; i < G; i++) tilde[i] = tilde[i - 1] + h;

    int *box = (int *) malloc(sizeof(int) * (size_t) N);
    double *Lx = (double *) malloc(sizeof(double) * (size_t) N * p);
    double *Ly = (double *) malloc(sizeof(double) * (size_t) N * p);
    for (int i = 0; i < N; i++) {
        int xi = (int) ((xs[i] - coord_min) / bw2);                  /* nbodyfft.cpp:86-101 */
        int yi = (int) ((ys[i] - coord_min) / bw2);
        if (xi >= B) xi = B - 1; else if (xi < 0) xi = 0;
        if (yi >= B) yi = B - 1; else if (yi < 0) yi = 0;
        box[i] = yi * B + xi;
        double x_lo = xi * bw + cmin, y_lo = yi * bw + cmin;         /* nbodyfft.cpp:21,24,110-113 */
        lagrange(p, nodes, denom, (xs[i] - x_lo) / bw2, Lx + (size_t) i * p);
        lagrange(p, nodes, denom, (ys[i] - y_lo) / bw2, Ly + (size_t) i * p);
    }

    /* spread: grid row index comes from x, column from y (nbodyfft.cpp:130-147) */
    double *w = (double *) calloc((size_t) G * G * T, sizeof(double));
    for (int i = 0; i < N; i++) {
        int bi = box[i] % B, bj = box[i] / B;
        for (int a = 0; a < p; a++)
            for (int b = 0; b < p; b++) {
                size_t node = (size_t) (bi * p + a) * G + (size_t) bj * p + b;
                double l = Ly[(size_t) i * p + b] * Lx[(size_t) i * p + a];
                for (int t = 0; t < T; t++) w[node * T + t] += l * q[(size_t) i * T + t];
            }
    }

    /* Toeplitz convolution v = K * w with K(di,dj) sampled at node offsets (nbodyfft.cpp:52-61,170-209) */
    const int M = next_pow2(2 * G);
    cplx *tw = make_twiddles(M);
    cplx *kh = (cplx *) calloc((size_t) M * M, sizeof(cplx));
    for (int i = 0; i < G; i++)
        for (int j = 0; j < G; j++) {
            double dy = tilde[0] - tilde[i], dx = tilde[0] - tilde[j];
            double k = kernel_r2(kind, dy * dy + dx * dx, df);
            int ri[2] = { i, (M - i) % M }, cj[2] = { j, (M - j) % M };
            for (int s = 0; s < 2; s++) for (int u = 0; u < 2; u++) kh[(size_t) ri[s] * M + cj[u]].re = k;
        }
    fft2_pow2(kh, M, 0, tw);
    double *v = (double *) malloc(sizeof(double) * (size_t) G * G * T);
    cplx *buf = (cplx *) malloc(sizeof(cplx) * (size_t) M * M);
    for (int t = 0; t < T; t += 2) {           /* two real terms per complex transform */
        memset(buf, 0, sizeof(cplx) * (size_t) M * M);
        for (int i = 0; i < G; i++)
            for (int j = 0; j < G; j++) {
                buf[(size_t) i * M + j].re = w[((size_t) i * G + j) * T + t];
                if (t + 1 < T) buf[(size_t) i * M + j].im = w[((size_t) i * G + j) * T + t + 1];
            }
        fft2_pow2(buf, M, 0, tw);
        for (size_t k = 0; k < (size_t) M * M; k++) {
            cplx a = buf[k], b = kh[k];
            buf[k].re = a.re * b.re - a.im * b.im;
            buf[k].im = a.re * b.im + a.im * b.re;
        }
        fft2_pow2(buf, M, 1, tw);
        double sc = 1.0 / ((double) M * (double) M);
        for (int i = 0; i < G; i++)
            for (int j = 0; j < G; j++) {
                v[((size_t) i * G + j) * T + t] = buf[(size_t) i * M + j].re * sc;
                if (t + 1 < T) v[((size_t) i * G + j) * T + t + 1] = buf[(size_t) i * M + j].im * sc;
            }
    }

    /* gather (nbodyfft.cpp:222-239); pot is zero-initialised by the caller */
    for (int i = 0; i < N; i++) {
        int bi = box[i] % B, bj = box[i] / B;
        for (int a = 0; a < p; a++)
            for (int b = 0; b < p; b++) {
                size_t node = (size_t) (bi * p + a) * G + (size_t) bj * p + b;
                double l = Lx[(size_t) i * p + a] * Ly[(size_t) i * p + b];
                for (int t = 0; t < T; t++) pot[(size_t) i * T + t] += l * v[node * T + t];
            }
    }
    free(tilde); free(box); free(Lx); free(Ly); free(w); free(tw); free(kh); free(v); free(buf);
}

/* ------------------------------------------------------------ 1-D solver -- */
static void nbody_1d(int N, int T, const double *Y, const double *q, int B, int p, double ymin, double ymax,
                     int kind, double df, double *pot) {
    const int G = p * B;
    const double bw = (ymax - ymin) / (double) B;                    /* nbodyfft.cpp:257 */
    const double coord_min = 0 * bw + ymin;                          /* nbodyfft.cpp:344-345 */
    const double bw2 = (1 * bw + ymin) - coord_min;
    double nodes[64], denom[64];
    lagrange_setup(p, nodes, denom);
    double *tilde = (double *) malloc(sizeof(double) * (size_t) G);
    double h = (1 / (double) p) * bw;
    tilde[0] = ymin + h / 2;                                         /* nbodyfft.cpp:276-280 */
    for (int i = 1; i < G; i++) tilde[i] = tilde[i - 1] + h;

    int *box = (int *) malloc(sizeof(int) * (size_t) N);
    double *L = (double *) malloc(sizeof(double) * (size_t) N * p);
    for (int i = 0; i < N; i++) {
        int b = (int) ((Y[i] - coord_min) / bw2);                    /* nbodyfft.cpp:350-355: upper clamp only */
        if (b >= B) b = B - 1;
        box[i] = b;
        lagrange(p, nodes, denom, (Y[i] - (b * bw + ymin)) / bw2, L + (size_t) i * p);
    }
    double *w = (double *) calloc((size_t) G * T, sizeof(double));
    for (int i = 0; i < N; i++)
        for (int a = 0; a < p; a++)
            for (int t = 0; t < T; t++)
                w[(size_t) (box[i] * p + a) * T + t] += L[(size_t) i * p + a] * q[(size_t) i * T + t];

    const int M = next_pow2(2 * G);
    cplx *tw = make_twiddles(M);
    cplx *kh = (cplx *) calloc((size_t) M, sizeof(cplx));
    for (int i = 0; i < G; i++) {                                    /* nbodyfft.cpp:290-298 */
        double d = tilde[0] - tilde[i];
        double k = kernel_r2(kind, d * d, df);
        kh[i].re = k; kh[(M - i) % M].re = k;
    }
    fft_pow2(kh, M, 1, 0, tw);
    double *v = (double *) malloc(sizeof(double) * (size_t) G * T);
    cplx *buf = (cplx *) malloc(sizeof(cplx) * (size_t) M);
    for (int t = 0; t < T; t++) {
        memset(buf, 0, sizeof(cplx) * (size_t) M);
        for (int i = 0; i < G; i++) buf[i].re = w[(size_t) i * T + t];
        fft_pow2(buf, M, 1, 0, tw);
        for (int k = 0; k < M; k++) {
            cplx a = buf[k], b = kh[k];
            buf[k].re = a.re * b.re - a.im * b.im;
            buf[k].im = a.re * b.im + a.im * b.re;
        }
        fft_pow2(buf, M, 1, 1, tw);
        for (int i = 0; i < G; i++) v[(size_t) i * T + t] = buf[i].re / (double) M;
    }
    for (int i = 0; i < N; i++)                                      /* nbodyfft.cpp:439-447 */
        for (int a = 0; a < p; a++)
            for (int t = 0; t < T; t++)
                pot[(size_t) i * T + t] += L[(size_t) i * p + a] * v[(size_t) (box[i] * p + a) * T + t];
    free(tilde); free(box); free(L); free(w); free(tw); free(kh); free(v); free(buf);
}

/* ------------------------------------------------------------- gradient -- */

static int boxes_2d(double span, double ipi, int min_int) {
    /* tsne.cpp:1065-1077 */
    static const int allowed[20] = { 25, 36, 50, 55, 60, 65, 70, 75, 80, 85, 90, 96, 100, 110, 120, 130, 140, 150, 175, 200 };
    int n = (int) fmax(min_int, span / ipi);
    if (n < allowed[19]) {
        int c = 0;
        while (allowed[c] < n) c++;
        n = allowed[c];
    }
    return n;
}

/* Grid sizing exactly as the reference would pick it for this Y; exported so the
 * tests can cross-check the device's choice.  out = {min, max, n_boxes}. */
void fitsne_oracle_grid(int N, int no_dims, const double *Y, double ipi, int min_int, double *out) {
    double mn = INFINITY, mx = -INFINITY;
    if (no_dims == 2) {
        for (int i = 0; i < N; i++) {                                /* tsne.cpp:1042-1049: note the else-if */
            double x = Y[2 * i], y = Y[2 * i + 1];
            if (x > mx) mx = x; else if (x < mn) mn = x;
            if (y > mx) mx = y; else if (y < mn) mn = y;
        }
        out[2] = boxes_2d(mx - mn, ipi, min_int);
    } else {
        for (int i = 0; i < N; i++) {                                /* tsne.cpp:769-772 */
            if (Y[i] < mn) mn = Y[i];
            if (Y[i] > mx) mx = Y[i];
        }
        out[2] = (int) fmax(min_int, (mx - mn) / ipi);               /* tsne.cpp:774 -- no rounding list in 1-D */
    }
    out[0] = mn; out[1] = mx;
}

/* dC = F_attr - F_rep/Z for identical (Y, P); returns 0, writes *sum_Q_out.
 * Only the repulsive part when row_P is all zeros. */
int fitsne_oracle_gradient(int N, int no_dims, const uint32_t *row_P, const uint32_t *col_P, const double *val_P,
                           const double *Y, double *dC, int nterms, double ipi, int min_int, double df,
                           double *sum_Q_out) {
    if (no_dims != 1 && no_dims != 2) return -1;
    if (nterms < 1 || nterms > 64) return -1;
    double g[3];
    fitsne_oracle_grid(N, no_dims, Y, ipi, min_int, g);
    const double mn = g[0], mx = g[1];
    const int B = (int) g[2];
    double sum_Q = 0;
    double *neg = (double *) malloc(sizeof(double) * (size_t) N * no_dims);

    if (no_dims == 2) {
        double *xs = (double *) malloc(sizeof(double) * (size_t) N), *ys = (double *) malloc(sizeof(double) * (size_t) N);
        for (int i = 0; i < N; i++) { xs[i] = Y[2 * i]; ys[i] = Y[2 * i + 1]; }
        if (df == 1.0) {
            const int T = 4;                                         /* tsne.cpp:1052-1062 */
            double *q = (double *) malloc(sizeof(double) * (size_t) N * T), *phi = (double *) calloc((size_t) N * T, sizeof(double));
            for (int i = 0; i < N; i++) {
                q[4 * i] = 1; q[4 * i + 1] = xs[i]; q[4 * i + 2] = ys[i]; q[4 * i + 3] = xs[i] * xs[i] + ys[i] * ys[i];
            }
            nbody_2d(N, T, xs, ys, q, B, nterms, mn, mx, 0, 1.0, phi);
            for (int i = 0; i < N; i++)                              /* tsne.cpp:1101-1110 */
                sum_Q += (1 + xs[i] * xs[i] + ys[i] * ys[i]) * phi[4 * i] - 2 * (xs[i] * phi[4 * i + 1] + ys[i] * phi[4 * i + 2]) + phi[4 * i + 3];
            sum_Q -= N;
            for (int i = 0; i < N; i++) {                            /* tsne.cpp:1149-1151 */
                neg[2 * i] = (xs[i] * phi[4 * i] - phi[4 * i + 1]) / sum_Q;
                neg[2 * i + 1] = (ys[i] * phi[4 * i] - phi[4 * i + 2]) / sum_Q;
            }
            free(q); free(phi);
        } else {
            double *q3 = (double *) malloc(sizeof(double) * (size_t) N * 3), *h3 = (double *) calloc((size_t) N * 3, sizeof(double));
            double *q1 = (double *) malloc(sizeof(double) * (size_t) N), *h1 = (double *) calloc((size_t) N, sizeof(double));
            for (int i = 0; i < N; i++) { q3[3 * i] = xs[i]; q3[3 * i + 1] = ys[i]; q3[3 * i + 2] = 1; q1[i] = 1; }  /* tsne.cpp:906-910,936-938 */
            nbody_2d(N, 3, xs, ys, q3, B, nterms, mn, mx, 2, df, h3);
            nbody_2d(N, 1, xs, ys, q1, B, nterms, mn, mx, 1, df, h1);
            for (int i = 0; i < N; i++) sum_Q += h1[i];              /* tsne.cpp:950-955 */
            sum_Q -= N;
            for (int i = 0; i < N; i++) {                            /* tsne.cpp:986-991 */
                neg[2 * i] = (xs[i] * h3[3 * i + 2] - h3[3 * i]) / sum_Q;
                neg[2 * i + 1] = (ys[i] * h3[3 * i + 2] - h3[3 * i + 1]) / sum_Q;
            }
            free(q3); free(h3); free(q1); free(h1);
        }
        free(xs); free(ys);
    } else {
        if (df == 1.0) {
            const int T = 3;                                         /* tsne.cpp:777-787 */
            double *q = (double *) malloc(sizeof(double) * (size_t) N * T), *phi = (double *) calloc((size_t) N * T, sizeof(double));
            for (int i = 0; i < N; i++) { q[3 * i] = 1; q[3 * i + 1] = Y[i]; q[3 * i + 2] = Y[i] * Y[i]; }
            nbody_1d(N, T, Y, q, B, nterms, mn, mx, 0, 1.0, phi);
            for (int i = 0; i < N; i++) sum_Q += (1 + Y[i] * Y[i]) * phi[3 * i] - 2 * (Y[i] * phi[3 * i + 1]) + phi[3 * i + 2];  /* :809-816 */
            sum_Q -= N;
            for (int i = 0; i < N; i++) neg[i] = (Y[i] * phi[3 * i] - phi[3 * i + 1]) / sum_Q;   /* :847 */
            free(q); free(phi);
        } else {
            double *q2 = (double *) malloc(sizeof(double) * (size_t) N * 2), *h2 = (double *) calloc((size_t) N * 2, sizeof(double));
            double *q1 = (double *) malloc(sizeof(double) * (size_t) N), *h1 = (double *) calloc((size_t) N, sizeof(double));
            for (int i = 0; i < N; i++) { q2[2 * i] = Y[i]; q2[2 * i + 1] = 1; q1[i] = 1; }     /* :670-673,693-695 */
            nbody_1d(N, 2, Y, q2, B, nterms, mn, mx, 2, df, h2);
            nbody_1d(N, 1, Y, q1, B, nterms, mn, mx, 1, df, h1);
            for (int i = 0; i < N; i++) sum_Q += h1[i];              /* :705-710 */
            sum_Q -= N;
            for (int i = 0; i < N; i++) neg[i] = (Y[i] * h2[2 * i + 1] - h2[2 * i]) / sum_Q;    /* :737-739 */
            free(q2); free(h2); free(q1); free(h1);
        }
    }

    /* attractive term over the CSR edges, q_ij = 1/(1+d2/df) */
    for (int i = 0; i < N; i++) {
        double acc[2] = { 0, 0 };
        for (uint32_t e = row_P[i]; e < row_P[i + 1]; e++) {
            uint32_t j = col_P[e];
            double d2 = 0, diff[2];
            for (int d = 0; d < no_dims; d++) { diff[d] = Y[(size_t) i * no_dims + d] - Y[(size_t) j * no_dims + d]; d2 += diff[d] * diff[d]; }
            double qij = 1 / (1 + d2 / df);
            for (int d = 0; d < no_dims; d++) acc[d] += val_P[e] * qij * diff[d];
        }
        for (int d = 0; d < no_dims; d++) dC[(size_t) i * no_dims + d] = acc[d] - neg[(size_t) i * no_dims + d];
    }
    free(neg);
    if (sum_Q_out) *sum_Q_out = sum_Q;
    return 0;
}

/* tsne.cpp:1329-1355, accumulated serially (the reference's threaded version races on C, :1349) */
double fitsne_oracle_kl(int N, int no_dims, const uint32_t *row_P, const uint32_t *col_P, const double *val_P,
                        const double *Y, double sum_Q, double df) {
    double C = 0;
    for (int i = 0; i < N; i++) {
        double temp = 0;
        for (uint32_t e = row_P[i]; e < row_P[i + 1]; e++) {
            uint32_t j = col_P[e];
            double Q = 0;
            for (int d = 0; d < no_dims; d++) { double b = Y[(size_t) i * no_dims + d] - Y[(size_t) j * no_dims + d]; Q += b * b; }
            Q = pow(1.0 / (1.0 + Q / df), df) / sum_Q;
            temp += val_P[e] * log((val_P[e] + FLT_MIN) / (Q + FLT_MIN));
        }
        C += temp;
    }
    return C;
}

static double sgn(double x) { return x == .0 ? .0 : (x < .0 ? -1.0 : 1.0); }   /* tsne.h:37 */

/* One optimiser step + zero-mean (tsne.cpp:479-531).  mode 0: gains+momentum with optional clipping (:492-513);
 * mode 1: gains+momentum without clipping (:481-485); mode 2: plain Y -= dY (:489). */
void fitsne_oracle_step(int N, int no_dims, double *Y, double *uY, double *gains, const double *dY, int mode,
                        double momentum, double learning_rate, double max_step_norm) {
    size_t n = (size_t) N * no_dims;
    if (mode == 2) {
        for (size_t i = 0; i < n; i++) Y[i] = Y[i] - dY[i];
    } else {
        for (size_t i = 0; i < n; i++) {
            gains[i] = (sgn(dY[i]) != sgn(uY[i])) ? (gains[i] + .2) : (gains[i] * .8);
            if (gains[i] < .01) gains[i] = .01;
            uY[i] = momentum * uY[i] - learning_rate * gains[i] * dY[i];
        }
        if (mode == 0 && max_step_norm > 0) {
            for (int i = 0; i < N; i++) {
                double s = 0;
                for (int d = 0; d < no_dims; d++) s += uY[(size_t) i * no_dims + d] * uY[(size_t) i * no_dims + d];
                s = sqrt(s);
                if (s > max_step_norm) for (int d = 0; d < no_dims; d++) uY[(size_t) i * no_dims + d] *= (max_step_norm / s);
            }
        }
        for (size_t i = 0; i < n; i++) Y[i] = Y[i] + uY[i];
    }
    for (int d = 0; d < no_dims; d++) {                              /* zeroMean, tsne.cpp:1851-1876 */
        double m = 0;
        for (int i = 0; i < N; i++) m += Y[(size_t) i * no_dims + d];
        m /= (double) N;
        for (int i = 0; i < N; i++) Y[(size_t) i * no_dims + d] -= m;
    }
}

/* The iteration loop of TSNE::run (tsne.cpp:404-412,437-577) on a caller-supplied P and Y0.
 * val_P is scaled in place like the reference does; costs[] (max_iter, pre-zeroed) is written every 50th
 * iteration and at the last one. */
int fitsne_oracle_run(int N, int no_dims, const uint32_t *row_P, const uint32_t *col_P, double *val_P, double *Y,
                      int max_iter, int stop_lying_iter, int mom_switch_iter, double momentum, double final_momentum,
                      double learning_rate, double early_exag_coeff, double *costs, int no_momentum_during_exag,
                      int start_late_exag_iter, double late_exag_coeff, int nterms, double ipi, int min_int, double df,
                      double max_step_norm) {
    size_t n = (size_t) N * no_dims, E = row_P[N];
    double *dY = (double *) malloc(sizeof(double) * n), *uY = (double *) calloc(n, sizeof(double));
    double *gains = (double *) malloc(sizeof(double) * n);
    for (size_t i = 0; i < n; i++) gains[i] = 1.0;
    if (early_exag_coeff == 0) {                                     /* tsne.cpp:392-402 */
        double mxs = 0;
        for (int r = 0; r < N; r++) {
            double s = 0;
            for (uint32_t e = row_P[r]; e < row_P[r + 1]; e++) s += val_P[e];
            if (s > mxs) mxs = s;
        }
        early_exag_coeff = 1.0 / (learning_rate * mxs);
    }
    for (size_t e = 0; e < E; e++) val_P[e] *= early_exag_coeff;
    double sum_Q = 0;
    for (int iter = 0; iter < max_iter; iter++) {
        int rc = fitsne_oracle_gradient(N, no_dims, row_P, col_P, val_P, Y, dY, nterms, ipi, min_int, df, &sum_Q);
        if (rc) return rc;
        int mode = 0;
        if (no_momentum_during_exag) mode = (iter > stop_lying_iter) ? 1 : 2;
        fitsne_oracle_step(N, no_dims, Y, uY, gains, dY, mode, momentum, learning_rate, max_step_norm);
        if (iter == stop_lying_iter) for (size_t e = 0; e < E; e++) val_P[e] /= early_exag_coeff;
        if (iter == start_late_exag_iter) for (size_t e = 0; e < E; e++) val_P[e] *= late_exag_coeff;
        if (iter == mom_switch_iter) momentum = final_momentum;
        if ((iter + 1) % 50 == 0 || iter == max_iter - 1) {
            double C = fitsne_oracle_kl(N, no_dims, row_P, col_P, val_P, Y, sum_Q, df);
            if (iter < stop_lying_iter && stop_lying_iter != -1) C = C / early_exag_coeff - log(early_exag_coeff);
            if (iter >= start_late_exag_iter && start_late_exag_iter != -1) C = C / late_exag_coeff - log(late_exag_coeff);
            costs[iter] = C;
        }
    }
    free(dY); free(uY); free(gains);
    return 0;
}
