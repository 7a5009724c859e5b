/* TEST INFRASTRUCTURE ONLY -- C restatement ("port") of the reference pyfft algorithm.
 *
 * Not part of the product: libb2fft.so never links or calls this.  It exists so that
 * bench.py can time "the reference's algorithm on the box's host cores" (the reference
 * itself needs Python 2 + Mako + PyCUDA/PyOpenCL and cannot run here) and so that tests can
 * cross-check the numpy restatement (oracle/pyfft_restatement.py) against an independent
 * second implementation.  One OpenMP thread transforms one line at a time.
 *
 * Follows (paths relative to /root/reference):
 *   planner            pyfft/plan.py:111-171          X: local kernel if x <= 2048 (sp) / 1024 (dp),
 *                                                     else global chain; Y, Z: always global chains
 *   radix tables       pyfft/kernel_helpers.py:10-65  (local)   67-122 (global: base radix 128 = R1 x R2)
 *   butterflies        pyfft/kernel.mako:93-214       fftKernel2/4/8/16, natural-order output
 *   local twiddle      pyfft/kernel.mako:566-597      ang = (scalar)(2*dir*pi*k/data_len) * (scalar)m, sincos(ang)
 *   global twiddles    pyfft/kernel.mako:918-930      ang = (scalar)(2*dir*pi*k/radix) * j
 *                      pyfft/kernel.mako:957-971      ang1 = (scalar)(2*dir*pi/curr_n) * l ; ang = ang1 * kk
 *   complex multiply   pyfft/kernel.mako:64
 *   scaling            pyfft/kernel.py:23-37, kernel.mako:271-278   division in the last kernel
 * Twiddles are recomputed with sincos for every element of every line, exactly as the
 * reference's kernels do (that cost is part of the baseline).
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -shared).  The file compiles itself twice,
 * once per precision.
 */
#ifndef PYFFT_PORT_IMPL

#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PI_D 3.14159265358979323846

static int ilog2_(long n) { int l = 0; while ((1L << l) < n) ++l; return l; }

/* kernel_helpers.py:10-65 with max_radix = 0 */
static int local_radix_array(int n, int* out) {
    switch (n) {
        case 2: case 4: case 8: out[0] = n; return 1;
        case 16: case 32: case 64: out[0] = 8; out[1] = n / 8; return 2;
        case 128: out[0] = 8; out[1] = 4; out[2] = 4; return 3;
        case 256: out[0] = 4; out[1] = 4; out[2] = 4; out[3] = 4; return 4;
        case 512: out[0] = 8; out[1] = 8; out[2] = 8; return 3;
        case 1024: out[0] = 16; out[1] = 16; out[2] = 4; return 3;
        case 2048: out[0] = 8; out[1] = 8; out[2] = 8; out[3] = 4; return 4;
    }
    return 0;
}

/* kernel_helpers.py:67-122 */
static int global_radix_info(long n, int* radix, int* r1, int* r2) {
    long base = n < 128 ? n : 128, N = n;
    int num = 0;
    while (N > base) { N /= base; radix[num++] = (int)base; }
    radix[num++] = (int)N;
    for (int i = 0; i < num; ++i) {
        int B = radix[i];
        if (B <= 8) { r1[i] = B; r2[i] = 1; }
        else {
            int a = 2, b = B / a;
            while (b > a) { a *= 2; b = B / a; }
            r1[i] = a; r2[i] = b;
        }
    }
    return num;
}

#define PYFFT_PORT_IMPL
#define REAL float
#define SUF(x) x##_f32
#define SINCOS sincosf
#include "pyfft_port.c"
#undef REAL
#undef SUF
#undef SINCOS
#define REAL double
#define SUF(x) x##_f64
#define SINCOS sincos
#include "pyfft_port.c"
#undef REAL
#undef SUF
#undef SINCOS

int pyfft_port_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

#else /* ------------------------------------------------------------ per-precision body */

typedef struct { REAL x, y; } SUF(cx);

static inline SUF(cx) SUF(cmul)(SUF(cx) a, SUF(cx) b) {      /* kernel.mako:64 */
    SUF(cx) r; r.x = -a.y * b.y + a.x * b.x; r.y = a.y * b.x + a.x * b.y; return r;
}
static inline SUF(cx) SUF(ctm)(SUF(cx) a, REAL d) {          /* kernel.mako:68 */
    SUF(cx) r; r.x = -a.y * d; r.y = a.x * d; return r;
}
static inline void SUF(k2s)(SUF(cx)* p, SUF(cx)* q) {        /* kernel.mako:102-109 */
    SUF(cx) c = *p;
    p->x = c.x + q->x; p->y = c.y + q->y;
    q->x = c.x - q->x; q->y = c.y - q->y;
}
static inline void SUF(swp)(SUF(cx)* p, SUF(cx)* q) { SUF(cx) c = *p; *p = *q; *q = c; }

static void SUF(k4s)(SUF(cx)* a0, SUF(cx)* a1, SUF(cx)* a2, SUF(cx)* a3, int d) {   /* kernel.mako:111-134 */
    SUF(k2s)(a0, a2); SUF(k2s)(a1, a3); SUF(k2s)(a0, a1);
    *a3 = SUF(ctm)(*a3, (REAL)d);
    SUF(k2s)(a2, a3);
    SUF(swp)(a1, a2);
}

static void SUF(k8)(SUF(cx)* a, int d) {                                            /* kernel.mako:136-166 */
    const REAL s = (REAL)0.70710678118654752440;
    SUF(cx) w1 = {s, s * d}, w3 = {-s, s * d};
    for (int i = 0; i < 4; ++i) SUF(k2s)(a + i, a + i + 4);
    a[5] = SUF(cmul)(w1, a[5]);
    a[6] = SUF(ctm)(a[6], (REAL)d);
    a[7] = SUF(cmul)(w3, a[7]);
    SUF(k2s)(a + 0, a + 2); SUF(k2s)(a + 1, a + 3); SUF(k2s)(a + 4, a + 6); SUF(k2s)(a + 5, a + 7);
    a[3] = SUF(ctm)(a[3], (REAL)d);
    a[7] = SUF(ctm)(a[7], (REAL)d);
    SUF(k2s)(a + 0, a + 1); SUF(k2s)(a + 2, a + 3); SUF(k2s)(a + 4, a + 5); SUF(k2s)(a + 6, a + 7);
    SUF(swp)(a + 1, a + 4); SUF(swp)(a + 3, a + 6);
}

static void SUF(k16)(SUF(cx)* a, int d) {                                           /* kernel.mako:168-214 */
    const REAL w0 = (REAL)0.92387953251128675613, w1 = (REAL)0.38268343236508977173,
               w2 = (REAL)0.70710678118654752440;
    SUF(cx) t;
    for (int i = 0; i < 4; ++i) SUF(k4s)(a + i, a + i + 4, a + i + 8, a + i + 12, d);
    t.x = w0; t.y = d * w1; a[5] = SUF(cmul)(a[5], t);
    t.x = w1; t.y = d * w0; a[7] = SUF(cmul)(a[7], t);
    t.x = w2; t.y = d * w2; a[6] = SUF(cmul)(a[6], t); a[9] = SUF(cmul)(a[9], t);
    a[10] = SUF(ctm)(a[10], (REAL)d);
    t.x = -w2; t.y = d * w2; a[11] = SUF(cmul)(a[11], t); a[14] = SUF(cmul)(a[14], t);
    t.x = w1; t.y = d * w0; a[13] = SUF(cmul)(a[13], t);
    t.x = -w0; t.y = -d * w1; a[15] = SUF(cmul)(a[15], t);
    for (int b = 0; b < 16; b += 4) SUF(k4s)(a + b, a + b + 1, a + b + 2, a + b + 3, d);
    SUF(swp)(a + 1, a + 4); SUF(swp)(a + 2, a + 8); SUF(swp)(a + 3, a + 12);
    SUF(swp)(a + 6, a + 9); SUF(swp)(a + 7, a + 13); SUF(swp)(a + 11, a + 14);
}

static void SUF(bfly)(SUF(cx)* a, int R, int d) {
    switch (R) {
        case 2: SUF(k2s)(a, a + 1); break;
        case 4: SUF(k4s)(a, a + 1, a + 2, a + 3, d); break;
        case 8: SUF(k8)(a, d); break;
        case 16: SUF(k16)(a, d); break;
        default: break;   /* radix 1 */
    }
}

/* One localKernel on one line (kernel.mako:725-803).  src/dst: n elements, ping-pong. */
static void SUF(local_line)(SUF(cx)* buf, SUF(cx)* tmp, int n, int d) {
    int radix[8];
    const int nr = local_radix_array(n, radix);
    long P = 1, L = n;
    SUF(cx)* src = buf; SUF(cx)* dst = tmp;
    for (int r = 0; r < nr; ++r) {
        const int R = radix[r];
        const long M = L / R;
        for (long K = 0; K < P; ++K)
            for (long m = 0; m < M; ++m) {
                SUF(cx) a[16];
                for (int j = 0; j < R; ++j) a[j] = src[K * L + m + M * j];
                SUF(bfly)(a, R, d);
                if (r < nr - 1) {
                    const REAL angf = (REAL)m;
                    for (int k = 1; k < R; ++k) {
                        const REAL ang = (REAL)(2.0 * d * PI_D * k / (double)L) * angf;
                        SUF(cx) w; SINCOS(ang, &w.y, &w.x);
                        a[k] = SUF(cmul)(a[k], w);
                    }
                }
                for (int k = 0; k < R; ++k) dst[(K + P * k) * M + m] = a[k];
            }
        P *= R; L = M;
        SUF(cx)* t = src; src = dst; dst = t;
    }
    if (src != buf) memcpy(buf, src, sizeof(SUF(cx)) * (size_t)n);
}

/* The globalKernel chain on one line (kernel.mako:805-1047, kernel.py:259-283). */
static void SUF(global_line)(SUF(cx)* buf, SUF(cx)* tmp, long n, int d) {
    int radix[8], r1[8], r2[8];
    const int np = global_radix_info(n, radix, r1, r2);
    long P = 1, L = n;
    SUF(cx)* src = buf; SUF(cx)* dst = tmp;
    for (int p = 0; p < np; ++p) {
        const int RX = radix[p], R1 = r1[p], R2 = r2[p];
        const long M = L / RX;
        for (long K = 0; K < P; ++K)
            for (long m = 0; m < M; ++m) {
                SUF(cx) a[128], o[128];
                for (int jj = 0; jj < RX; ++jj) a[jj] = src[K * L + m + M * jj];
                if (R2 > 1) {
                    for (int jt = 0; jt < R2; ++jt) {           /* R1-point FFT over elements jt + R2*i */
                        SUF(cx) b[16];
                        for (int i = 0; i < R1; ++i) b[i] = a[jt + R2 * i];
                        SUF(bfly)(b, R1, d);
                        for (int k = 1; k < R1; ++k) {
                            const REAL ang = (REAL)(2.0 * d * PI_D * k / (double)RX) * (REAL)jt;
                            SUF(cx) w; SINCOS(ang, &w.y, &w.x);
                            b[k] = SUF(cmul)(b[k], w);
                        }
                        for (int k = 0; k < R1; ++k) a[jt + R2 * k] = b[k];     /* a[jt + R2*k1] */
                    }
                    for (int k1 = 0; k1 < R1; ++k1) {           /* R2-point FFT over jt */
                        SUF(cx) b[16];
                        for (int t = 0; t < R2; ++t) b[t] = a[t + R2 * k1];
                        SUF(bfly)(b, R2, d);
                        for (int k2 = 0; k2 < R2; ++k2) o[k1 + R1 * k2] = b[k2];
                    }
                } else {
                    SUF(bfly)(a, R1, d);
                    for (int k = 0; k < R1; ++k) o[k] = a[k];
                }
                if (p < np - 1) {
                    const REAL ang1 = (REAL)(2.0 * d * PI_D / (double)L) * (REAL)m;
                    for (int kk = 0; kk < RX; ++kk) {
                        const REAL ang = ang1 * (REAL)kk;
                        SUF(cx) w; SINCOS(ang, &w.y, &w.x);
                        o[kk] = SUF(cmul)(o[kk], w);
                    }
                }
                for (int kk = 0; kk < RX; ++kk) dst[(K + P * kk) * M + m] = o[kk];
            }
        P *= RX; L = M;
        SUF(cx)* t = src; src = dst; dst = t;
    }
    if (src != buf) memcpy(buf, src, sizeof(SUF(cx)) * (size_t)n);
}

/* Restated FFTPlan.execute (plan.py:173-284) for `batch` transforms of (z, y, x).
 * interleaved != 0: in_re/out_re are complex arrays and in_im/out_im are ignored.
 * Returns 0, or -1 for bad sizes. */
int SUF(pyfft_port_execute)(const REAL* in_re, const REAL* in_im, REAL* out_re, REAL* out_im, long x, long y, long z,
                            long batch, int inverse, int normalize, double scale, int interleaved, int nthreads) {
    const long dims[3] = {x, y, z};
    for (int a = 0; a < 3; ++a)
        if (dims[a] < 1 || (dims[a] & (dims[a] - 1))) return -1;
    const int d = inverse ? 1 : -1;
    const long size = x * y * z, total = size * batch;
    const long max_local = sizeof(REAL) == 4 ? 2048 : 1024;       /* plan.py:32,46 */
    SUF(cx)* work = (SUF(cx)*)malloc(sizeof(SUF(cx)) * (size_t)total);
    if (!work) return -2;
    if (interleaved) memcpy(work, in_re, sizeof(SUF(cx)) * (size_t)total);
    else for (long i = 0; i < total; ++i) { work[i].x = in_re[i]; work[i].y = in_im[i]; }
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    const long stride[3] = {1, x, x * y};
    for (int ax = 0; ax < 3; ++ax) {
        const long n = dims[ax];
        if (n <= 1) continue;
        const long st = stride[ax], lines = total / n;
        const int use_local = (ax == 0 && n <= max_local);
#pragma omp parallel
        {
            SUF(cx)* line = (SUF(cx)*)malloc(sizeof(SUF(cx)) * (size_t)n * 2);
            SUF(cx)* tmp = line + n;
#pragma omp for schedule(static)
            for (long ln = 0; ln < lines; ++ln) {
                /* line index -> base offset: lines enumerate (outer, inner) with inner < st */
                const long outer = ln / st, inner = ln % st;
                SUF(cx)* base = work + outer * n * st + inner;
                for (long i = 0; i < n; ++i) line[i] = base[i * st];
                if (use_local) SUF(local_line)(line, tmp, (int)n, d);
                else SUF(global_line)(line, tmp, n, d);
                for (long i = 0; i < n; ++i) base[i * st] = line[i];
            }
            free(line);
        }
    }
    /* kernel.py:23-37 */
    double coeff = inverse ? ((normalize ? (double)size : 1.0) * scale) : (scale == 1.0 ? 1.0 : 1.0 / scale);
    if (coeff != 1.0) {
        const REAL c = (REAL)coeff;
        for (long i = 0; i < total; ++i) { work[i].x = work[i].x / c; work[i].y = work[i].y / c; }
    }
    if (interleaved) memcpy(out_re, work, sizeof(SUF(cx)) * (size_t)total);
    else for (long i = 0; i < total; ++i) { out_re[i] = work[i].x; out_im[i] = work[i].y; }
    free(work);
    return 0;
}

#endif
