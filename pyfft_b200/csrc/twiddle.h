// Host-side twiddle table generation (replaces the reference's on-the-fly sincos,
// pyfft/kernel.mako:35-44,566-597,918-930,957-971).  Tables are computed in long
// double with exact handling of the axis points and rounded once to the working
// precision, i.e. sincospi-class accuracy (<= 0.5 ulp + 1e-19).
#pragma once

#include <cmath>
#include <vector>

#include "fft_core.cuh"

namespace b2 {

// exp(-2*pi*i * e / n), 0 <= e < n, n a power of two.
inline void unit_root(long long e, long long n, long double& c, long double& s) {
    e %= n;
    if (e == 0) { c = 1; s = 0; return; }
    if (n % 4 == 0) {
        const long long q = n / 4;
        if (e == q) { c = 0; s = -1; return; }
        if (e == 2 * q) { c = -1; s = 0; return; }
        if (e == 3 * q) { c = 0; s = 1; return; }
    } else if (n == 2) { c = -1; s = 0; return; }
    const long double pi = 3.14159265358979323846264338327950288L;
    // reduce to the first octant so that cosl/sinl see a small argument
    long long oct = (8 * e) / n;                 // 0..7
    long long r = 8 * e - oct * n;               // remainder in units of 1/(8n) turn ... angle = 2*pi*(oct*n + r)/(8n)
    long double cc, ss;
    if (oct % 2 == 0) {
        long double a = 2.0L * pi * (long double)r / (8.0L * (long double)n);
        cc = cosl(a); ss = sinl(a);
    } else {
        long double a = 2.0L * pi * (long double)(n - r) / (8.0L * (long double)n);
        cc = sinl(a); ss = cosl(a);              // angle within octant = pi/4 - a
    }
    // (cc, ss) = (cos, sin) of the angle inside the even/odd octant pair; rotate by (oct/2) quarter turns
    long double co, so;
    switch ((oct / 2) & 3) {
        case 0: co = cc; so = ss; break;
        case 1: co = -ss; so = cc; break;
        case 2: co = -cc; so = -ss; break;
        default: co = ss; so = -cc; break;
    }
    c = co; s = -so;                              // exp(-i theta)
}

// Stage table for remaining length NS and radix R (M = NS/R):
//   tab[(k-1)*M + m] = exp(-2*pi*i * m*k / NS),  k = 1..R-1, m = 0..M-1
template <typename T>
std::vector<vec2<T>> make_stage_table(int NS, int R) {
    const int M = NS / R;
    std::vector<vec2<T>> tab((size_t)(R - 1) * M);
    for (int k = 1; k < R; ++k)
        for (int m = 0; m < M; ++m) {
            long double c, s;
            unit_root((long long)m * k, NS, c, s);
            tab[(size_t)(k - 1) * M + m].x = (T)c;
            tab[(size_t)(k - 1) * M + m].y = (T)s;
        }
    return tab;
}

// Four-step inter-pass twiddle tables (PassParams::fs_t1 / fs_t2):
//   tab[r*N2 + n2] = exp(-2*pi*i * mult*r*n2 / N),  r = 0..rows-1, n2 = 0..N2-1
// fs_t1: rows = TPC, mult = 1;  fs_t2: rows = E, mult = TPC.
template <typename T>
std::vector<vec2<T>> make_fs_table(long long N, long long N2, int rows, long long mult) {
    std::vector<vec2<T>> tab((size_t)rows * (size_t)N2);
    for (int r = 0; r < rows; ++r)
        for (long long n2 = 0; n2 < N2; ++n2) {
            long double c, s;
            unit_root((mult * r * n2) % N, N, c, s);
            tab[(size_t)r * N2 + n2].x = (T)c;
            tab[(size_t)r * N2 + n2].y = (T)s;
        }
    return tab;
}

}  // namespace b2
