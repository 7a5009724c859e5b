// TEST INFRASTRUCTURE ONLY: runs the exact tile thread program of
// pyfft_b200/csrc/fft_core.cuh on the CPU, one CTA at a time, phase by phase
// (all threads do phase k, then "barrier", then phase k+1), and checks the result
// against a long-double reference FFT.  It validates index arithmetic, twiddle
// tables, butterflies and padding without a GPU.  The product never links this.
//
// Build: g++ -std=c++17 -O1 -I pyfft_b200/csrc tests/host_emu/emu.cpp -o tests/host_emu/emu
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <array>
#include <vector>

#include "fft_core.cuh"
#include "twiddle.h"

using namespace b2;

typedef std::complex<long double> cld;

static void ref_fft(std::vector<cld>& a) {  // in-place recursive radix-2, forward
    const size_t n = a.size();
    if (n == 1) return;
    std::vector<cld> ev(n / 2), od(n / 2);
    for (size_t i = 0; i < n / 2; ++i) { ev[i] = a[2 * i]; od[i] = a[2 * i + 1]; }
    ref_fft(ev);
    ref_fft(od);
    for (size_t k = 0; k < n / 2; ++k) {
        long double c, s;
        unit_root((long long)k, (long long)n, c, s);
        cld t = cld(c, s) * od[k];
        a[k] = ev[k] + t;
        a[k + n / 2] = ev[k] - t;
    }
}

template <class Cfg, bool SPLIT, bool INV, int s, class TH>
static void emu_stages(std::vector<TH>& th, const PassParams<typename Cfg::T>& p,
                       std::vector<vec2<typename Cfg::T>>& smem) {
    for (auto& t : th) t.template compute<s>(p);
    if constexpr (s + 1 < Cfg::S) {
        // poison to catch reads of unwritten slots
        for (auto& v : smem) { v.x = NAN; v.y = NAN; }
        for (auto& t : th) t.template xwrite<s>(smem.data());
        for (auto& t : th) t.template xread<s>(smem.data());
        emu_stages<Cfg, SPLIT, INV, s + 1>(th, p, smem);
    }
}

// Runs one pass over an [outer][N][inner] array and returns the max relative error.
template <class Cfg, bool SPLIT, bool INV>
static double run_case(long long outer, long long inner, unsigned seed, bool staged = false) {
    using T = typename Cfg::T;
    using T2 = vec2<T>;
    const int N = Cfg::N;
    const long long total = outer * N * inner;
    std::mt19937_64 rng(seed);
    std::normal_distribution<double> nd(0.0, 1.0);
    std::vector<T> in_re(total), in_im(total);
    for (long long i = 0; i < total; ++i) { in_re[i] = (T)nd(rng); in_im[i] = (T)nd(rng); }
    std::vector<T2> in_c(total), out_c(total);
    std::vector<T> out_re(total, (T)NAN), out_im(total, (T)NAN);
    for (long long i = 0; i < total; ++i) { in_c[i].x = in_re[i]; in_c[i].y = in_im[i]; out_c[i].x = NAN; out_c[i].y = NAN; }

    PassParams<T> p{};
    std::vector<std::vector<T2>> tabs;
    for (int s = 0; s + 1 < Cfg::S; ++s) tabs.push_back(make_stage_table<T>(Cfg::NS(s), Cfg::R(s)));
    for (int s = 0; s + 1 < Cfg::S; ++s) p.tw[s] = reinterpret_cast<const T*>(tabs[s].data());
    if (SPLIT) { p.in0 = in_re.data(); p.in1 = in_im.data(); p.out0 = out_re.data(); p.out1 = out_im.data(); }
    else { p.in0 = reinterpret_cast<const T*>(in_c.data()); p.out0 = reinterpret_cast<T*>(out_c.data()); }
    p.inner = inner;
    p.inner_blocks = inner / Cfg::W;
    p.outer_stride = (long long)N * inner;
    p.n_tiles = outer * p.inner_blocks;
    p.out_inner = inner;
    p.out_outer_stride = (long long)N * inner;
    p.out_blk_log2 = -1;
    p.scale = (T)0.5;
    p.scale_mode = 1;
    if (inner % Cfg::W != 0) { std::printf("bad inner\n"); std::exit(2); }

    const long long ctas = (p.n_tiles + Cfg::G - 1) / Cfg::G;
    std::vector<T2> smem((size_t)Cfg::COL_SMEM * Cfg::W * Cfg::G + 1);
    for (long long bid = 0; bid < ctas; ++bid) {
        std::vector<TileThread<Cfg, SPLIT, INV>> th(Cfg::THREADS);
        // staged = the persistent TMA kernels' path: the group's input is bulk-copied to "shared
        // memory" (dense, tile g at g*N*W) and stage 0 reads it from there via load_smem()
        std::vector<T2> st_c((size_t)Cfg::G * N * Cfg::W);
        std::vector<T> st_re((size_t)Cfg::G * N * Cfg::W), st_im((size_t)Cfg::G * N * Cfg::W);
        if (staged) {
            long long tiles = p.n_tiles - bid * Cfg::G;
            if (tiles > Cfg::G) tiles = Cfg::G;
            for (long long g = 0; g < tiles; ++g) {                          // dense [g][n][w] staging, as TMA delivers it
                const long long tile = bid * Cfg::G + g, o = tile / p.inner_blocks, ib = tile % p.inner_blocks;
                for (int n = 0; n < N; ++n)
                    for (int w = 0; w < Cfg::W; ++w) {
                        const long long src = (o * N + n) * inner + ib * Cfg::W + w, dst = (g * N + n) * Cfg::W + w;
                        st_c[dst] = in_c[src]; st_re[dst] = in_re[src]; st_im[dst] = in_im[src];
                    }
            }
        }
        for (int tid = 0; tid < Cfg::THREADS; ++tid) {
            th[tid].setup(tid, bid, p);
            if (!staged) th[tid].load(p);
            else if (SPLIT) th[tid].load_smem(st_re.data(), st_im.data());
            else th[tid].load_smem(st_c.data(), nullptr);
        }
        emu_stages<Cfg, SPLIT, INV, 0>(th, p, smem);
        for (auto& t : th) t.store(p);
    }

    // reference
    double max_err = 0, max_mag = 0;
    std::vector<cld> line(N);
    for (long long o = 0; o < outer; ++o)
        for (long long i = 0; i < inner; ++i) {
            for (int n = 0; n < N; ++n) {
                long long idx = (o * N + n) * inner + i;
                line[n] = INV ? cld(in_im[idx], in_re[idx]) : cld(in_re[idx], in_im[idx]);   // inverse via swap identity
            }
            ref_fft(line);
            for (int n = 0; n < N; ++n) {
                long long idx = (o * N + n) * inner + i;
                cld want = INV ? cld(line[n].imag(), line[n].real()) : line[n];
                want *= 0.5L;
                cld got = SPLIT ? cld(out_re[idx], out_im[idx]) : cld(out_c[idx].x, out_c[idx].y);
                double e = (double)std::abs(got - want);
                if (!(e == e)) e = 1e30;   // NaN
                if (e > max_err) max_err = e;
                double m = (double)std::abs(want);
                if (m > max_mag) max_mag = m;
            }
        }
    return max_err / (max_mag > 0 ? max_mag : 1);
}

static int g_fail = 0;

// Destination-blocked stores (slab exchange): the transformed axis is cut into nblk blocks, block h
// is written to its own buffer laid out [outer][N/nblk][inner].  Interleaved forward only.
template <class Cfg>
static void check_blocked(const char* name, long long outer, long long inner, int nblk) {
    using T = typename Cfg::T;
    using T2 = vec2<T>;
    const int N = Cfg::N;
    if (N % nblk) return;
    const int blk = N / nblk;
    const long long total = outer * N * inner;
    std::mt19937_64 rng(77);
    std::normal_distribution<double> nd(0.0, 1.0);
    std::vector<T2> in_c(total);
    for (auto& v : in_c) { v.x = (T)nd(rng); v.y = (T)nd(rng); }
    std::vector<std::vector<T2>> outb(nblk, std::vector<T2>((size_t)outer * blk * inner));
    for (auto& b : outb) for (auto& v : b) { v.x = NAN; v.y = NAN; }
    PassParams<T> p{};
    std::vector<std::vector<T2>> tabs;
    for (int s = 0; s + 1 < Cfg::S; ++s) tabs.push_back(make_stage_table<T>(Cfg::NS(s), Cfg::R(s)));
    for (int s = 0; s + 1 < Cfg::S; ++s) p.tw[s] = reinterpret_cast<const T*>(tabs[s].data());
    p.in0 = reinterpret_cast<const T*>(in_c.data());
    p.inner = inner; p.inner_blocks = inner / Cfg::W; p.outer_stride = (long long)N * inner;
    p.n_tiles = outer * p.inner_blocks;
    p.out_inner = inner; p.out_outer_stride = (long long)blk * inner;
    p.out_blk_log2 = ilog2(blk);
    for (int h = 0; h < nblk; ++h) p.out_blk0[h] = reinterpret_cast<T*>(outb[h].data());
    p.scale = 1; p.scale_mode = 0;
    const long long ctas = (p.n_tiles + Cfg::G - 1) / Cfg::G;
    std::vector<T2> smem((size_t)Cfg::COL_SMEM * Cfg::W * Cfg::G + 1);
    for (long long bid = 0; bid < ctas; ++bid) {
        std::vector<TileThread<Cfg, false, false>> th(Cfg::THREADS);
        for (int tid = 0; tid < Cfg::THREADS; ++tid) { th[tid].setup(tid, bid, p); th[tid].load(p); }
        emu_stages<Cfg, false, false, 0>(th, p, smem);
        for (auto& t : th) t.template store<true>(p);
    }
    double max_err = 0, max_mag = 0;
    std::vector<cld> line(N);
    for (long long o = 0; o < outer; ++o)
        for (long long i = 0; i < inner; ++i) {
            for (int n = 0; n < N; ++n) { auto v = in_c[(o * N + n) * inner + i]; line[n] = cld(v.x, v.y); }
            ref_fft(line);
            for (int n = 0; n < N; ++n) {
                const T2 g = outb[n / blk][((size_t)o * blk + (n % blk)) * inner + i];
                double e = (double)std::abs(cld(g.x, g.y) - line[n]);
                if (!(e == e)) e = 1e30;
                if (e > max_err) max_err = e;
                double m = (double)std::abs(line[n]);
                if (m > max_mag) max_mag = m;
            }
        }
    const double tol = sizeof(T) == 4 ? 3e-6 : 1e-14;
    const bool ok = max_err / max_mag < tol;
    std::printf("%-44s outer=%lld inner=%lld blocked x%d err=%.2e %s\n", name, outer, inner, nblk, max_err / max_mag, ok ? "ok" : "FAIL");
    if (!ok) ++g_fail;
}

// The x-slab exchange pass (pyfft_b200/dist.py _init_xslab): a contiguous-axis pass over the rows
// {all local z} x {y chunk c} of a z-slab [Zl][Y][X], walked through the two-level outer index of
// b2fft_plan_set_outer_split, with destination-blocked stores: x-block h of row (z, y) goes to
// buffer h at [(y*Z + zoff + z)*Xb + xl].  Parameters are set exactly as dist.py / api.cu set them.
template <class Cfg>
static void check_xslab_rows(const char* name) {
    using T = typename Cfg::T;
    using T2 = vec2<T>;
    const int X = Cfg::N, G = 4, Xb = X / G;
    if (Cfg::W != 1 || X % G) return;
    const int Zl = 3, Y = 8, C = 2, Yc = Y / C, Z = Zl * G, zoff = Zl;      // "rank 1" of 4: z offset Zl
    std::mt19937_64 rng(99);
    std::normal_distribution<double> nd(0.0, 1.0);
    std::vector<T2> slab((size_t)Zl * Y * X);
    for (auto& v : slab) { v.x = (T)nd(rng); v.y = (T)nd(rng); }
    std::vector<std::vector<T2>> xs(G, std::vector<T2>((size_t)Y * Z * Xb));
    for (auto& b : xs) for (auto& v : b) { v.x = NAN; v.y = NAN; }
    std::vector<std::vector<T2>> tabs;
    for (int s = 0; s + 1 < Cfg::S; ++s) tabs.push_back(make_stage_table<T>(Cfg::NS(s), Cfg::R(s)));
    for (int c = 0; c < C; ++c) {
        PassParams<T> p{};
        for (int s = 0; s + 1 < Cfg::S; ++s) p.tw[s] = reinterpret_cast<const T*>(tabs[s].data());
        p.in0 = reinterpret_cast<const T*>(slab.data() + (size_t)c * Yc * X);
        p.inner = 1; p.inner_blocks = 1;
        p.n_tiles = (long long)Yc * Zl;
        p.out_inner = 1;
        p.out_blk_log2 = ilog2(Xb);
        for (int h = 0; h < G; ++h) p.out_blk0[h] = reinterpret_cast<T*>(xs[h].data() + ((size_t)c * Yc * Z + zoff) * Xb);
        p.outer_div = Yc; p.outer_stride = X; p.in_stride_hi = (long long)Y * X;
        p.out_outer_stride = (long long)Z * Xb; p.out_stride_hi = Xb;
        p.scale = 1; p.scale_mode = 0;
        const long long ctas = (p.n_tiles + Cfg::G - 1) / Cfg::G;
        std::vector<T2> smem((size_t)Cfg::COL_SMEM * Cfg::W * Cfg::G + 1);
        for (long long bid = 0; bid < ctas; ++bid) {
            std::vector<TileThread<Cfg, false, false>> th(Cfg::THREADS);
            for (int tid = 0; tid < Cfg::THREADS; ++tid) { th[tid].setup(tid, bid, p); th[tid].load(p); }
            emu_stages<Cfg, false, false, 0>(th, p, smem);
            for (auto& t : th) t.template store<true>(p);
        }
    }
    double max_err = 0, max_mag = 0;
    std::vector<cld> line(X);
    for (int z = 0; z < Zl; ++z)
        for (int y = 0; y < Y; ++y) {
            for (int x = 0; x < X; ++x) { auto v = slab[((size_t)z * Y + y) * X + x]; line[x] = cld(v.x, v.y); }
            ref_fft(line);
            for (int x = 0; x < X; ++x) {
                const T2 g = xs[x / Xb][((size_t)y * Z + zoff + z) * Xb + x % Xb];
                double e = (double)std::abs(cld(g.x, g.y) - line[x]);
                if (!(e == e)) e = 1e30;
                if (e > max_err) max_err = e;
                double m = (double)std::abs(line[x]);
                if (m > max_mag) max_mag = m;
            }
        }
    // nothing outside this rank's z range may have been written
    bool clean = true;
    for (int h = 0; h < G; ++h)
        for (int y = 0; y < Y; ++y)
            for (int z = 0; z < Z; ++z)
                if (z < zoff || z >= zoff + Zl)
                    for (int x = 0; x < Xb; ++x) clean = clean && std::isnan((double)xs[h][((size_t)y * Z + z) * Xb + x].x);
    const double tol = sizeof(T) == 4 ? 3e-6 : 1e-14;
    const bool ok = max_err / max_mag < tol && clean;
    std::printf("%-44s x-slab exchange rows (outer split, 4 blocks) err=%.2e %s\n", name, max_err / max_mag, ok ? "ok" : "FAIL");
    if (!ok) ++g_fail;
}

// The inverse x-slab exchange pass (csrc/slab.cu b2fft_slab_inverse): a contiguous-axis pass whose LOADS are
// source-blocked -- piece h of row (z, y) comes from buffer h at [(y*Z + zoff + z)*Xb + xl] -- walked through the
// two-level outer index, with plain stores into the z-slab [Zl][Y][X].  Parameters as api.cu / slab.cu set them.
template <class Cfg>
static void check_xslab_pull_rows(const char* name) {
    using T = typename Cfg::T;
    using T2 = vec2<T>;
    const int X = Cfg::N, G = 4, Xb = X / G;
    if (Cfg::W != 1 || X % G) return;
    const int Zl = 3, Y = 8, C = 2, Yc = Y / C, Z = Zl * G, zoff = 2 * Zl;      // "rank 2" of 4
    std::mt19937_64 rng(123);
    std::normal_distribution<double> nd(0.0, 1.0);
    std::vector<std::vector<T2>> xs(G, std::vector<T2>((size_t)Y * Z * Xb));
    for (auto& b : xs) for (auto& v : b) { v.x = (T)nd(rng); v.y = (T)nd(rng); }
    std::vector<T2> slab((size_t)Zl * Y * X);
    for (auto& v : slab) { v.x = NAN; v.y = NAN; }
    std::vector<std::vector<T2>> tabs;
    for (int s = 0; s + 1 < Cfg::S; ++s) tabs.push_back(make_stage_table<T>(Cfg::NS(s), Cfg::R(s)));
    for (int c = 0; c < C; ++c) {
        PassParams<T> p{};
        for (int s = 0; s + 1 < Cfg::S; ++s) p.tw[s] = reinterpret_cast<const T*>(tabs[s].data());
        p.out0 = reinterpret_cast<T*>(slab.data() + (size_t)c * Yc * X);
        p.inner = 1; p.inner_blocks = 1;
        p.n_tiles = (long long)Yc * Zl;
        p.out_inner = 1;
        p.out_blk_log2 = -1;
        p.in_blk_log2 = ilog2(Xb);
        for (int h = 0; h < G; ++h) p.in_blk[h] = reinterpret_cast<const T*>(xs[h].data() + ((size_t)c * Yc * Z + zoff) * Xb);
        p.outer_div = Yc; p.outer_stride = (long long)Z * Xb; p.in_stride_hi = Xb;
        p.out_outer_stride = X; p.out_stride_hi = (long long)Y * X;
        p.scale = 1; p.scale_mode = 0;
        const long long ctas = (p.n_tiles + Cfg::G - 1) / Cfg::G;
        std::vector<T2> smem((size_t)Cfg::COL_SMEM * Cfg::W * Cfg::G + 1);
        for (long long bid = 0; bid < ctas; ++bid) {
            std::vector<TileThread<Cfg, false, true>> th(Cfg::THREADS);
            for (int tid = 0; tid < Cfg::THREADS; ++tid) { th[tid].setup(tid, bid, p); th[tid].load(p); }
            emu_stages<Cfg, false, true, 0>(th, p, smem);
            for (auto& t : th) t.store(p);
        }
    }
    double max_err = 0, max_mag = 0;
    std::vector<cld> line(X);
    for (int z = 0; z < Zl; ++z)
        for (int y = 0; y < Y; ++y) {
            for (int x = 0; x < X; ++x) { auto v = xs[x / Xb][((size_t)y * Z + zoff + z) * Xb + x % Xb]; line[x] = cld(v.y, v.x); }   // inverse via swap
            ref_fft(line);
            for (int x = 0; x < X; ++x) {
                const T2 g = slab[((size_t)z * Y + y) * X + x];
                double e = (double)std::abs(cld(g.x, g.y) - cld(line[x].imag(), line[x].real()));
                if (!(e == e)) e = 1e30;
                if (e > max_err) max_err = e;
                double m = (double)std::abs(line[x]);
                if (m > max_mag) max_mag = m;
            }
        }
    const double tol = sizeof(T) == 4 ? 3e-6 : 1e-14;
    const bool ok = max_err / max_mag < tol;
    std::printf("%-44s x-slab inverse rows (source-blocked loads, outer split) err=%.2e %s\n", name, max_err / max_mag, ok ? "ok" : "FAIL");
    if (!ok) ++g_fail;
}

template <class Cfg>
static void check(const char* name, long long outer, long long inner) {
    using T = typename Cfg::T;
    const double tol = sizeof(T) == 4 ? 3e-6 : 1e-14;
    double e0 = run_case<Cfg, false, false>(outer, inner, 1);
    double e1 = run_case<Cfg, false, true>(outer, inner, 2);
    double e2 = run_case<Cfg, true, false>(outer, inner, 3);
    bool ok = e0 < tol && e1 < tol && e2 < tol;
    std::printf("%-44s outer=%lld inner=%lld  err fwd=%.2e inv=%.2e split=%.2e  smem=%lld thr=%d %s\n", name, outer, inner,
                e0, e1, e2, (long long)Cfg::SMEM_BYTES, Cfg::THREADS, ok ? "ok" : "FAIL");
    if (!ok) ++g_fail;
}

template <class Cfg>
static void check_staged(const char* name, long long outer) {
    using T = typename Cfg::T;
    const double tol = sizeof(T) == 4 ? 3e-6 : 1e-14;
    const long long inner = Cfg::W == 1 ? 1 : 3 * Cfg::W;
    double e0 = run_case<Cfg, false, false>(outer, inner, 4, true);
    double e1 = run_case<Cfg, false, true>(outer, inner, 5, true);
    double e2 = run_case<Cfg, true, false>(outer, inner, 6, true);
    bool ok = e0 < tol && e1 < tol && e2 < tol;
    std::printf("%-44s outer=%lld staged   err fwd=%.2e inv=%.2e split=%.2e %s\n", name, outer, e0, e1, e2, ok ? "ok" : "FAIL");
    if (!ok) ++g_fail;
}

// The long-row kernel (kernels.cuh tile_fft_kernel_tma_row_alias): ONE buffer is staging slot and exchange buffer.  One CTA
// walks the rows in the kernel's order of events: stage-0 input is read from the buffer; every exchange poisons and rewrites
// it; right after the LAST exchange has been read the buffer is poisoned and refilled with the next group of rows (what the
// bulk copy does once every thread has arrived on the "empty" barrier), and only then the last radix stage and the stores
// of the current rows run -- so a stage that still needed the buffer after the hand-back, or a refill landing in the wrong
// place, shows up as a wrong result.
template <class Cfg, bool SPLIT, bool INV, int s, class TH, class F>
static void emu_stages_hook(std::vector<TH>& th, const PassParams<typename Cfg::T>& p, vec2<typename Cfg::T>* smem, size_t smem_elems,
                            F&& after_last_read) {
    for (auto& t : th) t.template compute<s>(p);
    if constexpr (s + 1 < Cfg::S) {
        for (size_t i = 0; i < smem_elems; ++i) { smem[i].x = NAN; smem[i].y = NAN; }
        for (auto& t : th) t.template xwrite<s>(smem);
        for (auto& t : th) t.template xread<s>(smem);
        if constexpr (s + 2 == Cfg::S) after_last_read();
        emu_stages_hook<Cfg, SPLIT, INV, s + 1>(th, p, smem, smem_elems, static_cast<F&&>(after_last_read));
    }
}

template <class Cfg, bool SPLIT, bool INV>
static double run_alias_row(long long rows, unsigned seed) {
    using T = typename Cfg::T;
    using T2 = vec2<T>;
    constexpr int N = Cfg::N, G = Cfg::G;
    static_assert(Cfg::W == 1 && Cfg::S >= 2, "long contiguous rows");
    const long long total = rows * N;
    std::mt19937_64 rng(seed);
    std::normal_distribution<double> nd(0.0, 1.0);
    std::vector<T> in_re(total), in_im(total), out_re(total, (T)NAN), out_im(total, (T)NAN);
    for (long long i = 0; i < total; ++i) { in_re[i] = (T)nd(rng); in_im[i] = (T)nd(rng); }
    std::vector<T2> in_c(total), out_c(total);
    for (long long i = 0; i < total; ++i) { in_c[i].x = in_re[i]; in_c[i].y = in_im[i]; out_c[i].x = NAN; out_c[i].y = NAN; }
    PassParams<T> p{};
    std::vector<std::vector<T2>> tabs;
    for (int s = 0; s + 1 < Cfg::S; ++s) tabs.push_back(make_stage_table<T>(Cfg::NS(s), Cfg::R(s)));
    for (int s = 0; s + 1 < Cfg::S; ++s) p.tw[s] = reinterpret_cast<const T*>(tabs[s].data());
    if (SPLIT) { p.in0 = in_re.data(); p.in1 = in_im.data(); p.out0 = out_re.data(); p.out1 = out_im.data(); }
    else { p.in0 = reinterpret_cast<const T*>(in_c.data()); p.out0 = reinterpret_cast<T*>(out_c.data()); }
    p.inner = 1; p.inner_blocks = 1; p.outer_stride = N; p.n_tiles = rows; p.out_inner = 1; p.out_outer_stride = N;
    p.out_blk_log2 = -1; p.in_blk_log2 = -1; p.scale = (T)0.5; p.scale_mode = 1;
    // the buffer: max(staging, exchange) bytes, as TmaRowAliasLayout sizes it
    const size_t in_bytes = (size_t)G * N * 2 * sizeof(T), x_bytes = (size_t)Cfg::COL_SMEM * G * 2 * sizeof(T);
    std::vector<T2> buf((in_bytes > x_bytes ? in_bytes : x_bytes) / sizeof(T2) + 1);
    unsigned char* raw = reinterpret_cast<unsigned char*>(buf.data());
    auto stage = [&](long long grp) {
        for (auto& v : buf) { v.x = NAN; v.y = NAN; }
        long long tiles = rows - grp * G;
        if (tiles > G) tiles = G;
        if (tiles <= 0) return;
        if (SPLIT) {
            std::memcpy(raw, in_re.data() + grp * G * N, (size_t)tiles * N * sizeof(T));
            std::memcpy(raw + in_bytes / 2, in_im.data() + grp * G * N, (size_t)tiles * N * sizeof(T));
        } else {
            std::memcpy(raw, in_c.data() + grp * G * N, (size_t)tiles * N * sizeof(T2));
        }
    };
    const long long n_groups = (rows + G - 1) / G;
    stage(0);
    for (long long grp = 0; grp < n_groups; ++grp) {
        std::vector<TileThread<Cfg, SPLIT, INV>> th(Cfg::THREADS);
        for (int tid = 0; tid < Cfg::THREADS; ++tid) {
            th[tid].setup(tid, grp, p);
            th[tid].load_smem(raw, raw + in_bytes / 2);
        }
        emu_stages_hook<Cfg, SPLIT, INV, 0>(th, p, buf.data(), buf.size(), [&] { stage(grp + 1); });
        for (auto& t : th) t.store(p);
    }
    double max_err = 0, max_mag = 0;
    std::vector<cld> line(N);
    for (long long r = 0; r < rows; ++r) {
        for (int n = 0; n < N; ++n) line[n] = INV ? cld(in_im[r * N + n], in_re[r * N + n]) : cld(in_re[r * N + n], in_im[r * N + n]);
        ref_fft(line);
        for (int n = 0; n < N; ++n) {
            cld want = INV ? cld(line[n].imag(), line[n].real()) : line[n];
            want *= 0.5L;
            cld got = SPLIT ? cld(out_re[r * N + n], out_im[r * N + n]) : cld(out_c[r * N + n].x, out_c[r * N + n].y);
            double e = (double)std::abs(got - want);
            if (!(e == e)) e = 1e30;
            max_err = std::max(max_err, e);
            max_mag = std::max(max_mag, (double)std::abs(want));
        }
    }
    return max_err / (max_mag > 0 ? max_mag : 1);
}

template <class Cfg>
static void check_alias_row(const char* name) {
    using T = typename Cfg::T;
    const double tol = sizeof(T) == 4 ? 3e-6 : 1e-14;
    double e0 = run_alias_row<Cfg, false, false>(2 * Cfg::G + 1, 71), e1 = run_alias_row<Cfg, false, true>(Cfg::G + 1, 72),
           e2 = run_alias_row<Cfg, true, false>(2 * Cfg::G + 1, 73);
    bool ok = e0 < tol && e1 < tol && e2 < tol;
    std::printf("%-44s rows, staging slot = exchange buffer  err fwd=%.2e inv=%.2e split=%.2e %s\n", name, e0, e1, e2, ok ? "ok" : "FAIL");
    if (!ok) ++g_fail;
}

// One pass of the thread program over the arrays `p` points at (plain or four-step "A" mode).
template <class Cfg, bool SPLIT, bool INV, bool FS>
static void emu_pass(const PassParams<typename Cfg::T>& p) {
    using T2 = vec2<typename Cfg::T>;
    const long long ctas = (p.n_tiles + Cfg::G - 1) / Cfg::G;
    std::vector<T2> smem((size_t)Cfg::COL_SMEM * Cfg::W * Cfg::G + 1);
    for (long long bid = 0; bid < ctas; ++bid) {
        std::vector<TileThread<Cfg, SPLIT, INV, FS>> th(Cfg::THREADS);
        for (int tid = 0; tid < Cfg::THREADS; ++tid) { th[tid].setup(tid, bid, p); th[tid].load(p); }
        emu_stages<Cfg, SPLIT, INV, 0>(th, p, smem);
        for (auto& t : th) t.store(p);
    }
}

// Four-step decomposition of an axis of length N = CfgA::N * CfgB::N with element stride inner0:
// transposing pass A (CfgA, FS mode) followed by the plain strided pass B (CfgB, stride N1*inner0),
// exactly as api.cu's add_axis()/launch_pass() set them up, against a direct length-N FFT.
template <class CfgA, class CfgB, bool SPLIT, bool INV>
static double run_fourstep(long long outer, long long inner0, unsigned seed) {
    using T = typename CfgA::T;
    using T2 = vec2<T>;
    const long long N1 = CfgA::N, N2 = CfgB::N, N = N1 * N2, total = outer * N * inner0;
    std::mt19937_64 rng(seed);
    std::normal_distribution<double> nd(0.0, 1.0);
    std::vector<T> a_re(total), a_im(total), b_re(total, (T)NAN), b_im(total, (T)NAN);
    for (long long i = 0; i < total; ++i) { a_re[i] = (T)nd(rng); a_im[i] = (T)nd(rng); }
    std::vector<T2> a_c(total), b_c(total);
    for (long long i = 0; i < total; ++i) { a_c[i].x = a_re[i]; a_c[i].y = a_im[i]; b_c[i].x = NAN; b_c[i].y = NAN; }
    auto set_io = [&](PassParams<T>& p, bool from_a) {
        if (SPLIT) {
            p.in0 = from_a ? a_re.data() : b_re.data(); p.in1 = from_a ? a_im.data() : b_im.data();
            p.out0 = b_re.data(); p.out1 = b_im.data();
        } else {
            p.in0 = reinterpret_cast<const T*>(from_a ? a_c.data() : b_c.data());
            p.out0 = reinterpret_cast<T*>(b_c.data());
        }
    };
    // pass A: [outer][N1][N2*inner0] -> [outer][N2][N1][inner0]
    {
        PassParams<T> p{};
        std::vector<std::vector<T2>> tabs;
        for (int s = 0; s + 1 < CfgA::S; ++s) tabs.push_back(make_stage_table<T>(CfgA::NS(s), CfgA::R(s)));
        for (int s = 0; s + 1 < CfgA::S; ++s) p.tw[s] = reinterpret_cast<const T*>(tabs[s].data());
        auto t1 = make_fs_table<T>(N, N2, CfgA::TPC, 1), t2 = make_fs_table<T>(N, N2, CfgA::E, CfgA::TPC);
        set_io(p, true);
        p.inner = N2 * inner0;
        if (p.inner % CfgA::W) { std::printf("bad inner\n"); std::exit(2); }
        p.inner_blocks = p.inner / CfgA::W;
        p.outer_stride = N * inner0;
        p.n_tiles = outer * p.inner_blocks;
        p.out_inner = inner0;
        p.out_outer_stride = p.outer_stride;
        p.out_blk_log2 = -1;
        p.scale = 1; p.scale_mode = 0;
        p.fs_log2_inner = ilog2((int)inner0);
        p.fs_n2 = N2;
        p.fs_col_stride = N1 * inner0;
        p.fs_t1 = reinterpret_cast<const T*>(t1.data());
        p.fs_t2 = reinterpret_cast<const T*>(t2.data());
        emu_pass<CfgA, SPLIT, INV, true>(p);
    }
    // pass B: [outer][N2][N1*inner0], in place
    {
        PassParams<T> p{};
        std::vector<std::vector<T2>> tabs;
        for (int s = 0; s + 1 < CfgB::S; ++s) tabs.push_back(make_stage_table<T>(CfgB::NS(s), CfgB::R(s)));
        for (int s = 0; s + 1 < CfgB::S; ++s) p.tw[s] = reinterpret_cast<const T*>(tabs[s].data());
        set_io(p, false);
        p.inner = N1 * inner0;
        if (p.inner % CfgB::W) { std::printf("bad inner\n"); std::exit(2); }
        p.inner_blocks = p.inner / CfgB::W;
        p.outer_stride = N * inner0;
        p.n_tiles = outer * p.inner_blocks;
        p.out_inner = p.inner;
        p.out_outer_stride = p.outer_stride;
        p.out_blk_log2 = -1;
        p.scale = (T)0.5; p.scale_mode = 1;
        emu_pass<CfgB, SPLIT, INV, false>(p);
    }
    double max_err = 0, max_mag = 0;
    std::vector<cld> line(N);
    for (long long o = 0; o < outer; ++o)
        for (long long i = 0; i < inner0; ++i) {
            for (long long n = 0; n < N; ++n) {
                const long long idx = (o * N + n) * inner0 + i;
                line[n] = INV ? cld(a_im[idx], a_re[idx]) : cld(a_re[idx], a_im[idx]);
            }
            ref_fft(line);
            for (long long n = 0; n < N; ++n) {
                const long long idx = (o * N + n) * inner0 + i;
                cld want = INV ? cld(line[n].imag(), line[n].real()) : line[n];
                want *= 0.5L;
                cld got = SPLIT ? cld(b_re[idx], b_im[idx]) : cld(b_c[idx].x, b_c[idx].y);
                double e = (double)std::abs(got - want);
                if (!(e == e)) e = 1e30;
                if (e > max_err) max_err = e;
                double m = (double)std::abs(want);
                if (m > max_mag) max_mag = m;
            }
        }
    return max_err / (max_mag > 0 ? max_mag : 1);
}

template <class CfgA>
static void check_fourstep(const char* name) {
    using T = typename CfgA::T;
    using CfgB = TileCfg<T, 4, (sizeof(T) == 4 ? 16 : 8), (sizeof(T) == 4 ? 8 : 16), 16, 1, 1, 1>;
    const double tol = sizeof(T) == 4 ? 4e-6 : 2e-14;
    double e0 = run_fourstep<CfgA, CfgB, false, false>(1, 1, 11);
    double e1 = run_fourstep<CfgA, CfgB, false, true>(1, 2, 12);
    double e2 = run_fourstep<CfgA, CfgB, true, false>(2, 2 * CfgA::W, 13);
    bool ok = e0 < tol && e1 < tol && e2 < tol;
    std::printf("%-44s four-step N=%dx16  err fwd(i1)=%.2e inv(i2)=%.2e split(i2W)=%.2e %s\n", name, CfgA::N, e0, e1, e2,
                ok ? "ok" : "FAIL");
    if (!ok) ++g_fail;
}

// The fused two-step strided kernel (kernels.cuh fused2_fft_kernel): per super-tile, step A sub-tiles into the
// CTA's scratch slot, then step B sub-tiles out of it, with the parameters VariantOpsFused2::launch builds.
// `grid` CTAs stride over the super-tiles (slot reuse is exercised); in_place: out == in.
template <class CfgA, class CfgB, bool INV>
static double run_fused2(long long outer, long long inner, int grid, bool in_place, unsigned seed) {
    using T = typename CfgA::T;
    using T2 = vec2<T>;
    constexpr int W = CfgA::W, N1 = CfgA::N, N2 = CfgB::N, N = N1 * N2;
    const long long total = outer * N * inner;
    std::mt19937_64 rng(seed);
    std::normal_distribution<double> nd(0.0, 1.0);
    std::vector<T2> in_c(total), out_c(total), ref_in(total);
    for (long long i = 0; i < total; ++i) { in_c[i].x = (T)nd(rng); in_c[i].y = (T)nd(rng); out_c[i].x = NAN; out_c[i].y = NAN; }
    ref_in = in_c;
    std::vector<T2> scratch((size_t)grid * N * W);
    for (auto& v : scratch) { v.x = NAN; v.y = NAN; }
    std::vector<std::vector<T2>> ta, tb;
    for (int s = 0; s + 1 < CfgA::S; ++s) ta.push_back(make_stage_table<T>(CfgA::NS(s), CfgA::R(s)));
    for (int s = 0; s + 1 < CfgB::S; ++s) tb.push_back(make_stage_table<T>(CfgB::NS(s), CfgB::R(s)));
    auto t1 = make_fs_table<T>(N, N2, CfgA::TPC, 1), t2 = make_fs_table<T>(N, N2, CfgA::E, CfgA::TPC);
    PassParams<T> p{};
    p.in0 = reinterpret_cast<const T*>(in_c.data());
    p.out0 = reinterpret_cast<T*>(in_place ? in_c.data() : out_c.data());
    p.inner = inner; p.inner_blocks = inner / W; p.outer_stride = (long long)N * inner;
    p.n_tiles = outer * p.inner_blocks;
    p.out_inner = inner; p.out_outer_stride = p.outer_stride; p.out_blk_log2 = -1;
    p.scale = (T)0.5; p.scale_mode = 1;
    p.fs_t1 = reinterpret_cast<const T*>(t1.data()); p.fs_t2 = reinterpret_cast<const T*>(t2.data());
    for (int s = 0; s + 1 < CfgA::S; ++s) p.tw[s] = reinterpret_cast<const T*>(ta[s].data());
    for (int s = 0; s + 1 < CfgB::S; ++s) p.tw_b[s] = reinterpret_cast<const T*>(tb[s].data());
    p.scratch = reinterpret_cast<T*>(scratch.data());
    // ---- as VariantOpsFused2::launch
    PassParams<T> pa = p, pb = p;
    pa.inner = (long long)N2 * p.inner; pa.out0 = p.scratch; pa.out_inner = W; pa.scale_mode = 0; pa.fs_n2 = N2;
    pb.in0 = p.scratch; pb.inner = (long long)N1 * W; pb.out_inner = (long long)N1 * p.out_inner;
    for (int s = 0; s < 3; ++s) pb.tw[s] = p.tw_b[s];
    pb.fs_t1 = pb.fs_t2 = nullptr;
    std::vector<T2> smem((size_t)(CfgA::COL_SMEM * CfgA::G > CfgB::COL_SMEM * CfgB::G ? CfgA::COL_SMEM * CfgA::G : CfgB::COL_SMEM * CfgB::G) * W + 1);
    for (int bid = 0; bid < grid; ++bid) {
        const long long slot = (long long)bid * N * W;
        for (long long sidx = bid; sidx < pa.n_tiles; sidx += grid) {
            const long long o = sidx / pa.inner_blocks, ib = sidx - o * pa.inner_blocks;
            for (int c = 0; c < N2 / CfgA::G; ++c) {
                std::vector<TileThread<CfgA, false, INV, true>> th(CfgA::THREADS);
                for (int tid = 0; tid < CfgA::THREADS; ++tid) {
                    fused2_setup_a<CfgA, CfgB>(th[tid], tid, c, o * pa.outer_stride + ib * W, p.inner, slot);
                    th[tid].load(pa);
                }
                emu_stages<CfgA, false, INV, 0>(th, pa, smem);
                for (auto& t : th) t.store(pa);
            }
            for (int c = 0; c < N1 / CfgB::G; ++c) {
                std::vector<TileThread<CfgB, false, INV, false>> th(CfgB::THREADS);
                for (int tid = 0; tid < CfgB::THREADS; ++tid) {
                    fused2_setup_b<CfgA, CfgB>(th[tid], tid, c, o * pb.out_outer_stride + ib * W, p.out_inner, slot);
                    th[tid].load(pb);
                }
                emu_stages<CfgB, false, INV, 0>(th, pb, smem);
                for (auto& t : th) t.store(pb);
            }
            for (long long i = 0; i < (long long)N * W; ++i) { scratch[slot + i].x = NAN; scratch[slot + i].y = NAN; }   // poison
        }
    }
    const std::vector<T2>& got_c = in_place ? in_c : out_c;
    double max_err = 0, max_mag = 0;
    std::vector<cld> line(N);
    for (long long o = 0; o < outer; ++o)
        for (long long i = 0; i < inner; ++i) {
            for (long long n = 0; n < N; ++n) {
                const T2 v = ref_in[(o * N + n) * inner + i];
                line[n] = INV ? cld(v.y, v.x) : cld(v.x, v.y);
            }
            ref_fft(line);
            for (long long n = 0; n < N; ++n) {
                cld want = INV ? cld(line[n].imag(), line[n].real()) : line[n];
                want *= 0.5L;
                const T2 g = got_c[(o * N + n) * inner + i];
                double e = (double)std::abs(cld(g.x, g.y) - want);
                if (!(e == e)) e = 1e30;
                if (e > max_err) max_err = e;
                double m = (double)std::abs(want);
                if (m > max_mag) max_mag = m;
            }
        }
    return max_err / (max_mag > 0 ? max_mag : 1);
}

template <class CfgA, class CfgB>
static void check_fused2(const char* name) {
    using T = typename CfgA::T;
    const double tol = sizeof(T) == 4 ? 4e-6 : 2e-14;
    static_assert(CfgA::THREADS == CfgB::THREADS && CfgA::W == CfgB::W, "fused steps share the CTA shape");
    double e0 = run_fused2<CfgA, CfgB, false>(1, 3 * CfgA::W, 2, false, 21);
    double e1 = run_fused2<CfgA, CfgB, true>(2, CfgA::W, 1, true, 22);
    bool ok = e0 < tol && e1 < tol;
    std::printf("%-44s fused two-step N=%dx%d  err fwd(oop)=%.2e inv(in place)=%.2e thr=%d %s\n", name, CfgA::N, CfgB::N, e0, e1,
                CfgA::THREADS, ok ? "ok" : "FAIL");
    if (!ok) ++g_fail;
}

// The shared-memory-resident fused two-step kernel (kernels.cuh fused2s_fft_kernel): step A threads store through
// fused2s_store_a into the [k1][n2][W] intermediate (rows k1 < KS in "shared memory", the rest in the scratch slot), step B
// threads read their column back with fused2s_load_b; parameters as VariantOpsFused2S::launch builds them.
template <class CfgA, class CfgB, int KS, bool INV>
static double run_fused2s(long long outer, long long inner, int grid, bool in_place, unsigned seed) {
    using T = typename CfgA::T;
    using T2 = vec2<T>;
    constexpr int W = CfgA::W, N1 = CfgA::N, N2 = CfgB::N, N = N1 * N2;
    const long long total = outer * N * inner;
    std::mt19937_64 rng(seed);
    std::normal_distribution<double> nd(0.0, 1.0);
    std::vector<T2> in_c(total), out_c(total), ref_in(total);
    for (long long i = 0; i < total; ++i) { in_c[i].x = (T)nd(rng); in_c[i].y = (T)nd(rng); out_c[i].x = NAN; out_c[i].y = NAN; }
    ref_in = in_c;
    const long long slot_elems = (long long)(N1 - KS) * N2 * W;
    std::vector<T2> scratch((size_t)grid * slot_elems + 1), smem_i((size_t)KS * N2 * W + 1);
    auto t1 = make_fs_table<T>(N, N2, CfgA::TPC, 1), t2 = make_fs_table<T>(N, N2, CfgA::E, CfgA::TPC);
    PassParams<T> p{};
    p.in0 = reinterpret_cast<const T*>(in_c.data());
    p.out0 = reinterpret_cast<T*>(in_place ? in_c.data() : out_c.data());
    p.inner = inner; p.inner_blocks = inner / W; p.outer_stride = (long long)N * inner;
    p.n_tiles = outer * p.inner_blocks;
    p.out_inner = inner; p.out_outer_stride = p.outer_stride; p.out_blk_log2 = -1; p.in_blk_log2 = -1;
    p.scale = (T)0.5; p.scale_mode = 1;
    p.fs_t1 = reinterpret_cast<const T*>(t1.data()); p.fs_t2 = reinterpret_cast<const T*>(t2.data());
    p.scratch = reinterpret_cast<T*>(scratch.data());
    PassParams<T> pa = p, pb = p;
    pa.inner = (long long)N2 * p.inner; pa.out0 = p.scratch; pa.scale_mode = 0; pa.fs_n2 = N2;
    pb.in0 = p.scratch; pb.out_inner = (long long)N1 * p.out_inner; pb.fs_t1 = pb.fs_t2 = nullptr;
    for (int bid = 0; bid < grid; ++bid) {
        T2* slot = scratch.data() + (long long)bid * slot_elems;
        for (long long sidx = bid; sidx < pa.n_tiles; sidx += grid) {
            const long long o = sidx / pa.inner_blocks, ib = sidx - o * pa.inner_blocks;
            for (auto& v : smem_i) { v.x = NAN; v.y = NAN; }
            for (long long i = 0; i < slot_elems; ++i) { slot[i].x = NAN; slot[i].y = NAN; }
            for (int c = 0; c < N2 / CfgA::G; ++c)
                for (int tid = 0; tid < CfgA::THREADS; ++tid) {
                    TileThread<CfgA, false, INV, true, true> th;
                    fused2_setup_a<CfgA, CfgB>(th, tid, c, o * pa.outer_stride + ib * W, p.inner, 0);
                    th.load(pa);
                    th.template compute<0>(pa);
                    th.apply_fs_twiddle(pa);
                    fused2s_store_a<CfgA, CfgB, KS>(th, smem_i.data(), slot, 0ull);
                }
            for (int c = 0; c < N1 / CfgB::G; ++c)
                for (int tid = 0; tid < CfgB::THREADS; ++tid) {
                    TileThread<CfgB, false, INV, false, true> th;
                    fused2_setup_b<CfgA, CfgB>(th, tid, c, o * pb.out_outer_stride + ib * W, p.out_inner, 0);
                    fused2s_load_b<CfgA, CfgB, KS>(th, c * CfgB::G + th.g, smem_i.data(), slot, 0ull);
                    th.template compute<0>(pb);
                    th.store(pb);
                }
        }
    }
    const std::vector<T2>& got_c = in_place ? in_c : out_c;
    double max_err = 0, max_mag = 0;
    std::vector<cld> line(N);
    for (long long o = 0; o < outer; ++o)
        for (long long i = 0; i < inner; ++i) {
            for (long long n = 0; n < N; ++n) {
                const T2 v = ref_in[(o * N + n) * inner + i];
                line[n] = INV ? cld(v.y, v.x) : cld(v.x, v.y);
            }
            ref_fft(line);
            for (long long n = 0; n < N; ++n) {
                cld want = INV ? cld(line[n].imag(), line[n].real()) : line[n];
                want *= 0.5L;
                const T2 g = got_c[(o * N + n) * inner + i];
                double e = (double)std::abs(cld(g.x, g.y) - want);
                if (!(e == e)) e = 1e30;
                if (e > max_err) max_err = e;
                double m = (double)std::abs(want);
                if (m > max_mag) max_mag = m;
            }
        }
    return max_err / (max_mag > 0 ? max_mag : 1);
}

template <class CfgA, class CfgB, int KS>
static void check_fused2s(const char* name) {
    using T = typename CfgA::T;
    const double tol = sizeof(T) == 4 ? 4e-6 : 2e-14;
    double e0 = run_fused2s<CfgA, CfgB, KS, false>(1, 3 * CfgA::W, 2, false, 31);
    double e1 = run_fused2s<CfgA, CfgB, KS, true>(2, CfgA::W, 1, true, 32);
    bool ok = e0 < tol && e1 < tol;
    std::printf("%-44s fused two-step (smem, ks=%d) N=%dx%d  err fwd(oop)=%.2e inv(in place)=%.2e thr=%d %s\n", name, KS, CfgA::N,
                CfgB::N, e0, e1, CfgA::THREADS, ok ? "ok" : "FAIL");
    if (!ok) ++g_fail;
}

// The streamed in-place fused two-step kernel (kernels.cuh fused2p_fft_kernel), in the kernel's own order of events: a
// warp refills the rows it owns (Fused2PRows / Fused2PChunk, as fused2p_refill does, lane by lane and segment by segment)
// in the shared-memory tile; step A reads them back with fused2p_load_a (rows n1 >= KS from global memory), transforms in
// place; in step B sub-tile c every WARP first loads all its lanes, then the rows it has freed are poisoned and refilled
// with the CTA's next super-tile (what its cp.async copies do), then its lanes compute and store -- so a refill that
// touched a row another warp still has to read, or a wrong chunk address, shows up as a wrong result of the current or the
// next super-tile.
template <class CfgA, class CfgB, int KS, bool INV>
static double run_fused2p(long long outer, long long inner, int grid, bool in_place, unsigned seed) {
    using T = typename CfgA::T;
    using T2 = vec2<T>;
    using C = cpx<T>;
    using CH = Fused2PChunk<CfgA>;
    using M = Fused2PRows<CfgA, CfgB>;
    constexpr int W = CfgA::W, N1 = CfgA::N, N2 = CfgB::N, N = N1 * N2, NX = N1 - KS, THREADS = CfgA::THREADS, GB = CfgB::G;
    const long long total = outer * N * inner;
    std::mt19937_64 rng(seed);
    std::normal_distribution<double> nd(0.0, 1.0);
    std::vector<T2> in_c(total), out_c(total), ref_in(total);
    for (long long i = 0; i < total; ++i) { in_c[i].x = (T)nd(rng); in_c[i].y = (T)nd(rng); out_c[i].x = NAN; out_c[i].y = NAN; }
    ref_in = in_c;
    const long long slot_elems = (long long)(NX > 0 ? NX : 1) * N2 * W;
    std::vector<T2> scratch((size_t)grid * slot_elems + 1);
    auto t1 = make_fs_table<T>(N, N2, CfgA::TPC, 1), t2 = make_fs_table<T>(N, N2, CfgA::E, CfgA::TPC);
    PassParams<T> p{};
    p.in0 = reinterpret_cast<const T*>(in_c.data());
    p.out0 = reinterpret_cast<T*>(in_place ? in_c.data() : out_c.data());
    p.inner = inner; p.inner_blocks = inner / W; p.outer_stride = (long long)N * inner;
    p.n_tiles = outer * p.inner_blocks;
    p.out_inner = inner; p.out_outer_stride = p.outer_stride; p.out_blk_log2 = -1; p.in_blk_log2 = -1;
    p.scale = (T)0.5; p.scale_mode = 1;
    p.fs_t1 = reinterpret_cast<const T*>(t1.data()); p.fs_t2 = reinterpret_cast<const T*>(t2.data());
    p.scratch = reinterpret_cast<T*>(scratch.data());
    PassParams<T> pa = p, pb = p;
    pa.inner = (long long)N2 * p.inner; pa.out0 = p.scratch; pa.scale_mode = 0; pa.fs_n2 = N2;
    pb.in0 = p.scratch; pb.out_inner = (long long)N1 * p.out_inner; pb.fs_t1 = pb.fs_t2 = nullptr;
    auto tile_base = [&](long long sidx) {
        const long long o = sidx / pa.inner_blocks, ib = sidx - o * pa.inner_blocks;
        return o * pa.outer_stride + ib * W;
    };
    // one warp's refill of the rows k0 .. k0 + RPW - 1 from the tile at `base` (fused2p_refill, all lanes and segments)
    auto refill = [&](std::vector<T2>& smem, long long base, int k0) {
        for (int lane = 0; lane < 32; ++lane) {
            const int q = CH::elem(lane), pl = (int)CH::row(lane);
            for (int sg = 0; sg < M::NSEG; ++sg)
                for (int it = 0; it < M::ITERS; ++it) {
                    if (M::k_of(k0, it * M::PPI + pl) >= KS) continue;
                    const long long n = M::row(k0, sg, it * M::PPI + pl);
                    for (int e = 0; e < CH::EPC; ++e) smem[n * W + q + e] = in_c[base + n * inner + q + e];
                }
        }
    };
    for (int bid = 0; bid < grid; ++bid) {
        T2* slot = scratch.data() + (long long)bid * slot_elems;
        std::vector<T2> smem_i((size_t)KS * N2 * W + 1);
        for (auto& v : smem_i) { v.x = NAN; v.y = NAN; }
        if (bid < pa.n_tiles)
            for (int c = 0; c < N1 / GB; ++c)
                for (int warp = 0; warp < THREADS / 32; ++warp) refill(smem_i, tile_base(bid), c * GB + warp * M::RPW);
        for (long long sidx = bid; sidx < pa.n_tiles; sidx += grid) {
            const long long in_base = tile_base(sidx);
            const long long o = sidx / pa.inner_blocks, ib = sidx - o * pa.inner_blocks;
            for (long long i = 0; i < slot_elems; ++i) { slot[i].x = NAN; slot[i].y = NAN; }
            for (int c = 0; c < N2 / CfgA::G; ++c)
                for (int tid = 0; tid < THREADS; ++tid) {
                    TileThread<CfgA, false, INV, true, true> th;
                    C extra[NX > 0 ? NX : 1];
                    for (int i = 0; i < NX; ++i)
                        extra[i] = ld_c(in_c.data() + in_base + ((long long)(KS + i) * N2 + c * CfgA::G + tid / W) * inner + tid % W);
                    fused2_setup_a<CfgA, CfgB>(th, tid, c, in_base, p.inner, 0);
                    C lo[4], hi[N1 / 4];
                    th.fs_base_load(pa, lo, hi);
                    fused2p_load_a<CfgA, CfgB, KS>(th, smem_i.data(), extra);
                    th.template compute<0>(pa);
                    th.fs_base_apply(lo, hi);
                    fused2s_store_a<CfgA, CfgB, KS>(th, smem_i.data(), slot, 0ull);
                }
            const long long s2 = sidx + grid;
            for (int c = 0; c < N1 / GB; ++c)
                for (int warp = 0; warp < THREADS / 32; ++warp) {
                    std::vector<TileThread<CfgB, false, INV, false, true>> th(32);
                    for (int lane = 0; lane < 32; ++lane) {
                        const int tid = warp * 32 + lane;
                        fused2_setup_b<CfgA, CfgB>(th[lane], tid, c, o * pb.out_outer_stride + ib * W, p.out_inner, 0);
                        fused2s_load_b<CfgA, CfgB, KS>(th[lane], c * GB + th[lane].g, smem_i.data(), slot, 0ull);
                    }
                    const int k0 = c * GB + warp * M::RPW;
                    for (int k = k0; k < k0 + M::RPW && k < KS; ++k)      // freed rows: poison, then refill if there is a next tile
                        for (int e = 0; e < N2 * W; ++e) { smem_i[(size_t)k * N2 * W + e].x = NAN; smem_i[(size_t)k * N2 * W + e].y = NAN; }
                    if (s2 < pa.n_tiles) refill(smem_i, tile_base(s2), k0);
                    for (int lane = 0; lane < 32; ++lane) {
                        th[lane].template compute<0>(pb);
                        th[lane].store(pb);
                    }
                }
        }
    }
    const std::vector<T2>& got_c = in_place ? in_c : out_c;
    double max_err = 0, max_mag = 0;
    std::vector<cld> line(N);
    for (long long o = 0; o < outer; ++o)
        for (long long i = 0; i < inner; ++i) {
            for (long long n = 0; n < N; ++n) {
                const T2 v = ref_in[(o * N + n) * inner + i];
                line[n] = INV ? cld(v.y, v.x) : cld(v.x, v.y);
            }
            ref_fft(line);
            for (long long n = 0; n < N; ++n) {
                cld want = INV ? cld(line[n].imag(), line[n].real()) : line[n];
                want *= 0.5L;
                const T2 g = got_c[(o * N + n) * inner + i];
                double e = (double)std::abs(cld(g.x, g.y) - want);
                if (!(e == e)) e = 1e30;
                if (e > max_err) max_err = e;
                double m = (double)std::abs(want);
                if (m > max_mag) max_mag = m;
            }
        }
    return max_err / (max_mag > 0 ? max_mag : 1);
}

template <class CfgA, class CfgB, int KS>
static void check_fused2p(const char* name) {
    using T = typename CfgA::T;
    const double tol = sizeof(T) == 4 ? 4e-6 : 2e-14;
    double e0 = run_fused2p<CfgA, CfgB, KS, false>(1, 5 * CfgA::W, 2, false, 51);     // 5 super-tiles over 2 CTAs: refills + tails
    double e1 = run_fused2p<CfgA, CfgB, KS, true>(3, CfgA::W, 1, true, 52);           // one CTA, in place
    bool ok = e0 < tol && e1 < tol;
    std::printf("%-44s fused two-step (streamed, ks=%d) N=%dx%d  err fwd(oop)=%.2e inv(in place)=%.2e thr=%d %s\n", name, KS, CfgA::N,
                CfgB::N, e0, e1, CfgA::THREADS, ok ? "ok" : "FAIL");
    if (!ok) ++g_fail;
}

// The short-row kernel (kernels.cuh row_shfl_kernel): the lanes of a row run ShflRow stage by stage, the warp shuffle is
// replaced by handing every lane the registers of lane ^ mask; loads / stores use the kernel's addressing (positions 2l, 2l+1
// in, X[brev(l)] and X[brev(l) + N/2] out), interleaved and split, forward and conjugated, with the scale applied.
template <int LOG2N, bool INV, bool SPLIT>
static double run_shfl(long long rows, unsigned seed) {
    using R = ShflRow<LOG2N, INV>;
    using C = cpx<float>;
    constexpr int N = R::N, LP = R::LP;
    std::mt19937_64 rng(seed);
    std::normal_distribution<double> nd(0.0, 1.0);
    std::vector<float> re(rows * N), im(rows * N), ore(rows * N, NAN), oim(rows * N, NAN);
    for (auto& v : re) v = (float)nd(rng);
    for (auto& v : im) v = (float)nd(rng);
    for (long long row = 0; row < rows; ++row) {
        std::vector<R> th(LP);
        std::vector<std::array<C, 2>> v(LP), o(LP);
        for (int l = 0; l < LP; ++l) {
            th[l].init(l);
            for (int b = 0; b < 2; ++b) v[l][b] = cmake<float>(re[row * N + 2 * l + b], im[row * N + 2 * l + b]);
        }
        static_for<0, R::NST>([&](auto sc) {
            constexpr int s = decltype(sc)::value;
            constexpr int mask = (N >> (s + 1)) >> 1;
            for (int l = 0; l < LP; ++l) o[l] = v[l ^ mask];
            for (int l = 0; l < LP; ++l) {
                C a[2] = {v[l][0], v[l][1]}, b[2] = {o[l][0], o[l][1]};
                th[l].template stage<s>(a, b);
                v[l] = {a[0], a[1]};
            }
        });
        for (int l = 0; l < LP; ++l) {
            C a[2] = {v[l][0], v[l][1]};
            R::last(a);
            R::scale(a, 0.5f, 1);
            const int k0 = R::out_index(l);
            for (int b = 0; b < 2; ++b) { ore[row * N + k0 + b * (N / 2)] = a[b].x; oim[row * N + k0 + b * (N / 2)] = a[b].y; }
        }
    }
    (void)SPLIT;      // the two layouts differ in the kernel's load / store instructions only; the emulation holds (re, im) apart
    double max_err = 0, max_mag = 0;
    std::vector<cld> line(N);
    for (long long row = 0; row < rows; ++row) {
        for (int n = 0; n < N; ++n) line[n] = INV ? cld(im[row * N + n], re[row * N + n]) : cld(re[row * N + n], im[row * N + n]);
        ref_fft(line);
        for (int n = 0; n < N; ++n) {
            cld want = INV ? cld(line[n].imag(), line[n].real()) : line[n];
            want *= 0.5L;
            double e = (double)std::abs(cld(ore[row * N + n], oim[row * N + n]) - want);
            if (!(e == e)) e = 1e30;
            max_err = std::max(max_err, e);
            max_mag = std::max(max_mag, (double)std::abs(want));
        }
    }
    return max_err / (max_mag > 0 ? max_mag : 1);
}

// ... and the four-elements-per-lane version (ShflRow4: 16-byte loads AND stores): positions 2l + b + (N/2) h in, X[h + 2r +
// (N/2) b] out, exactly as row_shfl4_kernel addresses them.
template <int LOG2N, bool INV>
static double run_shfl4(long long rows, unsigned seed) {
    using R = ShflRow4<LOG2N, INV>;
    using C = cpx<float>;
    constexpr int N = R::N, LP = R::LP;
    std::mt19937_64 rng(seed);
    std::normal_distribution<double> nd(0.0, 1.0);
    std::vector<float> re(rows * N), im(rows * N), ore(rows * N, NAN), oim(rows * N, NAN);
    for (auto& v : re) v = (float)nd(rng);
    for (auto& v : im) v = (float)nd(rng);
    for (long long row = 0; row < rows; ++row) {
        std::vector<R> th(LP);
        std::vector<std::array<C, 4>> v(LP), o(LP);
        for (int l = 0; l < LP; ++l) {
            th[l].init(l);
            for (int h = 0; h < 2; ++h)
                for (int b = 0; b < 2; ++b) {
                    const long long p = row * N + 2 * l + b + (N / 2) * h;
                    v[l][2 * h + b] = cmake<float>(re[p], im[p]);
                }
            C a[4] = {v[l][0], v[l][1], v[l][2], v[l][3]};
            th[l].first(a);
            v[l] = {a[0], a[1], a[2], a[3]};
        }
        static_for<0, R::NSX>([&](auto sc) {
            constexpr int sx = decltype(sc)::value;
            constexpr int mask = (N >> (sx + 2)) >> 1;
            for (int l = 0; l < LP; ++l) o[l] = v[l ^ mask];
            for (int l = 0; l < LP; ++l) {
                C a[4] = {v[l][0], v[l][1], v[l][2], v[l][3]}, b[4] = {o[l][0], o[l][1], o[l][2], o[l][3]};
                th[l].template stage<sx>(a, b);
                v[l] = {a[0], a[1], a[2], a[3]};
            }
        });
        for (int l = 0; l < LP; ++l) {
            C a[4] = {v[l][0], v[l][1], v[l][2], v[l][3]};
            R::last(a);
            R::scale(a, 0.5f, 1);
            const int r = R::out_index(l);
            for (int h = 0; h < 2; ++h)
                for (int b = 0; b < 2; ++b) {
                    const long long k = row * N + h + 2 * r + (N / 2) * b;
                    ore[k] = a[2 * h + b].x; oim[k] = a[2 * h + b].y;
                }
        }
    }
    double max_err = 0, max_mag = 0;
    std::vector<cld> line(N);
    for (long long row = 0; row < rows; ++row) {
        for (int n = 0; n < N; ++n) line[n] = INV ? cld(im[row * N + n], re[row * N + n]) : cld(re[row * N + n], im[row * N + n]);
        ref_fft(line);
        for (int n = 0; n < N; ++n) {
            cld want = INV ? cld(line[n].imag(), line[n].real()) : line[n];
            want *= 0.5L;
            double e = (double)std::abs(cld(ore[row * N + n], oim[row * N + n]) - want);
            if (!(e == e)) e = 1e30;
            max_err = std::max(max_err, e);
            max_mag = std::max(max_mag, (double)std::abs(want));
        }
    }
    return max_err / (max_mag > 0 ? max_mag : 1);
}

template <int LOG2N>
static void check_shfl4(const char* name) {
    double e0 = run_shfl4<LOG2N, false>(5, 81), e1 = run_shfl4<LOG2N, true>(3, 82);
    bool ok = e0 < 2e-6 && e1 < 2e-6;
    std::printf("%-44s short rows (4 per lane, lane shuffles) N=%d  err fwd=%.2e inv=%.2e %s\n", name, 1 << LOG2N, e0, e1, ok ? "ok" : "FAIL");
    if (!ok) ++g_fail;
}

template <int LOG2N>
static void check_shfl(const char* name) {
    double e0 = run_shfl<LOG2N, false, false>(5, 61), e1 = run_shfl<LOG2N, true, false>(3, 62);
    bool ok = e0 < 2e-6 && e1 < 2e-6;
    std::printf("%-44s short rows (lane shuffles) N=%d  err fwd=%.2e inv=%.2e %s\n", name, 1 << LOG2N, e0, e1, ok ? "ok" : "FAIL");
    if (!ok) ++g_fail;
}

// The lane-pair fused two-step kernel (kernels.cuh fused2w_fft_kernel): the same thread-level functions (Fused2W /
// PairFFT), with the warp shuffle replaced by handing each lane its partner's `send` array.
template <int LOG2A, int LOG2B, int KS, bool INV>
static double run_fused2w(long long outer, long long inner, int grid, bool in_place, unsigned seed) {
    using F = Fused2W<LOG2A, LOG2B, 16, KS, INV>;
    using T = float;
    using T2 = vec2<T>;
    using C = cpx<T>;
    constexpr int W = F::W, N1 = F::N1, N2 = F::N2, N = N1 * N2, G = 512 / (2 * W), THREADS = 512;
    const long long total = outer * N * inner;
    std::mt19937_64 rng(seed);
    std::normal_distribution<double> nd(0.0, 1.0);
    std::vector<T2> in_c(total), out_c(total), ref_in(total);
    for (long long i = 0; i < total; ++i) { in_c[i].x = (T)nd(rng); in_c[i].y = (T)nd(rng); out_c[i].x = NAN; out_c[i].y = NAN; }
    ref_in = in_c;
    const long long slot_elems = (long long)(N1 - KS) * N2 * W;
    std::vector<T2> scratch((size_t)grid * slot_elems + 1), smem_i((size_t)KS * N2 * W + 1);
    auto tab = make_fs_table<T>(N, N2, N1, 1);                       // [k1][N2]: w_N^(k1*n2), as api.cu fused_tables() requests it
    const long long inner_blocks = inner / W, outer_stride = (long long)N * inner, n_super = outer * inner_blocks;
    T2* out_base = in_place ? in_c.data() : out_c.data();
    for (int bid = 0; bid < grid; ++bid) {
        T2* slot = scratch.data() + (long long)bid * slot_elems;
        for (long long sidx = bid; sidx < n_super; sidx += grid) {
            const long long o = sidx / inner_blocks, ib = sidx - o * inner_blocks;
            for (auto& v : smem_i) { v.x = NAN; v.y = NAN; }
            for (long long i = 0; i < slot_elems; ++i) { slot[i].x = NAN; slot[i].y = NAN; }
            for (int c = 0; c < N2 / G; ++c)
                for (int pair = 0; pair < THREADS / 2; ++pair) {         // the two lanes (w, t = 0 / 1) of column (g, w)
                    const int w = pair % W, g = pair / W;
                    C v[2][F::EA], send[2][F::HA];
                    for (int t = 0; t < 2; ++t) {
                        F::a_load(v[t], in_c.data() + o * outer_stride + ib * W + w + (long long)(c * G + g) * inner, (long long)N2 * inner, t, 0ull);
                        F::PA::pre(v[t], t, send[t]);
                    }
                    for (int t = 0; t < 2; ++t) {
                        F::PA::post(v[t], t, send[1 - t]);
                        F::a_store(v[t], t, c * G + g, w, tab.data(), smem_i.data(), slot, 0ull);
                    }
                }
            std::vector<T2> result((size_t)N * W);
            for (int c = 0; c < N1 / G; ++c)
                for (int pair = 0; pair < THREADS / 2; ++pair) {
                    const int w = pair % W, g = pair / W, k1 = c * G + g;
                    C v[2][F::EB], send[2][F::HB];
                    for (int t = 0; t < 2; ++t) {
                        F::b_load(v[t], t, k1, w, smem_i.data(), slot, 0ull);
                        F::PB::pre(v[t], t, send[t]);
                    }
                    for (int t = 0; t < 2; ++t) {
                        F::PB::post(v[t], t, send[1 - t]);
                        F::b_store(v[t], t, out_base + o * outer_stride + ib * W + w + (long long)k1 * inner, (long long)N1 * inner, (T)0.5, 1, 0ull);
                    }
                }
        }
    }
    const std::vector<T2>& got_c = in_place ? in_c : out_c;
    double max_err = 0, max_mag = 0;
    std::vector<cld> line(N);
    for (long long o = 0; o < outer; ++o)
        for (long long i = 0; i < inner; ++i) {
            for (long long n = 0; n < N; ++n) {
                const T2 v = ref_in[(o * N + n) * inner + i];
                line[n] = INV ? cld(v.y, v.x) : cld(v.x, v.y);
            }
            ref_fft(line);
            for (long long n = 0; n < N; ++n) {
                cld want = INV ? cld(line[n].imag(), line[n].real()) : line[n];
                want *= 0.5L;
                const T2 g = got_c[(o * N + n) * inner + i];
                double e = (double)std::abs(cld(g.x, g.y) - want);
                if (!(e == e)) e = 1e30;
                if (e > max_err) max_err = e;
                double m = (double)std::abs(want);
                if (m > max_mag) max_mag = m;
            }
        }
    return max_err / (max_mag > 0 ? max_mag : 1);
}

template <int LOG2A, int LOG2B, int KS>
static void check_fused2w(const char* name) {
    double e0 = run_fused2w<LOG2A, LOG2B, KS, false>(1, 48, 2, false, 41);
    double e1 = run_fused2w<LOG2A, LOG2B, KS, true>(2, 16, 1, true, 42);
    bool ok = e0 < 4e-6 && e1 < 4e-6;
    std::printf("%-44s fused two-step (lane pairs, ks=%d) N=%dx%d  err fwd(oop)=%.2e inv(in place)=%.2e %s\n", name, KS, 1 << LOG2A,
                1 << LOG2B, e0, e1, ok ? "ok" : "FAIL");
    if (!ok) ++g_fail;
}

#define CHKT(T, L, G, R0, R1, R2, R3) \
    check_staged<TileCfg<T, L, 1, G, R0, R1, R2, R3>>(#T " n" #L " g" #G " r" #R0 "x" #R1 "x" #R2 "x" #R3 " tma", 2 * (G) + 1);

#define CHKC(T, L, W, G, R0, R1, R2, R3) \
    check_staged<TileCfg<T, L, W, G, R0, R1, R2, R3>>(#T " n" #L " w" #W " g" #G " r" #R0 "x" #R1 "x" #R2 "x" #R3 " tmac", 3);

#define CHK(T, L, W, G, R0, R1, R2, R3, OUTER, INNER) \
    check<TileCfg<T, L, W, G, R0, R1, R2, R3>>(#T " n" #L " w" #W " g" #G " r" #R0 "x" #R1 "x" #R2 "x" #R3, OUTER, INNER); \
    if ((W) > 1 && (L) >= 3) check_blocked<TileCfg<T, L, W, G, R0, R1, R2, R3>>(#T " n" #L " w" #W " g" #G " r" #R0 "x" #R1 "x" #R2 "x" #R3, OUTER, INNER, (L) >= 8 ? 8 : 2);

#define CHKXS(T, L, W, G, R0, R1, R2, R3) \
    check_xslab_rows<TileCfg<T, L, W, G, R0, R1, R2, R3>>(#T " n" #L " w" #W " g" #G " r" #R0 "x" #R1 "x" #R2 "x" #R3); \
    check_xslab_pull_rows<TileCfg<T, L, W, G, R0, R1, R2, R3>>(#T " n" #L " w" #W " g" #G " r" #R0 "x" #R1 "x" #R2 "x" #R3);

#define CHKF(T, L, W, G, R0, R1, R2, R3) \
    check_fourstep<TileCfg<T, L, W, G, R0, R1, R2, R3>>(#T " n" #L " w" #W " g" #G " r" #R0 "x" #R1 "x" #R2 "x" #R3);

int main() {
#include "emu_cases.inc"
    std::printf(g_fail ? "FAILED %d\n" : "ALL OK\n", g_fail);
    return g_fail ? 1 : 0;
}
