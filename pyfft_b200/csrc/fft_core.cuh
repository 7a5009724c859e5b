// b2fft core: in-register radix butterflies and the "tile FFT" thread program.
//
// One CTA processes G tiles.  A tile is W adjacent columns of an [outer][N][inner]
// array (FFT along N, element stride `inner`); W = 1, inner = 1 is the contiguous
// (row / X-axis) case, W > 1 the strided-axis (Y / Z) case where the W columns are
// W consecutive elements of the contiguous dimension, so every global access of a
// warp covers whole 128-byte lines.  Each thread owns E = R0 elements of one
// column for the whole transform: data goes HBM -> registers once, through S
// Stockham/DIF stages with S-1 shared-memory exchanges, and registers -> HBM once.
// This replaces the reference's localKernel (pyfft/kernel.mako:725-803) and
// globalKernel (805-1047); the mathematics per stage is the same DIF step
//     n_rem = m + M*j  --radix-R butterfly over j, times w_{N_s}^{m k}-->  (K + P*k, m)
// with natural-order output, but twiddles come from precomputed tables instead of
// on-the-fly sincos (kernel.mako:35-44,566-597).
//
// Everything here is __host__ __device__ so tests/host_emu can run the exact same
// thread program on the CPU, phase by phase (test infrastructure; the product only
// ever launches the __global__ wrappers in kernels.cuh).
#pragma once

#include <stdint.h>
#include <cmath>
#include <type_traits>
#include <utility>

#if defined(__CUDACC__)
#define B2_HD __host__ __device__ __forceinline__
#else
#define B2_HD inline
#endif

namespace b2 {

// ------------------------------------------------------------------ small utilities
template <typename T> struct vec2;
template <> struct alignas(8) vec2<float> { float x, y; };
template <> struct alignas(16) vec2<double> { double x, y; };

template <int B, int E_, class F>
B2_HD void static_for(F&& f) {
    if constexpr (B < E_) {
        f(std::integral_constant<int, B>{});
        static_for<B + 1, E_>(static_cast<F&&>(f));
    }
}

constexpr int ilog2(int n) { return n <= 1 ? 0 : 1 + ilog2(n >> 1); }
constexpr int brev(int v, int bits) {
    int r = 0;
    for (int i = 0; i < bits; ++i) r |= ((v >> i) & 1) << (bits - 1 - i);
    return r;
}

// cos(2*pi*e/64), e = 0..16 (first quadrant incl. end points), round-to-nearest doubles.
constexpr double kCosQ64[17] = {
    1.0,
    0.99518472667219688624, 0.98078528040323044913, 0.95694033573220886494,
    0.92387953251128675613, 0.88192126434835502971, 0.83146961230254523708,
    0.77301045336273696081, 0.70710678118654752440, 0.63439328416364549822,
    0.55557023301960222474, 0.47139673682599764856, 0.38268343236508977173,
    0.29028467725446236764, 0.19509032201612826785, 0.09801714032956060199,
    0.0};
constexpr double cos64(int e) {  // e in [0, 32]
    return e <= 16 ? kCosQ64[e] : -kCosQ64[32 - e];
}
constexpr double sin64(int e) {  // e in [0, 32]
    return e <= 16 ? kCosQ64[16 - e] : kCosQ64[e - 16];
}

// ------------------------------------------------------------------ complex values in registers
// cpx<T> is one complex number held by a thread.  For double (and for float on the host, where
// tests/host_emu runs this code) it is a plain scalar pair.  For float on the device it is ONE
// 64-bit register pair operated on with Blackwell's packed-FP32 instructions (PTX add/mul/fma
// .f32x2 -> SASS FADD2 / FMUL2 / FFMA2): a complex add is one instruction instead of two, a
// complex multiply three instead of four, which is what the radix butterflies are made of.
// The real/imaginary swap and scalar broadcast these formulas need are operand modifiers of the
// packed instructions (ptxas folds the mov.b64 re-packs below into .LO_HI / .F32 operands).
template <typename T> struct cpx { T x, y; };
template <typename T> struct is_packed { static constexpr bool value = false; };
#if defined(__CUDA_ARCH__)
template <> struct cpx<float> { unsigned long long v; };
template <> struct is_packed<float> { static constexpr bool value = true; };
#endif

#if defined(__CUDA_ARCH__)
namespace pk {
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float lo, float hi) { u64 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
__device__ __forceinline__ void unpack(u64 a, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a)); }
__device__ __forceinline__ u64 swp(u64 a) { float lo, hi; unpack(a, lo, hi); return pack(hi, lo); }
__device__ __forceinline__ u64 blo(u64 a) { float lo, hi; unpack(a, lo, hi); return pack(lo, lo); }
__device__ __forceinline__ u64 bhi(u64 a) { float lo, hi; unpack(a, lo, hi); return pack(hi, hi); }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
}  // namespace pk
#endif

template <typename T> B2_HD cpx<T> cmake(T x, T y) {
#if defined(__CUDA_ARCH__)
    if constexpr (is_packed<T>::value) { cpx<T> c; c.v = pk::pack(x, y); return c; } else
#endif
    { cpx<T> c; c.x = x; c.y = y; return c; }
}
template <typename T> B2_HD void csplit(const cpx<T>& c, T& x, T& y) {
#if defined(__CUDA_ARCH__)
    if constexpr (is_packed<T>::value) { pk::unpack(c.v, x, y); } else
#endif
    { x = c.x; y = c.y; }
}
template <typename T> B2_HD cpx<T> cadd(const cpx<T>& a, const cpx<T>& b) {
#if defined(__CUDA_ARCH__)
    if constexpr (is_packed<T>::value) { cpx<T> c; c.v = pk::add2(a.v, b.v); return c; } else
#endif
    { cpx<T> c; c.x = a.x + b.x; c.y = a.y + b.y; return c; }
}
template <typename T> B2_HD cpx<T> csub(const cpx<T>& a, const cpx<T>& b) {
#if defined(__CUDA_ARCH__)
    if constexpr (is_packed<T>::value) { cpx<T> c; c.v = pk::sub2(a.v, b.v); return c; } else
#endif
    { cpx<T> c; c.x = a.x - b.x; c.y = a.y - b.y; return c; }
}
// a * s, s real
template <typename T> B2_HD cpx<T> cscale(const cpx<T>& a, T s) {
#if defined(__CUDA_ARCH__)
    if constexpr (is_packed<T>::value) { cpx<T> c; c.v = pk::mul2(a.v, pk::pack(s, s)); return c; } else
#endif
    { cpx<T> c; c.x = a.x * s; c.y = a.y * s; return c; }
}
// a * w (CONJ: a * conj(w)), w a run-time twiddle
template <bool CONJ, typename T> B2_HD cpx<T> cmul(const cpx<T>& a, const cpx<T>& w) {
#if defined(__CUDA_ARCH__)
    if constexpr (is_packed<T>::value) {
        // (ar*wr, ai*wr) -/+ (ai*wi, ar*wi) * (1, -1): three packed instructions
        const pk::u64 d = pk::mul2(a.v, pk::blo(w.v));
        const pk::u64 e = pk::mul2(pk::swp(a.v), pk::bhi(w.v));
        cpx<T> c;
        c.v = pk::fma2(e, CONJ ? pk::pack(1.0f, -1.0f) : pk::pack(-1.0f, 1.0f), d);
        return c;
    } else
#endif
    {
        cpx<T> c;
        if constexpr (CONJ) { c.x = a.x * w.x + a.y * w.y; c.y = a.y * w.x - a.x * w.y; }
        else { c.x = a.x * w.x - a.y * w.y; c.y = a.x * w.y + a.y * w.x; }
        return c;
    }
}
// a * exp(-2*pi*i*e/64) (CONJ: exp(+...)), e in [0, 32) compile-time.
template <int e, bool CONJ, typename T> B2_HD cpx<T> cmulc(const cpx<T>& a) {
    if constexpr (e == 0) {
        return a;
    } else {
#if defined(__CUDA_ARCH__)
        if constexpr (is_packed<T>::value) {
            cpx<T> c;
            if constexpr (e == 16) {        // * -i: (ai, -ar);  * +i: (-ai, ar)
                c.v = pk::mul2(pk::swp(a.v), CONJ ? pk::pack(-1.0f, 1.0f) : pk::pack(1.0f, -1.0f));
            } else {                        // (ar*c, ai*c) + (ai, ar) * (s, -s)   [CONJ: (-s, s)]
                constexpr float cs = (float)cos64(e), sn = (float)sin64(e);
                c.v = pk::fma2(pk::swp(a.v), CONJ ? pk::pack(-sn, sn) : pk::pack(sn, -sn), pk::mul2(a.v, pk::pack(cs, cs)));
            }
            return c;
        } else
#endif
        {
            cpx<T> c;
            const T dr = a.x, di = a.y;
            if constexpr (e == 16) {
                if constexpr (CONJ) { c.x = -di; c.y = dr; } else { c.x = di; c.y = -dr; }
            } else if constexpr (e == 8) {
                constexpr T h = (T)0.70710678118654752440;
                if constexpr (CONJ) { c.x = (dr - di) * h; c.y = (dr + di) * h; }
                else { c.x = (dr + di) * h; c.y = (di - dr) * h; }
            } else if constexpr (e == 24) {
                constexpr T h = (T)0.70710678118654752440;
                if constexpr (CONJ) { c.x = -(dr + di) * h; c.y = (dr - di) * h; }
                else { c.x = (di - dr) * h; c.y = -(dr + di) * h; }
            } else {
                constexpr T cs = (T)cos64(e);
                constexpr T sn = (T)sin64(e);
                if constexpr (CONJ) { c.x = dr * cs - di * sn; c.y = di * cs + dr * sn; }
                else { c.x = dr * cs + di * sn; c.y = di * cs - dr * sn; }
            }
            return c;
        }
    }
}

// In-register radix-R DIF FFT on v[OFF .. OFF+R), forward (CONJ: inverse, conjugated roots).
// X[k] ends up at register OFF + brev(k, log2 R) (no data movement for the permutation).
template <int R, int OFF, bool CONJ, typename T>
B2_HD void butterfly(cpx<T>* v) {
    static_assert(R >= 1 && R <= 64 && (R & (R - 1)) == 0, "radix must be a power of two <= 64");
    constexpr int LG = ilog2(R);
    static_for<0, LG>([&](auto lc) {
        constexpr int len = R >> decltype(lc)::value;
        constexpr int half = len / 2;
        static_for<0, R / 2>([&](auto qc) {
            constexpr int q = decltype(qc)::value;
            constexpr int ia = OFF + (q / half) * len + (q % half);
            constexpr int ib = ia + half;
            constexpr int e = (q % half) * (64 / len);
            const cpx<T> u = v[ia], w = v[ib];
            v[ia] = cadd(u, w);
            v[ib] = cmulc<e, CONJ>(csub(u, w));
        });
    });
}

// ------------------------------------------------------------------ pass parameters
#define B2_MAX_BLOCKS 16
template <typename T>
struct PassParams {
    const T* in0;   // interleaved: vec2<T>* ; split: re plane
    const T* in1;   // split: im plane (else unused)
    T* out0;
    T* out1;
    const T* tw[3];          // per-exchange-stage twiddle tables (vec2<T>), layout [(k-1)*M + m]
    long long n_tiles;       // total number of tiles (columns / W)
    long long inner;         // element stride between consecutive n (1 for the contiguous axis)
    long long inner_blocks;  // inner / W
    long long outer_stride;  // N * inner
    // output side (equal to the input side unless the pass re-lays-out its result)
    long long out_inner;         // element stride between consecutive n in the output
    long long out_outer_stride;  // element stride between consecutive outer blocks in the output
    T scale;                 // multiplier (scale_mode 1) or divisor (scale_mode 2)
    int scale_mode;          // 0: none
    // destination-blocked stores (slab-decomposed multi-GPU transforms): output index n goes to
    // block h = n >> out_blk_log2 at in-block position n & (2^out_blk_log2 - 1); block h lives at
    // out_blk0[h] (+ out_blk1[h] for the split layout), which may be local or peer (NVLink) memory.
    int out_blk_log2;        // < 0: plain stores to out0/out1
    T* out_blk0[B2_MAX_BLOCKS];
    T* out_blk1[B2_MAX_BLOCKS];
    // source-blocked loads (inverse of the above, contiguous-axis passes only): input index n comes from block
    // h = n >> in_blk_log2 at in-block position n & (2^in_blk_log2 - 1) of in_blk[h], local or peer (NVLink) memory
    int in_blk_log2;         // < 0 (or 0 with in_blk[0] == nullptr): plain loads from in0/in1
    const T* in_blk[B2_MAX_BLOCKS];
    // two-level outer index (slab exchange passes): when outer_div > 0 the outer index o is split as
    // (o_hi, o_lo) = (o / outer_div, o % outer_div) and the tile starts at
    //   in:  o_hi*in_stride_hi  + o_lo*outer_stride       out: o_hi*out_stride_hi + o_lo*out_outer_stride
    // so a pass can walk a sub-range of one axis of a larger array and re-lay-out its result.
    long long outer_div, in_stride_hi, out_stride_hi;
    int fs_bulk;              // four-step pass A, inner0 == 1: whole output rows leave as cp.async.bulk (kernels.cuh)
    int tma_store;            // strided persistent kernels: output tiles leave as TMA tensor stores (kernels.cuh)
    int blk_bulk;             // blocked stores of a contiguous-axis pass go out as TMA bulk copies (kernels.cuh)
    int split_bulk;           // split-layout rows of the TMA-staged row kernel: output planes leave as TMA bulk copies
    int max_ctas;             // > 0: cap on the grid of the plain kernels (CTAs then stride over the tiles)
    // four-step "A" pass (FS kernels only): the transformed axis of length N = N1*N2 is split as
    // n = n1*N2 + n2; this pass transforms over n1 (length Cfg::N = N1, element stride N2*inner0,
    // so the pass's `inner` is N2*inner0), multiplies output k1 of column n2 by w_N^(k1*n2) and
    // stores it TRANSPOSED at [outer][n2][k1][inner0] (out_inner = inner0).  The following pass
    // over n2 (element stride N1*inner0) then leaves X[k1 + N1*k2] in natural order.
    int fs_log2_inner;        // log2(inner0)
    long long fs_n2;          // N2
    long long fs_col_stride;  // N1 * inner0: output element stride between consecutive n2
    const T* fs_t1;           // vec2<T> [TPC][N2]: w_N^(t*n2)
    const T* fs_t2;           // vec2<T> [E][N2]:   w_N^(TPC*c*n2)
    // fused two-step strided kernel (kernels.cuh fused2_fft_kernel): stage tables of step B and the scratch
    // buffer that holds the intermediate [n2][k1][W] of one super-tile per resident CTA (stays in L2)
    const T* tw_b[3];
    T* scratch;
    long long scratch_slots;  // number of super-tile slots the scratch buffer holds
    int fused_flags;          // bit 0: discard scratch lines after step B read them, 1: prefetch the next super-tile into L2,
                              // 2: evict-first policy on the DRAM streams, 3: evict-last policy on the scratch slot
    // progress counters (fused2p kernel; slab pipeline): the CTA that has stored super-tile s adds 1 to
    // progress[s / progress_tiles] with release semantics at device scope, so that a consumer on another stream can start
    // on chunk k of the output while this launch is still working on the chunks behind it
    unsigned* progress;       // nullptr: none
    long long progress_tiles; // super-tiles per chunk
};

// ------------------------------------------------------------------ compile-time plan
// ESZ_ = bytes per complex element in shared memory (8 or 16), used for the
// bank-conflict padding rule only.
// TWP_: stage twiddles w^1..w^(R-1) are loaded in full from the table (0) or only w^1..w^3 and
// w^4, w^8, .. with the rest formed as one product each (1): the tables are read through L1, which
// shares its data pipe with the shared-memory exchanges that bound these kernels, while the FP
// pipe has slack (profiles/).  One extra rounding (<= 1 ulp) on 9 of 15 twiddles.
template <typename T_, int LOG2N_, int W_, int G_, int R0_, int R1_ = 1, int R2_ = 1, int R3_ = 1, int TWP_ = 1>
struct TileCfg {
    using T = T_;
    static constexpr int TWP = TWP_;
    static constexpr int LOG2N = LOG2N_;
    static constexpr int N = 1 << LOG2N_;
    static constexpr int W = W_;
    static constexpr int G = G_;
    static constexpr int E = R0_;
    static constexpr int S = 1 + (R1_ > 1) + (R2_ > 1) + (R3_ > 1);
    static constexpr int TPC = N / E;            // threads per column
    static constexpr int THREADS = TPC * W * G;  // threads per CTA
    static_assert(R0_ * R1_ * R2_ * R3_ == N, "radices must multiply to N");
    static_assert(R0_ >= R1_ && R0_ >= R2_ && R0_ >= R3_, "R0 must be the largest radix");
    static_assert(R1_ > 1 || (R2_ == 1 && R3_ == 1), "radices must be packed to the front");
    static_assert(R2_ > 1 || R3_ == 1, "radices must be packed to the front");

    static constexpr int R(int s) { return s == 0 ? R0_ : s == 1 ? R1_ : s == 2 ? R2_ : R3_; }
    static constexpr int P(int s) { return s <= 0 ? 1 : P(s - 1) * R(s - 1); }   // radices done before s
    static constexpr int NS(int s) { return N / P(s); }                         // remaining length
    static constexpr int M(int s) { return NS(s) / R(s); }
    static constexpr int BPT(int s) { return E / R(s); }                        // butterflies / thread

    // conflict-free group: consecutive column-threads that share one shared-memory wavefront
    static constexpr int CF = (sizeof(T) == 4 ? 16 : 8);
    static constexpr int CFE = (CF / W) < 1 ? 1 : (CF / W);
    // exchange s sits between stage s and s+1
    static constexpr int PAD(int s) { return M(s + 1) < CFE ? M(s + 1) : 0; }
    static constexpr int ROW(int s) { return M(s); }
    static constexpr int PADN(int s) { return N + (N / ROW(s)) * PAD(s); }
    static constexpr int max_padn() {
        int m = 0;
        for (int s = 0; s + 1 < S; ++s) m = PADN(s) > m ? PADN(s) : m;
        return m;
    }
    // complex elements per column: the largest exchange, plus 16 bytes so that the four-step pass can
    // stage its output transposed with a bank-conflict-free, 16-byte aligned row pitch (kernels.cuh)
    static constexpr int FS_PADE = 16 / (2 * (int)sizeof(T_));
    static constexpr int COL_SMEM = max_padn() > 0 ? max_padn() + FS_PADE : 0;
    static constexpr long long SMEM_BYTES = (long long)COL_SMEM * W * G * 2 * sizeof(T);
};

// ------------------------------------------------------------------ memory access
// streaming global load of one interleaved complex element (bypasses L1 allocation)
template <typename T>
B2_HD cpx<T> ld_stream_c(const vec2<T>* p) {
#if defined(__CUDA_ARCH__)
    cpx<T> c;
    if constexpr (is_packed<T>::value) {
        asm volatile("ld.global.L1::no_allocate.b64 %0, [%1];" : "=l"(c.v) : "l"(p));
    } else if constexpr (sizeof(T) == 4) {
        asm volatile("ld.global.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(c.x), "=f"(c.y) : "l"(p));
    } else {
        asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(c.x), "=d"(c.y) : "l"(p));
    }
    return c;
#else
    return cmake<T>(p->x, p->y);
#endif
}
// ... the same with an L2 cache policy (createpolicy): the fused two-step kernel streams its DRAM traffic through
// L2 as evict-first and keeps its scratch slot evict-last
template <typename T>
B2_HD cpx<T> ld_stream_c_pol(const vec2<T>* p, unsigned long long pol) {
#if defined(__CUDA_ARCH__)
    cpx<T> c;
    if constexpr (is_packed<T>::value) {
        asm volatile("ld.global.L1::no_allocate.L2::cache_hint.b64 %0, [%1], %2;" : "=l"(c.v) : "l"(p), "l"(pol));
    } else if constexpr (sizeof(T) == 4) {
        asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(c.x), "=f"(c.y) : "l"(p), "l"(pol));
    } else {
        asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(c.x), "=d"(c.y) : "l"(p), "l"(pol));
    }
    return c;
#else
    (void)pol;
    return cmake<T>(p->x, p->y);
#endif
}
template <typename T>
B2_HD void st_c_pol(vec2<T>* p, const cpx<T>& c, unsigned long long pol) {
#if defined(__CUDA_ARCH__)
    if constexpr (is_packed<T>::value) {
        asm volatile("st.global.L2::cache_hint.b64 [%0], %1, %2;" ::"l"(p), "l"(c.v), "l"(pol) : "memory");
    } else if constexpr (sizeof(T) == 4) {
        asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(p), "f"(c.x), "f"(c.y), "l"(pol) : "memory");
    } else {
        asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(c.x), "d"(c.y), "l"(pol) : "memory");
    }
#else
    (void)pol;
    vec2<T> v; v.x = c.x; v.y = c.y; *p = v;
#endif
}
template <typename T>
B2_HD T ld_stream1(const T* p) {
#if defined(__CUDA_ARCH__)
    T v;
    if constexpr (sizeof(T) == 4) {
        asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    } else {
        asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    }
    return v;
#else
    return *p;
#endif
}
// plain load / store of one complex element (shared memory, twiddle tables, global stores)
template <typename T>
B2_HD cpx<T> ld_c(const vec2<T>* p) {
#if defined(__CUDA_ARCH__)
    if constexpr (is_packed<T>::value) { cpx<T> c; c.v = *reinterpret_cast<const unsigned long long*>(p); return c; } else
#endif
    { const vec2<T> v = *p; return cmake<T>(v.x, v.y); }
}
template <typename T>
B2_HD void st_c(vec2<T>* p, const cpx<T>& c) {
#if defined(__CUDA_ARCH__)
    if constexpr (is_packed<T>::value) { *reinterpret_cast<unsigned long long*>(p) = c.v; } else
#endif
    { vec2<T> v; v.x = c.x; v.y = c.y; *p = v; }
}

// ------------------------------------------------------------------ the thread program
// INV (interleaved layout only) runs the same program with conjugated roots, i.e. the unscaled
// inverse DFT.  Split-layout inverses are done by the host swapping the re/im plane pointers
// (IDFT(z) = swap(DFT(swap z))), so SPLIT kernels are always compiled with INV = false.
// POL: interleaved global loads / plain stores carry the L2 cache policies pol_in / pol_out (fused two-step kernel).
template <class Cfg, bool SPLIT, bool INV, bool FS = false, bool POL = false>
struct TileThread {
    using T = typename Cfg::T;
    using T2 = vec2<T>;
    using C = cpx<T>;
    static constexpr int E = Cfg::E, N = Cfg::N, W = Cfg::W, S = Cfg::S, TPC = Cfg::TPC;
    static_assert(!(SPLIT && INV), "split inverses swap the planes on the host");

    C v[E];
    int t, w, g;          // thread-in-column, column-in-tile, tile-in-CTA
    bool active;
    long long base;       // element offset of (n = 0, this column) in the input
    long long obase;      // ... and in the output
    long long fs_n2i;     // FS: this column's n2
    unsigned long long pol_in = 0, pol_out = 0;   // POL: L2 cache policies of the global loads / stores

    B2_HD void setup(int tid, long long bid, const PassParams<T>& p) {
        w = tid % W;
        t = (tid / W) % TPC;
        g = tid / (W * TPC);
        long long tile = bid * Cfg::G + g;
        active = tile < p.n_tiles;
        long long o = 0, ib = tile;
        if (p.inner_blocks > 1) { o = tile / p.inner_blocks; ib = tile - o * p.inner_blocks; }
        else { o = tile; ib = 0; }
        long long obase0;
        if (p.outer_div > 0) {
            const long long oh = o / p.outer_div, ol = o - oh * p.outer_div;
            base = oh * p.in_stride_hi + ol * p.outer_stride + ib * W + w;
            obase0 = oh * p.out_stride_hi + ol * p.out_outer_stride;
        } else {
            base = o * p.outer_stride + ib * W + w;
            obase0 = o * p.out_outer_stride;
        }
        if constexpr (FS) {
            const long long c = ib * W + w;                       // column inside [N2][inner0]
            fs_n2i = c >> p.fs_log2_inner;
            obase = obase0 + fs_n2i * p.fs_col_stride + (c - (fs_n2i << p.fs_log2_inner));
        } else {
            fs_n2i = 0;
            obase = obase0 + ib * W + w;
        }
    }

    B2_HD void clear() {
        static_for<0, E>([&](auto jc) { v[decltype(jc)::value] = cmake<T>((T)0, (T)0); });
    }

    // ---- stage 0 input: element n = t + TPC*j  (BPT(0) == 1)
    B2_HD void load(const PassParams<T>& p) {
        if (!active) { clear(); return; }
        const long long step = (long long)TPC * p.inner;
        if constexpr (SPLIT) {
            const T* pr = p.in0 + base + (long long)t * p.inner;
            const T* pi = p.in1 + base + (long long)t * p.inner;
            static_for<0, E>([&](auto jc) {
                constexpr int j = decltype(jc)::value;
                const T xr = ld_stream1(pr + j * step);
                const T xi = ld_stream1(pi + j * step);
                v[j] = cmake<T>(xr, xi);
            });
        } else {
            if constexpr (W == 1 && !FS && !POL) {
                if (p.in_blk_log2 >= 0 && p.in_blk[0] != nullptr) {      // source-blocked (pull) loads, see PassParams::in_blk
                    const int lg = p.in_blk_log2;
                    const int mask = (1 << lg) - 1;
                    static_for<0, E>([&](auto jc) {
                        constexpr int j = decltype(jc)::value;
                        const int n = t + TPC * j;
                        v[j] = ld_stream_c(reinterpret_cast<const T2*>(p.in_blk[n >> lg]) + base + (n & mask));
                    });
                    return;
                }
            }
            const T2* pc = reinterpret_cast<const T2*>(p.in0) + base + (long long)t * p.inner;
            static_for<0, E>([&](auto jc) {
                constexpr int j = decltype(jc)::value;
                if constexpr (POL) v[j] = ld_stream_c_pol(pc + j * step, pol_in);
                else v[j] = ld_stream_c(pc + j * step);
            });
        }
    }

    // ---- stage 0 input from a shared-memory staging buffer filled by TMA (persistent kernels):
    //      the buffer holds the CTA's G tiles densely, tile g at g*N*W, element (n, w) at n*W + w.
    //      SPLIT: `sre`/`sim` are two planes of T; interleaved: `sre` is a vec2<T> array, `sim` unused.
    B2_HD void load_smem(const void* sre, const void* sim) {
        if (!active) { clear(); return; }
        const long long off = ((long long)g * N + t) * W + w;
        if constexpr (SPLIT) {
            const T* pr = static_cast<const T*>(sre) + off;
            const T* pi = static_cast<const T*>(sim) + off;
            static_for<0, E>([&](auto jc) {
                constexpr int j = decltype(jc)::value;
                v[j] = cmake<T>(pr[j * TPC * W], pi[j * TPC * W]);
            });
        } else {
            const T2* pc = static_cast<const T2*>(sre) + off;
            static_for<0, E>([&](auto jc) {
                constexpr int j = decltype(jc)::value;
                v[j] = ld_c(pc + j * TPC * W);
            });
        }
    }

    // ---- butterflies of stage s, then the stage twiddle (not after the last stage)
    template <int s>
    B2_HD void compute(const PassParams<T>& p) {
        constexpr int R = Cfg::R(s);
        static_for<0, Cfg::BPT(s)>([&](auto ic) { butterfly<R, decltype(ic)::value * R, INV>(v); });
        if constexpr (s + 1 < S) {
            constexpr int M = Cfg::M(s);
            constexpr int LG = ilog2(R);
            const T2* tw = reinterpret_cast<const T2*>(p.tw[s]) + (t % M);
            if constexpr (Cfg::TWP != 0 && R >= 8) {
                // w^k = w^(4a) * w^b, k = 4a + b: load w^1..w^3 and w^4, w^8, .., multiply the rest
                C lo[4], hi[R / 4];
                static_for<1, 4>([&](auto bc) { lo[decltype(bc)::value] = ld_c(tw + (decltype(bc)::value - 1) * M); });
                static_for<1, R / 4>([&](auto ac) { hi[decltype(ac)::value] = ld_c(tw + (4 * decltype(ac)::value - 1) * M); });
                static_for<1, R>([&](auto kc) {
                    constexpr int k = decltype(kc)::value;
                    constexpr int a = k / 4, b = k % 4;
                    C wv;
                    if constexpr (a == 0) wv = lo[b];
                    else if constexpr (b == 0) wv = hi[a];
                    else wv = cmul<false>(hi[a], lo[b]);
                    static_for<0, Cfg::BPT(s)>([&](auto ic) {
                        constexpr int q = decltype(ic)::value * R + brev(k, LG);
                        v[q] = cmul<INV>(v[q], wv);
                    });
                });
            } else {
                static_for<1, R>([&](auto kc) {
                    constexpr int k = decltype(kc)::value;
                    const C wv = ld_c(tw + (k - 1) * M);
                    static_for<0, Cfg::BPT(s)>([&](auto ic) {
                        constexpr int q = decltype(ic)::value * R + brev(k, LG);
                        v[q] = cmul<INV>(v[q], wv);
                    });
                });
            }
        }
    }

    // ---- exchange s: write outputs of stage s.  pos = b + k*(N/R), b = t + TPC*i, padded.
    template <int s>
    B2_HD void xwrite(T2* smem) const {
        constexpr int R = Cfg::R(s), LG = ilog2(R), BPT = Cfg::BPT(s);
        constexpr int ROW = Cfg::ROW(s), PAD = Cfg::PAD(s);
        static_assert(TPC % ROW == 0, "threads per column must be a multiple of the row length");
        constexpr int TW = TPC + (TPC / ROW) * PAD;       // padded distance between consecutive (i + k*BPT)
        T2* dst = smem + ((long long)g * Cfg::COL_SMEM + t + (t / ROW) * PAD) * W + w;
        static_for<0, BPT>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            static_for<0, R>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                constexpr int q = i * R + brev(k, LG);
                st_c(dst + (i + k * BPT) * TW * W, v[q]);
            });
        });
    }

    // ---- exchange s: read inputs of stage s+1.  b' = t + TPC*i' -> (K', m'); element j' at
    //      K'*(M(s)+PAD) + m' + M(s+1)*j'
    template <int s>
    B2_HD void xread(const T2* smem) {
        constexpr int R1 = Cfg::R(s + 1), BPT1 = Cfg::BPT(s + 1);
        constexpr int M1 = Cfg::M(s + 1);
        constexpr int ROWP = Cfg::M(s) + Cfg::PAD(s);
        static_assert(TPC % M1 == 0, "threads per column must be a multiple of M(s+1)");
        const T2* src = smem + ((long long)g * Cfg::COL_SMEM + (t / M1) * ROWP + (t % M1)) * W + w;
        static_for<0, BPT1>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            static_for<0, R1>([&](auto jc) {
                constexpr int j = decltype(jc)::value;
                v[i * R1 + j] = ld_c(src + (i * (TPC / M1) * ROWP + j * M1) * W);
            });
        });
    }

    // ---- four-step pass A: inter-pass twiddle w_N^(k1*n2), k1 = t + TPC*c, as the product of two table entries
    B2_HD void apply_fs_twiddle(const PassParams<T>& p) {
        constexpr int s = S - 1;
        constexpr int R = Cfg::R(s), LG = ilog2(R), BPT = Cfg::BPT(s);
        const T2* t2p = reinterpret_cast<const T2*>(p.fs_t2) + fs_n2i;
        if constexpr (TPC == 1) {
            // one thread per column (t == 0): the first table is all ones, w_N^(k1*n2) is the second table's entry itself
            static_for<0, BPT>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                static_for<0, R>([&](auto kc) {
                    constexpr int k = decltype(kc)::value;
                    constexpr int q = i * R + brev(k, LG);
                    constexpr int c = i + k * BPT;
                    if constexpr (c > 0) v[q] = cmul<INV>(v[q], ld_c(t2p + (long long)c * p.fs_n2));
                });
            });
            return;
        }
        const T2* t1p = reinterpret_cast<const T2*>(p.fs_t1) + (long long)t * p.fs_n2 + fs_n2i;
        const C b = ld_c(t1p);
        static_for<0, BPT>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            static_for<0, R>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                constexpr int q = i * R + brev(k, LG);
                constexpr int c = i + k * BPT;
                if constexpr (c > 0) v[q] = cmul<INV>(v[q], cmul<false>(b, ld_c(t2p + (long long)c * p.fs_n2)));
                else v[q] = cmul<INV>(v[q], b);
            });
        });
    }

    // ---- the same for one thread per column (TPC == 1), split so that the table loads can be issued ahead of the
    //      butterflies: w_N^(k1*n2), k1 = 4a + b, is the product of w^(4a*n2) and w^(b*n2) -- E/4 + 2 table entries per
    //      thread instead of E - 1 (one extra rounding on the products, as for the stage twiddles with TWP)
    B2_HD void fs_base_load(const PassParams<T>& p, C (&lo)[4], C (&hi)[E / 4]) const {
        static_assert(TPC == 1 && S == 1 && E >= 8, "one register FFT per thread");
        const T2* t2p = reinterpret_cast<const T2*>(p.fs_t2) + fs_n2i;
        static_for<1, 4>([&](auto bc) { lo[decltype(bc)::value] = ld_c(t2p + (long long)decltype(bc)::value * p.fs_n2); });
        static_for<1, E / 4>([&](auto ac) { hi[decltype(ac)::value] = ld_c(t2p + (long long)(4 * decltype(ac)::value) * p.fs_n2); });
    }
    B2_HD void fs_base_apply(const C (&lo)[4], const C (&hi)[E / 4]) {
        constexpr int LG = ilog2(E);
        static_for<1, E>([&](auto kc) {
            constexpr int k = decltype(kc)::value;            // k1
            constexpr int a = k / 4, b = k % 4;
            constexpr int q = brev(k, LG);
            C wv;
            if constexpr (a == 0) wv = lo[b];
            else if constexpr (b == 0) wv = hi[a];
            else wv = cmul<false>(hi[a], lo[b]);
            v[q] = cmul<INV>(v[q], wv);
        });
    }

    // ---- scale / normalise (last pass only; pyfft/kernel.py:23-37)
    B2_HD void apply_scale(const PassParams<T>& p) {
        if (p.scale_mode == 1) {
            static_for<0, E>([&](auto jc) { v[decltype(jc)::value] = cscale(v[decltype(jc)::value], p.scale); });
        } else if (p.scale_mode == 2) {
            static_for<0, E>([&](auto jc) {
                T xr, xi;
                csplit(v[decltype(jc)::value], xr, xi);
                v[decltype(jc)::value] = cmake<T>(xr / p.scale, xi / p.scale);
            });
        }
    }

    // output element offset of tile `tile` (contiguous-axis passes: one line per tile)
    static B2_HD long long line_out_base(long long tile, const PassParams<T>& p) {
        if (p.outer_div > 0) {
            const long long oh = tile / p.outer_div, ol = tile - oh * p.outer_div;
            return oh * p.out_stride_hi + ol * p.out_outer_stride;
        }
        return tile * p.out_outer_stride;
    }

    // ---- final output: n = t + TPC*(i + k*BPT) in natural order, scaled.
    //      BLK: destination-blocked stores (see PassParams::out_blk_log2).
    template <bool BLK = false>
    B2_HD void store(const PassParams<T>& p) {
        if (!active) return;
        constexpr int s = S - 1;
        constexpr int R = Cfg::R(s), LG = ilog2(R), BPT = Cfg::BPT(s);
        if constexpr (FS) apply_fs_twiddle(p);
        apply_scale(p);
        if constexpr (BLK) {
            const int lg = p.out_blk_log2;
            const int mask = (1 << lg) - 1;
            static_for<0, BPT>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                static_for<0, R>([&](auto kc) {
                    constexpr int k = decltype(kc)::value;
                    constexpr int q = i * R + brev(k, LG);
                    const int n = t + TPC * (i + k * BPT);
                    const int h = n >> lg;
                    const long long off = obase + (long long)(n & mask) * p.out_inner;
                    if constexpr (SPLIT) {
                        T xr, xi;
                        csplit(v[q], xr, xi);
                        p.out_blk0[h][off] = xr;
                        p.out_blk1[h][off] = xi;
                    } else {
                        st_c(reinterpret_cast<T2*>(p.out_blk0[h]) + off, v[q]);
                    }
                });
            });
            return;
        }
        const long long step = (long long)TPC * p.out_inner;
        if constexpr (SPLIT) {
            T* pr = p.out0 + obase + (long long)t * p.out_inner;
            T* pi = p.out1 + obase + (long long)t * p.out_inner;
            static_for<0, BPT>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                static_for<0, R>([&](auto kc) {
                    constexpr int k = decltype(kc)::value;
                    constexpr int q = i * R + brev(k, LG);
                    T xr, xi;
                    csplit(v[q], xr, xi);
                    pr[(i + k * BPT) * step] = xr;
                    pi[(i + k * BPT) * step] = xi;
                });
            });
        } else {
            T2* pc = reinterpret_cast<T2*>(p.out0) + obase + (long long)t * p.out_inner;
            static_for<0, BPT>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                static_for<0, R>([&](auto kc) {
                    constexpr int k = decltype(kc)::value;
                    constexpr int q = i * R + brev(k, LG);
                    if constexpr (POL) st_c_pol(pc + (i + k * BPT) * step, v[q], pol_out);
                    else st_c(pc + (i + k * BPT) * step, v[q]);
                });
            });
        }
    }
};

// ------------------------------------------------------------------ fused two-step strided program
// An axis of length N = N1*N2 whose elements are `inner` apart, W adjacent columns at a time (a "super-tile"
// of N x W elements, W*sizeof(complex) = 128 bytes per row), done by ONE CTA in one launch:
//   step A  for every n2: length-N1 FFT over n1 (rows n1*N2 + n2), times w_N^(k1*n2), stored to the CTA's scratch
//           slot as [n2][k1][W]              (sub-tiles of CfgA::G values of n2)
//   step B  for every k1: length-N2 FFT over n2 read back from the scratch slot, output row k1 + N1*k2
//           written to its final place       (sub-tiles of CfgB::G values of k1)
// i.e. the four-step decomposition of api.cu add_axis(), but with both steps inside one kernel so that every
// DRAM access is a full 128-byte line and the intermediate never leaves the L2 cache.  These two functions
// place a thread of step A / step B; load(), the stages and store() are the ordinary tile thread program.
template <class CfgA, class CfgB, class TH>
B2_HD void fused2_setup_a(TH& th, int tid, int c, long long in_base, long long inner_in, long long slot) {
    constexpr int W = CfgA::W;
    th.w = tid % W;
    th.t = (tid / W) % CfgA::TPC;
    th.g = tid / (W * CfgA::TPC);
    th.active = true;
    const long long n2 = (long long)c * CfgA::G + th.g;
    th.fs_n2i = n2;
    th.base = in_base + n2 * inner_in + th.w;              // element n1 at + n1 * (N2*inner_in)
    th.obase = slot + n2 * CfgA::N * W + th.w;             // scratch [n2][k1][W], k1 stride W
}
template <class CfgA, class CfgB, class TH>
B2_HD void fused2_setup_b(TH& th, int tid, int c, long long out_base, long long inner_out, long long slot) {
    constexpr int W = CfgB::W;
    th.w = tid % W;
    th.t = (tid / W) % CfgB::TPC;
    th.g = tid / (W * CfgB::TPC);
    th.active = true;
    const long long k1 = (long long)c * CfgB::G + th.g;
    th.fs_n2i = 0;
    th.base = slot + k1 * W + th.w;                        // element n2 at + n2 * (N1*W)
    th.obase = out_base + k1 * inner_out + th.w;           // output k = k1 + N1*k2 at + k2 * (N1*inner_out)
}

// ---- shared-memory-resident variant (kernels.cuh fused2s_fft_kernel): both steps are single-stage register FFTs
// (CfgA::S == CfgB::S == 1, one column per thread), so the CTA needs no exchange buffer and the intermediate of the
// super-tile itself lives in shared memory, laid out [k1][n2][W]; only the rows k1 >= KS that do not fit go through the
// (then tiny) global scratch slot, laid out [k1 - KS][n2][W].
template <class CfgA, class CfgB, int KS, class TH>
B2_HD void fused2s_store_a(const TH& th, vec2<typename CfgA::T>* smem_i, vec2<typename CfgA::T>* scratch_slot,
                           unsigned long long pol) {
    static_assert(CfgA::S == 1 && CfgA::TPC == 1, "step A is one register FFT per thread");
    constexpr int R = CfgA::E, LG = ilog2(R), W = CfgA::W, N2 = CfgB::N;
    const long long n2 = th.fs_n2i;
    static_for<0, R>([&](auto kc) {
        constexpr int k = decltype(kc)::value;             // k1
        constexpr int q = brev(k, LG);
        if constexpr (k < KS) st_c(smem_i + ((long long)k * N2 + n2) * W + th.w, th.v[q]);
        else st_c_pol(scratch_slot + ((long long)(k - KS) * N2 + n2) * W + th.w, th.v[q], pol);
    });
}
template <class CfgA, class CfgB, int KS, class TH>
B2_HD void fused2s_load_b(TH& th, int k1, const vec2<typename CfgA::T>* smem_i, const vec2<typename CfgA::T>* scratch_slot,
                          unsigned long long pol) {
    static_assert(CfgB::S == 1 && CfgB::TPC == 1, "step B is one register FFT per thread");
    constexpr int W = CfgA::W, N2 = CfgB::N;
    if (k1 < KS) {
        const vec2<typename CfgA::T>* src = smem_i + (long long)k1 * N2 * W + th.w;
        static_for<0, N2>([&](auto jc) { th.v[decltype(jc)::value] = ld_c(src + decltype(jc)::value * W); });
    } else {
        const vec2<typename CfgA::T>* src = scratch_slot + (long long)(k1 - KS) * N2 * W + th.w;
        static_for<0, N2>([&](auto jc) { th.v[decltype(jc)::value] = ld_stream_c_pol(src + decltype(jc)::value * W, pol); });
    }
}

// ---- in-place streamed variant (kernels.cuh fused2p_fft_kernel): the input rows n = n1*N2 + n2 with n1 < KS arrive in
// shared memory as the dense tile [n1][n2][W] (asynchronous 16-byte copies issued while the previous super-tile is in its
// step B), step A transforms column n2 IN PLACE (entry n1 is replaced by entry k1 of the same column, so [k1][n2][W] is
// what step B finds), the rows n1 >= KS that do not fit come straight from global memory one sub-tile ahead (`extra`).
template <class CfgA, class CfgB, int KS, class TH>
B2_HD void fused2p_load_a(TH& th, const vec2<typename CfgA::T>* smem_i, const cpx<typename CfgA::T>* extra) {
    static_assert(CfgA::S == 1 && CfgA::TPC == 1, "step A is one register FFT per thread");
    constexpr int W = CfgA::W, N2 = CfgB::N;
    const vec2<typename CfgA::T>* src = smem_i + th.fs_n2i * W + th.w;
    static_for<0, CfgA::N>([&](auto jc) {
        constexpr int j = decltype(jc)::value;                  // n1
        if constexpr (j < KS) th.v[j] = ld_c(src + (long long)j * N2 * W);
        else th.v[j] = extra[j - KS];
    });
}
// Chunk `idx` (16 bytes) of the staged part of a super-tile: row n = idx / CPP of the tile (CPP chunks per 128-byte row
// piece), returns the element offset inside the row piece; the source is row n of the tile in global memory, the
// destination row n of the dense shared-memory tile.
template <class CfgA>
struct Fused2PChunk {
    static constexpr int EPC = 16 / (2 * (int)sizeof(typename CfgA::T));     // complex elements per 16-byte chunk
    static constexpr int CPP = CfgA::W / EPC;                                // chunks per row piece
    static B2_HD long long row(long long idx) { return idx / CPP; }
    static B2_HD int elem(long long idx) { return (int)(idx % CPP) * EPC; }
};
// Rows of the shared-memory tile owned by one warp of a step-B sub-tile (thread tid reads row k1 = c*GB + tid/W, so warp v
// owns the RPW = 32/W rows c*GB + v*RPW + i) and the order in which the warp refills them: segment sg = the GA values of
// n2 of step-A sub-tile sg; piece p < RPW*GA of a segment is row k0 + p/GA, n2 = sg*GA + p%GA; a warp iteration
// covers PPI = 32/CPP pieces.
template <class CfgA, class CfgB>
struct Fused2PRows {
    static constexpr int RPW = 32 / CfgA::W;
    static constexpr int NSEG = CfgB::N / CfgA::G;
    static constexpr int PPI = 32 / Fused2PChunk<CfgA>::CPP;
    static constexpr int ITERS = RPW * CfgA::G / PPI;
    static_assert(32 % CfgA::W == 0 && (RPW * CfgA::G) % PPI == 0, "whole warp iterations per segment");
    static B2_HD int k_of(int k0, int p) { return k0 + p / CfgA::G; }
    static B2_HD int row(int k0, int sg, int p) { return k_of(k0, p) * CfgB::N + sg * CfgA::G + p % CfgA::G; }   // tile row n
};

// ------------------------------------------------------------------ short contiguous rows: 16-byte accesses + warp shuffles
// A row of N = 4 .. 64 complex64 elements is shared by N/2 lanes of one warp; lane l holds positions p = 2l, 2l + 1 (ONE
// 16-byte load per lane: every warp load covers whole 128-byte lines, where the tile program's 8-byte loads of a 4-thread
// column touch eight lines per instruction).  Radix-2 DIF over the positions: stage s (half = N >> (s+1) >= 2) pairs
// position p with p ^ half, i.e. lane l with lane l ^ (half/2), same register -- one shfl.sync.bfly per register instead
// of a shared-memory round trip; the lane that holds the upper element keeps (lower - upper) * w_{2 half}^(p mod half), the
// other one the sum.  The last stage pairs the lane's own two registers.  Position p then holds X[brev(p)]: lane l stores
// X[r] and X[r + N/2], r = brev_{log2N - 1}(l) -- per store instruction the lanes of a row cover one contiguous half row.
// The twiddles depend on the lane only, so they are set up once per thread (sincospi) and reused for every row it handles.
template <int LOG2N, bool INV>
struct ShflRow {
    using T = float;
    using C = cpx<T>;
    static_assert(LOG2N >= 2 && LOG2N <= 6, "N/2 lanes of one warp share a row");
    static constexpr int N = 1 << LOG2N, LP = N / 2, NST = LOG2N - 1;
    C tw[NST][2];
    bool upper[NST];
    B2_HD void init(int l) {
        static_for<0, NST>([&](auto sc) {
            constexpr int s = decltype(sc)::value;
            constexpr int half = N >> (s + 1);                 // >= 2, partner lane = l ^ (half / 2)
            upper[s] = (l & (half >> 1)) != 0;
            for (int b = 0; b < 2; ++b) {
                const int q = 2 * (l % (half / 2)) + b;       // position inside the lower half of the sub-transform
                T sn, cs;
#if defined(__CUDA_ARCH__)
                sincospif((T)q / (T)half, &sn, &cs);          // w_{2 half}^q = exp(-i pi q / half)
#else
                sn = (T)std::sin(3.14159265358979323846 * q / half);
                cs = (T)std::cos(3.14159265358979323846 * q / half);
#endif
                tw[s][b] = cmake<T>(cs, -sn);
            }
        });
    }
    B2_HD static C sel(bool c, const C& a, const C& b) {
#if defined(__CUDA_ARCH__)
        C r; r.v = c ? a.v : b.v; return r;
#else
        return c ? a : b;
#endif
    }
    // stage s: o = the partner lane's two registers
    template <int s>
    B2_HD void stage(C (&v)[2], const C (&o)[2]) const {
        for (int b = 0; b < 2; ++b) {
            const C up = cmul<INV>(csub(o[b], v[b]), tw[s][b]);
            const C lo = cadd(v[b], o[b]);
            v[b] = sel(upper[s], up, lo);
        }
    }
    B2_HD static void last(C (&v)[2]) {
        const C a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    }
    // output index of register 0 (register 1: + N/2)
    B2_HD static int out_index(int l) {
        int r = 0;
        for (int i = 0; i < NST; ++i) r |= ((l >> i) & 1) << (NST - 1 - i);
        return r;
    }
    B2_HD static void scale(C (&v)[2], T sc, int mode) {
        if (mode == 1) { v[0] = cscale(v[0], sc); v[1] = cscale(v[1], sc); }
        else if (mode == 2) {
            for (int b = 0; b < 2; ++b) { T xr, xi; csplit(v[b], xr, xi); v[b] = cmake<T>(xr / sc, xi / sc); }
        }
    }
};

// ---- the same with FOUR elements per lane and 16-byte stores as well: N/4 lanes share a row, lane l holds positions
// p = 2l + b + (N/2) h (b, h in {0, 1}): two 16-byte loads, one from each half of the row, both fully coalesced.  Stage 0
// (half = N/2) pairs the lane's own h = 0 / 1 registers (lane-dependent twiddle w_N^(2l+b)); the stages with half = N/4 .. 2
// pair lane l with lane l ^ (half/2) through shfl.sync.bfly (all four registers); the last stage (half = 1) pairs b = 0 / 1.
// Position p holds X[brev(p)] = X[h + 2 brev_{L-2}(l) + (N/2) b]: the lane's h = 0 / 1 outputs are ADJACENT, so the results
// leave as two 16-byte stores {X[2r], X[2r+1]} and {X[2r + N/2], X[2r+1 + N/2]}, r = brev_{L-2}(l) -- per instruction the
// lanes of a row cover one contiguous half row, like the loads.  One shuffle stage fewer than ShflRow for the same N.
template <int LOG2N, bool INV>
struct ShflRow4 {
    using T = float;
    using C = cpx<T>;
    static_assert(LOG2N >= 2 && LOG2N <= 7, "N/4 lanes of one warp share a row");
    static constexpr int N = 1 << LOG2N, LP = N / 4, NSX = LOG2N - 2;     // lanes per row, cross-lane stages
    C tw0[2];                         // stage 0: w_N^(2l + b)
    C tw[NSX > 0 ? NSX : 1][2];       // cross-lane stage s = 1 .. NSX: upper-output twiddle for b = 0, 1 (the same for h = 0, 1)
    bool upper[NSX > 0 ? NSX : 1];
    B2_HD static C root(int q, int len) {                      // exp(-2 pi i q / len)
        T sn, cs;
#if defined(__CUDA_ARCH__)
        sincospif((T)(2 * q) / (T)len, &sn, &cs);
#else
        sn = (T)std::sin(2.0 * 3.14159265358979323846 * q / len);
        cs = (T)std::cos(2.0 * 3.14159265358979323846 * q / len);
#endif
        return cmake<T>(cs, -sn);
    }
    B2_HD void init(int l) {
        for (int b = 0; b < 2; ++b) tw0[b] = root(2 * l + b, N);
        static_for<0, NSX>([&](auto sc) {
            constexpr int s = decltype(sc)::value + 1;         // stage index, half = N >> (s+1) in [2, N/4]
            constexpr int half = N >> (s + 1);
            upper[s - 1] = (l & (half >> 1)) != 0;
            for (int b = 0; b < 2; ++b) tw[s - 1][b] = root(2 * (l % (half / 2)) + b, 2 * half);
        });
    }
    B2_HD static C sel(bool c, const C& a, const C& b) {
#if defined(__CUDA_ARCH__)
        C r; r.v = c ? a.v : b.v; return r;
#else
        return c ? a : b;
#endif
    }
    // registers: v[2h + b]
    B2_HD void first(C (&v)[4]) const {
        for (int b = 0; b < 2; ++b) {
            const C lo = v[b], hi = v[2 + b];
            v[b] = cadd(lo, hi);
            v[2 + b] = cmul<INV>(csub(lo, hi), tw0[b]);
        }
    }
    template <int sx>                 // cross-lane stage sx = 0 .. NSX-1; o = the partner lane's four registers
    B2_HD void stage(C (&v)[4], const C (&o)[4]) const {
        for (int r = 0; r < 4; ++r) {
            const C up = cmul<INV>(csub(o[r], v[r]), tw[sx][r & 1]);
            const C lo = cadd(v[r], o[r]);
            v[r] = sel(upper[sx], up, lo);
        }
    }
    B2_HD static void last(C (&v)[4]) {
        for (int h = 0; h < 2; ++h) {
            const C a = v[2 * h], b = v[2 * h + 1];
            v[2 * h] = cadd(a, b);
            v[2 * h + 1] = csub(a, b);
        }
    }
    // v[2h + b] holds X[h + 2 r + (N/2) b], r = out_index(l)
    B2_HD static int out_index(int l) {
        int r = 0;
        for (int i = 0; i < NSX; ++i) r |= ((l >> i) & 1) << (NSX - 1 - i);
        return r;
    }
    B2_HD static void scale(C (&v)[4], T sc, int mode) {
        if (mode == 1) { for (int r = 0; r < 4; ++r) v[r] = cscale(v[r], sc); }
        else if (mode == 2) {
            for (int r = 0; r < 4; ++r) { T xr, xi; csplit(v[r], xr, xi); v[r] = cmake<T>(xr / sc, xi / sc); }
        }
    }
};

// ------------------------------------------------------------------ lane-pair FFT (warp-shuffle exchange)
// A length-N = 2E transform shared by two lanes of a warp (t = 0, 1): lane t holds x[2j + t] in v[j], does the E-point
// FFT of its residue class in registers, lane 1 applies w_N^ka, and the final radix-2 stage pairs Z_0[ka] with Z_1[ka]:
//     X[ka + E*kb] = Z_0[ka] + (-1)^kb Z_1[ka].
// Each lane computes the pairs of half of the ka values; the halves it does not own go to the partner in ONE exchange of
// E/2 complex values -- a warp shuffle on the device (shfl.sync.bfly, lane ^ 16), a plain swap in the host emulation --
// instead of a shared-memory round trip.  Register r holds ka = brev(r), so after post():
//     v[i]       = X[brev(i, L) + t]            i in [0, E/2)
//     v[E/2 + i] = X[brev(i, L) + t + E]
// i.e. the lane's outputs are a compile-time index plus t, which folds into its base pointers.
template <typename T, int LOG2N, bool INV>
struct PairFFT {
    static constexpr int N = 1 << LOG2N, E = N / 2, L = LOG2N - 1, H = E / 2;
    static_assert(LOG2N >= 2 && LOG2N <= 6, "the pair twiddles are 64th roots of unity");
    using C = cpx<T>;
    static constexpr int out_k(int i) { return brev(i, L); }       // + t (+ E for the upper half)
    B2_HD static C sel(bool c, const C& a, const C& b) {
#if defined(__CUDA_ARCH__)
        if constexpr (is_packed<T>::value) { C r; r.v = c ? a.v : b.v; return r; } else
#endif
        { return c ? a : b; }
    }
    B2_HD static void pre(C* v, int t, C* send) {
        butterfly<E, 0, INV>(v);
        static_for<1, E>([&](auto rc) {
            constexpr int r = decltype(rc)::value;
            constexpr int ka = brev(r, L);
            v[r] = sel(t != 0, cmulc<ka * (64 / N), INV>(v[r]), v[r]);
        });
        static_for<0, H>([&](auto ic) { constexpr int i = decltype(ic)::value; send[i] = sel(t != 0, v[i], v[H + i]); });
    }
    B2_HD static void post(C* v, int t, const C* recv) {
        const T sgn = t ? (T)-1 : (T)1;
        static_for<0, H>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            const C own = sel(t != 0, v[H + i], v[i]);
            v[i] = cadd(own, recv[i]);
            v[H + i] = cscale(csub(own, recv[i]), sgn);
        });
    }
#if defined(__CUDA_ARCH__)
    // the exchange itself: partner = lane ^ 16 (the two lanes of a pair sit in the two half-warps, one column each)
    __device__ __forceinline__ static void exchange(const C* send, C* recv) {
        static_for<0, H>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            if constexpr (is_packed<T>::value) {
                recv[i].v = __shfl_xor_sync(0xffffffffu, send[i].v, 16);
            } else {
                recv[i].x = __shfl_xor_sync(0xffffffffu, send[i].x, 16);
                recv[i].y = __shfl_xor_sync(0xffffffffu, send[i].y, 16);
            }
        });
    }
#endif
};

// ------------------------------------------------------------------ fused two-step strided program, lane-pair steps
// The thread program of kernels.cuh fused2w_fft_kernel: N = N1*N2 as in fused2s (intermediate [k1][n2][W] in shared
// memory, rows k1 >= KS in the global scratch slot), but every step is a PairFFT: a length-N1 (N2) transform is shared by
// the two lanes (w, t = 0 / 1) of a warp that handle the same column, each holding N1/2 (N2/2) elements, with the radix-2
// stage exchanged by warp shuffle.  Twice the threads per super-tile of the register-FFT version at half the registers.
template <int LOG2A, int LOG2B, int W_, int KS, bool INV>
struct Fused2W {
    using T = float;
    using C = cpx<T>;
    using T2 = vec2<T>;
    using PA = PairFFT<T, LOG2A, INV>;
    using PB = PairFFT<T, LOG2B, INV>;
    static constexpr int W = W_, N1 = PA::N, N2 = PB::N, EA = PA::E, EB = PB::E, HA = PA::H, HB = PB::H;
    static_assert(KS % 2 == 0 && KS <= N1, "lane pairs (k1, k1 + 1) stay on one side of the shared/global split");

    // step A input: x[n1][n2][w] with n1 = 2j + t; `col` points at (n1 = 0, this n2, this w), rows are `stride` apart
    B2_HD static void a_load(C* v, const T2* col, long long stride, int t, unsigned long long pol) {
        static_for<0, EA>([&](auto jc) { constexpr int j = decltype(jc)::value; v[j] = ld_stream_c_pol(col + (long long)(2 * j + t) * stride, pol); });
    }
    // step A output k1 = out_k(i) + t (+ EA): times w_N^(k1*n2) (table [k1][N2]), into the intermediate
    B2_HD static void a_store(const C* v, int t, long long n2, int w, const T2* fs_tab, T2* smem_i, T2* scratch_slot,
                              unsigned long long pol) {
        static_for<0, 2>([&](auto kbc) {
            constexpr int kb = decltype(kbc)::value;
            static_for<0, HA>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                constexpr int k0 = PA::out_k(i) + EA * kb;                 // k1 = k0 + t
                const C val = cmul<INV>(v[i + HA * kb], ld_c(fs_tab + (long long)(k0 + t) * N2 + n2));
                if constexpr (k0 < KS) st_c(smem_i + ((long long)(k0 + t) * N2 + n2) * W + w, val);
                else st_c_pol(scratch_slot + ((long long)(k0 + t - KS) * N2 + n2) * W + w, val, pol);
            });
        });
    }
    // step B input: intermediate row k1, n2 = 2j + t
    B2_HD static void b_load(C* v, int t, int k1, int w, const T2* smem_i, const T2* scratch_slot, unsigned long long pol) {
        if (k1 < KS) {
            const T2* src = smem_i + ((long long)k1 * N2 + t) * W + w;
            static_for<0, EB>([&](auto jc) { constexpr int j = decltype(jc)::value; v[j] = ld_c(src + 2 * j * W); });
        } else {
            const T2* src = scratch_slot + ((long long)(k1 - KS) * N2 + t) * W + w;
            static_for<0, EB>([&](auto jc) { constexpr int j = decltype(jc)::value; v[j] = ld_stream_c_pol(src + 2 * j * W, pol); });
        }
    }
    // step B output k2 = out_k(i) + t (+ EB) -> row k1 + N1*k2 of the result; `col` points at (k = k1, this w), k2 rows are
    // `stride` (= N1 * inner) apart; scale as TileThread::apply_scale (pyfft/kernel.py:23-37)
    B2_HD static void b_store(const C* v, int t, T2* col, long long stride, T scale, int scale_mode, unsigned long long pol) {
        static_for<0, 2>([&](auto kbc) {
            constexpr int kb = decltype(kbc)::value;
            static_for<0, HB>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                constexpr int k0 = PB::out_k(i) + EB * kb;
                C val = v[i + HB * kb];
                if (scale_mode == 1) val = cscale(val, scale);
                else if (scale_mode == 2) { T xr, xi; csplit(val, xr, xi); val = cmake<T>(xr / scale, xi / scale); }
                st_c_pol(col + (long long)(k0 + t) * stride, val, pol);
            });
        });
    }
};

}  // namespace b2
