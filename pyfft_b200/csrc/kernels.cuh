// __global__ wrappers around the tile thread program + the variant registry.
#pragma once

#include <cuda_runtime.h>

#include <vector>

#include "fft_core.cuh"

namespace b2 {

template <class Cfg, bool SPLIT, bool INV, int s>
__device__ __forceinline__ void run_stages(TileThread<Cfg, SPLIT, INV>& th, const PassParams<typename Cfg::T>& p,
                                           vec2<typename Cfg::T>* smem) {
    th.template compute<s>(p);
    if constexpr (s + 1 < Cfg::S) {
        if constexpr (s > 0) __syncthreads();   // previous exchange fully read before it is overwritten
        th.template xwrite<s>(smem);
        __syncthreads();
        th.template xread<s>(smem);
        run_stages<Cfg, SPLIT, INV, s + 1>(th, p, smem);
    }
}

template <class Cfg, bool SPLIT, bool INV, int MINB, bool BLK = false>
__global__ void __launch_bounds__(Cfg::THREADS, MINB)
tile_fft_kernel(const __grid_constant__ PassParams<typename Cfg::T> p) {
    extern __shared__ __align__(16) unsigned char b2_smem_raw[];
    auto* smem = reinterpret_cast<vec2<typename Cfg::T>*>(b2_smem_raw);
    TileThread<Cfg, SPLIT, INV> th;
    th.setup((int)threadIdx.x, (long long)blockIdx.x, p);
    th.load(p);
    run_stages<Cfg, SPLIT, INV, 0>(th, p, smem);
    th.template store<BLK>(p);
    if constexpr (BLK) __threadfence_system();   // peer (NVLink) stores visible before the kernel retires
}

// ------------------------------------------------------------------ persistent TMA-fed kernel (contiguous axis)
// One CTA per resident slot loops over groups of G transforms.  The next group's input is
// fetched by a single cp.async.bulk (TMA 1D bulk copy, SASS UBLKCP) into a shared-memory ring
// while the current group is in registers, so HBM reads stay in flight for the whole compute
// phase instead of only during each CTA's load phase (the plain kernel is latency-bound at
// 3-4 CTAs/SM, see profiles/).  Stage 0 then reads its elements from shared memory.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}

template <class Cfg, int NBUF>
struct TmaRowLayout {
    using T = typename Cfg::T;
    static constexpr size_t IN_BYTES = (size_t)Cfg::G * Cfg::N * 2 * sizeof(T);          // one ring slot (re+im)
    static constexpr size_t X_OFF = NBUF * IN_BYTES;
    static constexpr size_t X_BYTES = ((size_t)Cfg::SMEM_BYTES + 127) / 128 * 128;
    static constexpr size_t BAR_OFF = X_OFF + X_BYTES;         // full[NBUF] then empty[NBUF]
    static constexpr size_t TOTAL = BAR_OFF + 16 * NBUF;
};

template <class Cfg, bool SPLIT, bool INV, int MINB, int NBUF>
__global__ void __launch_bounds__(Cfg::THREADS, MINB)
tile_fft_kernel_tma_row(const __grid_constant__ PassParams<typename Cfg::T> p) {
    using T = typename Cfg::T;
    using T2 = vec2<T>;
    using L = TmaRowLayout<Cfg, NBUF>;
    static_assert(Cfg::W == 1, "bulk-copy staging is for the contiguous axis");
    extern __shared__ __align__(128) unsigned char b2_smem_raw[];
    T2* xbuf = reinterpret_cast<T2*>(b2_smem_raw + L::X_OFF);
    uint64_t* bars = reinterpret_cast<uint64_t*>(b2_smem_raw + L::BAR_OFF);      // "slot full": TMA bytes landed
    uint64_t* empty = bars + NBUF;                                              // "slot empty": every thread has read it
    const int tid = (int)threadIdx.x;
    const long long n_groups = (p.n_tiles + Cfg::G - 1) / Cfg::G;
    const long long stride = gridDim.x;

    // one elected thread fetches group `grp` into ring slot `slot`
    auto issue = [&](long long grp, int slot) {
        if (grp >= n_groups) return;
        long long tiles = p.n_tiles - grp * Cfg::G;
        if (tiles > Cfg::G) tiles = Cfg::G;
        const long long first = grp * Cfg::G * Cfg::N;                     // element offset of the group
        unsigned char* dst = b2_smem_raw + (size_t)slot * L::IN_BYTES;
        if constexpr (SPLIT) {
            const uint32_t bytes = (uint32_t)(tiles * Cfg::N * sizeof(T));
            mbar_expect_tx(&bars[slot], 2 * bytes);
            bulk_load(dst, p.in0 + first, bytes, &bars[slot]);
            bulk_load(dst + L::IN_BYTES / 2, p.in1 + first, bytes, &bars[slot]);
        } else {
            const uint32_t bytes = (uint32_t)(tiles * Cfg::N * 2 * sizeof(T));
            mbar_expect_tx(&bars[slot], bytes);
            bulk_load(dst, reinterpret_cast<const T2*>(p.in0) + first, bytes, &bars[slot]);
        }
    };

    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < NBUF; ++b) { mbar_init(&bars[b], 1); mbar_init(&empty[b], Cfg::THREADS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    long long grp = blockIdx.x;
    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < NBUF; ++b) issue(grp + b * stride, b);
    }
    TileThread<Cfg, SPLIT, INV> th;
    for (unsigned it = 0; grp < n_groups; ++it, grp += stride) {
        const int slot = (int)(it % NBUF);
        th.setup(tid, grp, p);
        mbar_wait(&bars[slot], (it / NBUF) & 1);
        const unsigned char* src = b2_smem_raw + (size_t)slot * L::IN_BYTES;
        th.load_smem(src, src + L::IN_BYTES / 2);
        // Release the slot to the async proxy (the TMA refill).  bar.sync alone is NOT enough: it orders
        // generic-proxy accesses only, and shared-memory loads that are still queued in the LSU were
        // observed to read the refilled data (tools/stress_variants.py).  Each thread therefore fences its
        // generic-proxy reads against the async proxy, then arrives on the slot's "empty" mbarrier; the
        // elected thread waits for that phase before it issues the bulk copy.
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(&empty[slot]);
        __syncthreads();                       // the previous exchange buffer contents are fully read
        if (tid == 0 && grp + (long long)NBUF * stride < n_groups) {
            mbar_wait(&empty[slot], (it / NBUF) & 1);
            issue(grp + (long long)NBUF * stride, slot);
        }
        run_stages<Cfg, SPLIT, INV, 0>(th, p, xbuf);
        th.store(p);
    }
}

// ------------------------------------------------------------------ registry
struct KernelVariant {
    const char* name;
    int prec;        // 0 = f32, 1 = f64
    int log2n;
    int W, G, E, S;
    int radix[4];
    int threads;
    long long smem_bytes;
    int minb;
    int kind;        // 0 = direct global loads, 1 = persistent + TMA bulk staging (contiguous axis only)
    int nbuf;        // ring depth for kind 1
    int blk;         // 1: also compiled with destination-blocked stores (slab exchange passes)
    // launches ceil(n_tiles / G) CTAs; params points at a PassParams<T> of the right T
    cudaError_t (*launch)(int split, int inv, const void* params, cudaStream_t stream);
    cudaError_t (*prepare)();   // one-time function attributes (dynamic smem opt-in)
    // occupancy (CTAs/SM) of the interleaved forward kernel, for the tuning report
    int (*occupancy)();
};

template <class Cfg, int MINB, bool BLKCAP = false>
struct VariantOps {
    using T = typename Cfg::T;
    static cudaError_t prepare() {
        cudaError_t e = cudaSuccess;
        if (Cfg::SMEM_BYTES > 48 * 1024) {
            const int b = (int)Cfg::SMEM_BYTES;
            if constexpr (BLKCAP) {
                e = cudaFuncSetAttribute(tile_fft_kernel<Cfg, false, false, MINB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, b);
                if (e != cudaSuccess) return e;
                e = cudaFuncSetAttribute(tile_fft_kernel<Cfg, false, true, MINB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, b);
                if (e != cudaSuccess) return e;
                e = cudaFuncSetAttribute(tile_fft_kernel<Cfg, true, false, MINB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, b);
                if (e != cudaSuccess) return e;
            }
            e = cudaFuncSetAttribute(tile_fft_kernel<Cfg, false, false, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, b);
            if (e != cudaSuccess) return e;
            e = cudaFuncSetAttribute(tile_fft_kernel<Cfg, false, true, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, b);
            if (e != cudaSuccess) return e;
            e = cudaFuncSetAttribute(tile_fft_kernel<Cfg, true, false, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, b);
        }
        return e;
    }
    static cudaError_t launch(int split, int inv, const void* params, cudaStream_t stream) {
        const PassParams<T>& p = *static_cast<const PassParams<T>*>(params);
        const long long ctas = (p.n_tiles + Cfg::G - 1) / Cfg::G;
        if (ctas <= 0) return cudaSuccess;
        if (ctas > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
        const dim3 grid((unsigned)ctas), block(Cfg::THREADS);
        const size_t sm = (size_t)Cfg::SMEM_BYTES;
        if (p.out_blk_log2 >= 0) {
            if constexpr (BLKCAP) {
                if (split) tile_fft_kernel<Cfg, true, false, MINB, true><<<grid, block, sm, stream>>>(p);
                else if (inv) tile_fft_kernel<Cfg, false, true, MINB, true><<<grid, block, sm, stream>>>(p);
                else tile_fft_kernel<Cfg, false, false, MINB, true><<<grid, block, sm, stream>>>(p);
                return cudaGetLastError();
            } else {
                return cudaErrorNotSupported;
            }
        }
        if (split) tile_fft_kernel<Cfg, true, false, MINB><<<grid, block, sm, stream>>>(p);
        else if (inv) tile_fft_kernel<Cfg, false, true, MINB><<<grid, block, sm, stream>>>(p);
        else tile_fft_kernel<Cfg, false, false, MINB><<<grid, block, sm, stream>>>(p);
        return cudaGetLastError();
    }
    static int occupancy() {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, tile_fft_kernel<Cfg, false, false, MINB>, Cfg::THREADS,
                                                          (size_t)Cfg::SMEM_BYTES) != cudaSuccess)
            return -1;
        return n;
    }
    static KernelVariant make(const char* name) {
        KernelVariant v{};
        v.name = name;
        v.prec = sizeof(T) == 4 ? 0 : 1;
        v.log2n = Cfg::LOG2N;
        v.W = Cfg::W; v.G = Cfg::G; v.E = Cfg::E; v.S = Cfg::S;
        for (int s = 0; s < 4; ++s) v.radix[s] = s < Cfg::S ? Cfg::R(s) : 1;
        v.threads = Cfg::THREADS;
        v.smem_bytes = Cfg::SMEM_BYTES;
        v.minb = MINB;
        v.blk = BLKCAP ? 1 : 0;
        v.launch = &launch;
        v.prepare = &prepare;
        v.occupancy = &occupancy;
        return v;
    }
};

template <class Cfg, int MINB, int NBUF>
struct VariantOpsTma {
    using T = typename Cfg::T;
    using L = TmaRowLayout<Cfg, NBUF>;
    static int& slots() { static int s = 0; return s; }   // resident CTAs on the device (SMs x occupancy)
    static cudaError_t prepare() {
        const int b = (int)L::TOTAL;
        cudaError_t e = cudaFuncSetAttribute(tile_fft_kernel_tma_row<Cfg, false, false, MINB, NBUF>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, b);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(tile_fft_kernel_tma_row<Cfg, false, true, MINB, NBUF>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, b);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(tile_fft_kernel_tma_row<Cfg, true, false, MINB, NBUF>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, b);
        if (e != cudaSuccess) return e;
        int dev = 0, sms = 0, occ = 0;
        e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tile_fft_kernel_tma_row<Cfg, false, false, MINB, NBUF>,
                                                          Cfg::THREADS, L::TOTAL);
        if (e != cudaSuccess) return e;
        slots() = sms * (occ > 0 ? occ : 1);
        return cudaSuccess;
    }
    static cudaError_t launch(int split, int inv, const void* params, cudaStream_t stream) {
        const PassParams<T>& p = *static_cast<const PassParams<T>*>(params);
        long long ctas = (p.n_tiles + Cfg::G - 1) / Cfg::G;
        if (ctas <= 0) return cudaSuccess;
        if (slots() <= 0) { cudaError_t e = prepare(); if (e != cudaSuccess) return e; }
        if (ctas > slots()) ctas = slots();
        const dim3 grid((unsigned)ctas), block(Cfg::THREADS);
        const size_t sm = L::TOTAL;
        if (split) tile_fft_kernel_tma_row<Cfg, true, false, MINB, NBUF><<<grid, block, sm, stream>>>(p);
        else if (inv) tile_fft_kernel_tma_row<Cfg, false, true, MINB, NBUF><<<grid, block, sm, stream>>>(p);
        else tile_fft_kernel_tma_row<Cfg, false, false, MINB, NBUF><<<grid, block, sm, stream>>>(p);
        return cudaGetLastError();
    }
    static int occupancy() {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, tile_fft_kernel_tma_row<Cfg, false, false, MINB, NBUF>,
                                                          Cfg::THREADS, L::TOTAL) != cudaSuccess)
            return -1;
        return n;
    }
    static KernelVariant make(const char* name) {
        KernelVariant v = VariantOps<Cfg, MINB, false>::make(name);
        v.smem_bytes = (long long)L::TOTAL;
        v.kind = 1;
        v.nbuf = NBUF;
        v.launch = &launch;
        v.prepare = &prepare;
        v.occupancy = &occupancy;
        return v;
    }
};

// B2_V(type, log2n, W, G, minblocks, R0, R1, R2, R3)
#define B2_STR2(x) #x
#define B2_STR(x) B2_STR2(x)
#define B2_V(T, L, W, G, MB, R0, R1, R2, R3)                                                         \
    out.push_back(::b2::VariantOps<::b2::TileCfg<T, L, W, G, R0, R1, R2, R3>, MB>::make(               \
        #T "_n" B2_STR(L) "_w" B2_STR(W) "_g" B2_STR(G) "_b" B2_STR(MB) "_r" B2_STR(R0) "x" B2_STR(R1) "x" B2_STR(R2) "x" B2_STR(R3)));

// B2_VB: like B2_V, additionally compiled with destination-blocked stores (used for the strided-axis defaults)
#define B2_VB(T, L, W, G, MB, R0, R1, R2, R3)                                                        \
    out.push_back(::b2::VariantOps<::b2::TileCfg<T, L, W, G, R0, R1, R2, R3>, MB, true>::make(         \
        #T "_n" B2_STR(L) "_w" B2_STR(W) "_g" B2_STR(G) "_b" B2_STR(MB) "_r" B2_STR(R0) "x" B2_STR(R1) "x" B2_STR(R2) "x" B2_STR(R3)));

// B2_VT(type, log2n, G, minblocks, ring depth, R0, R1, R2, R3): persistent TMA-staged contiguous-axis variant
#define B2_VT(T, L, G, MB, NB, R0, R1, R2, R3)                                                       \
    out.push_back(::b2::VariantOpsTma<::b2::TileCfg<T, L, 1, G, R0, R1, R2, R3>, MB, NB>::make(        \
        #T "_n" B2_STR(L) "_w1_g" B2_STR(G) "_b" B2_STR(MB) "_r" B2_STR(R0) "x" B2_STR(R1) "x" B2_STR(R2) "x" B2_STR(R3) "_tma" B2_STR(NB)));

void register_f32_row(std::vector<KernelVariant>& out);
void register_f32_col(std::vector<KernelVariant>& out);
void register_f64_row(std::vector<KernelVariant>& out);
void register_f64_col(std::vector<KernelVariant>& out);
void register_exp(std::vector<KernelVariant>& out);

}  // namespace b2
