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

template <class Cfg, bool SPLIT, bool INV, int MINB>
__global__ void __launch_bounds__(Cfg::THREADS, MINB)
tile_fft_kernel(const __grid_constant__ PassParams<typename Cfg::T> p) {
    extern __shared__ __align__(16) unsigned char b2_smem_raw[];
    auto* smem = reinterpret_cast<vec2<typename Cfg::T>*>(b2_smem_raw);
    TileThread<Cfg, SPLIT, INV> th;
    th.setup((int)threadIdx.x, (long long)blockIdx.x, p);
    th.load(p);
    run_stages<Cfg, SPLIT, INV, 0>(th, p, smem);
    th.store(p);
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
    // launches ceil(n_tiles / G) CTAs; params points at a PassParams<T> of the right T
    cudaError_t (*launch)(int split, int inv, const void* params, cudaStream_t stream);
    cudaError_t (*prepare)();   // one-time function attributes (dynamic smem opt-in)
    // occupancy (CTAs/SM) of the interleaved forward kernel, for the tuning report
    int (*occupancy)();
};

template <class Cfg, int MINB>
struct VariantOps {
    using T = typename Cfg::T;
    static cudaError_t prepare() {
        cudaError_t e = cudaSuccess;
        if (Cfg::SMEM_BYTES > 48 * 1024) {
            const int b = (int)Cfg::SMEM_BYTES;
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

void register_f32_row(std::vector<KernelVariant>& out);
void register_f32_col(std::vector<KernelVariant>& out);
void register_f64_row(std::vector<KernelVariant>& out);
void register_f64_col(std::vector<KernelVariant>& out);

}  // namespace b2
