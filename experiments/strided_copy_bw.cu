// Microbenchmark behind the strided-axis kernel design: what can the memory system
// deliver when a pass reads and writes tiles of [ROWS rows][W columns] whose rows are `pitch` apart
// (W*8-byte pieces at 2 KiB .. 32 MiB pitch), with no arithmetic at all?  It is the ceiling for the Y and Z
// passes of a 3-D transform as a function of piece width, pitch, and of who issues the accesses:
//   lsu: every thread moves 8-byte elements with ld/st.global, 16 in flight per thread (the plain kernels)
//   tma: one thread per CTA moves whole tiles with cp.async.bulk.tensor loads and stores through a
//        shared-memory ring (the persistent kernels), with the tensor map's L2 promotion as a parameter
// Not part of the product.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o strided_copy_bw strided_copy_bw.cu -lcuda
//   ./strided_copy_bw
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) { std::printf("%s: %s\n", #x, cudaGetErrorString(e_)); std::exit(1); } \
    } while (0)

struct Geo {
    long long pitch;         // elements (float2) between consecutive rows
    long long outer_stride;  // elements between consecutive outer blocks
    int nrows;               // rows per outer block (the transformed axis)
    int col_blocks;          // tiles along the contiguous dimension
    int row_blocks;          // nrows / ROWS
    long long n_tiles;
};

template <int W, int ROWS, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) copy_lsu(float2* __restrict__ a, Geo g) {
    constexpr int TPC = THREADS / W, E = ROWS / TPC;
    const int w = threadIdx.x % W, t = threadIdx.x / W;
    for (long long tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
        const long long ib = tile % g.col_blocks, r = tile / g.col_blocks;
        const long long rb = r % g.row_blocks, o = r / g.row_blocks;
        float2* base = a + o * g.outer_stride + (rb * ROWS + t) * g.pitch + ib * W + w;
        unsigned long long v[E];
#pragma unroll
        for (int j = 0; j < E; ++j)
            asm volatile("ld.global.L1::no_allocate.b64 %0, [%1];" : "=l"(v[j]) : "l"(base + (long long)j * TPC * g.pitch));
#pragma unroll
        for (int j = 0; j < E; ++j)
            asm volatile("st.global.b64 [%0], %1;" ::"l"(base + (long long)j * TPC * g.pitch), "l"(v[j] ^ 0x80000000ull) : "memory");
    }
}

// MODE 0: copy, 1: loads only (xor-reduced into one dummy store per thread), 2: stores only.
// CS > 1: the CTAs of a cluster take CS adjacent tiles and re-align with a cluster barrier before the load and the
// store phase, so the CS neighbouring W*8-byte pieces of one row are requested within a few hundred ns of each other.
template <int W, int ROWS, int THREADS, int MINB, int MODE, int CS>
__global__ void __launch_bounds__(THREADS, MINB) copy_lsu2(float2* __restrict__ a, Geo g, unsigned long long* sink) {
    constexpr int TPC = THREADS / W, E = ROWS / TPC;
    const int w = threadIdx.x % W, t = threadIdx.x / W;
    const long long tile = blockIdx.x;
    const long long ib = tile % g.col_blocks, r = tile / g.col_blocks;
    const long long rb = r % g.row_blocks, o = r / g.row_blocks;
    float2* base = a + o * g.outer_stride + (rb * ROWS + t) * g.pitch + ib * W + w;
    unsigned long long v[E];
    if constexpr (CS > 1) { asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory"); }
    if constexpr (MODE != 2) {
#pragma unroll
        for (int j = 0; j < E; ++j)
            asm volatile("ld.global.L1::no_allocate.b64 %0, [%1];" : "=l"(v[j]) : "l"(base + (long long)j * TPC * g.pitch));
    } else {
#pragma unroll
        for (int j = 0; j < E; ++j) v[j] = (unsigned long long)(tile + j);
    }
    if constexpr (CS > 1) {
        unsigned long long x = 0;
#pragma unroll
        for (int j = 0; j < E; ++j) x ^= v[j];
        if (x == 0x123456789abcdefull) sink[0] = x;      // wait for the loads before the barrier
        asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
    }
    if constexpr (MODE != 1) {
#pragma unroll
        for (int j = 0; j < E; ++j)
            asm volatile("st.global.b64 [%0], %1;" ::"l"(base + (long long)j * TPC * g.pitch), "l"(v[j] ^ 0x80000000ull) : "memory");
    } else {
        unsigned long long x = 0;
#pragma unroll
        for (int j = 0; j < E; ++j) x ^= v[j];
        if (x == 0x123456789abcdefull) sink[0] = x;
    }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int W, int ROWS, int S, int D>
__global__ void __launch_bounds__(32, 1) copy_tma(const __grid_constant__ CUtensorMap tm, Geo g) {
    extern __shared__ __align__(128) unsigned char sm[];
    constexpr int NB = ROWS < 256 ? ROWS : 256, NLOAD = ROWS / NB;
    constexpr uint32_t SLOT = (uint32_t)ROWS * W * 8;
    __shared__ __align__(8) uint64_t full[S];
    if (threadIdx.x != 0) return;
    for (int s = 0; s < S; ++s)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    long long n_mine = 0;
    for (long long tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) ++n_mine;
    auto coords = [&](long long i, int& c0, int& c1, int& c2) {
        const long long tile = blockIdx.x + i * gridDim.x;
        const long long ib = tile % g.col_blocks, r = tile / g.col_blocks;
        const long long rb = r % g.row_blocks, o = r / g.row_blocks;
        c0 = (int)(ib * W * 2); c1 = (int)(rb * ROWS); c2 = (int)o;
    };
    for (long long it = 0; it < n_mine + D; ++it) {
        if (it < n_mine) {
            const int slot = (int)(it % S);
            if (it >= S) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(S - D - 1) : "memory");
            int c0, c1, c2;
            coords(it, c0, c1, c2);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[slot])), "r"(SLOT) : "memory");
#pragma unroll
            for (int nb = 0; nb < NLOAD; ++nb)
                asm volatile(
                    "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                        smem_u32(sm + (size_t)slot * SLOT + (size_t)nb * NB * W * 8)),
                    "l"(&tm), "r"(c0), "r"(c1 + nb * NB), "r"(c2), "r"(smem_u32(&full[slot]))
                    : "memory");
        }
        const long long j = it - D;
        if (j >= 0) {
            const int slot = (int)(j % S);
            const uint32_t parity = (uint32_t)((j / S) & 1);
            uint32_t ok;
            do {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(smem_u32(&full[slot])), "r"(parity) : "memory");
            } while (!ok);
            int c0, c1, c2;
            coords(j, c0, c1, c2);
#pragma unroll
            for (int nb = 0; nb < NLOAD; ++nb)
                asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(&tm),
                             "r"(smem_u32(sm + (size_t)slot * SLOT + (size_t)nb * NB * W * 8)), "r"(c0), "r"(c1 + nb * NB), "r"(c2)
                             : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_fn get_encode() {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
    return (encode_fn)f;
}

static int g_sms = 148;
static float2* g_buf = nullptr;

struct Case { const char* name; long long pitch; int nrows; long long outer; long long cols; };

template <class F>
static double time_ms(F&& launch, int reps = 5) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

static Geo make_geo(const Case& c, int W, int ROWS) {
    Geo g;
    g.pitch = c.pitch; g.nrows = c.nrows; g.outer_stride = (long long)c.nrows * c.pitch;
    g.col_blocks = (int)(c.cols / W); g.row_blocks = c.nrows / ROWS;
    g.n_tiles = (long long)g.col_blocks * g.row_blocks * c.outer;
    return g;
}

template <int W, int ROWS, int THREADS, int MINB>
static void run_lsu(const Case& c) {
    if (ROWS > c.nrows) return;
    const Geo g = make_geo(c, W, ROWS);
    const double bytes = 2.0 * 8.0 * c.cols * c.nrows * c.outer;
    for (int cap : {0, 2}) {   // 0: one CTA per tile; else persistent grid of cap CTAs per SM
        const long long grid = cap ? (long long)cap * g_sms : g.n_tiles;
        const double ms = time_ms([&] { copy_lsu<W, ROWS, THREADS, MINB><<<(unsigned)grid, THREADS>>>(g_buf, g); });
        std::printf("%-14s lsu  W=%-2d piece=%-4dB rows/tile=%-4d thr=%-4d minb=%d grid=%-9s : %8.3f ms  %7.1f GB/s\n", c.name, W, W * 8,
                    ROWS, THREADS, MINB, cap ? "2/SM" : "tiles", ms, bytes / ms * 1e-6);
    }
}

template <int W, int ROWS, int THREADS, int MINB, int MODE, int CS>
static void run_lsu2(const Case& c) {
    if (ROWS > c.nrows) return;
    const Geo g = make_geo(c, W, ROWS);
    if (g.n_tiles % CS) return;
    static unsigned long long* sink = nullptr;
    if (!sink) CK(cudaMalloc(&sink, 64));
    const double bytes = (MODE == 0 ? 2.0 : 1.0) * 8.0 * c.cols * c.nrows * c.outer;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)g.n_tiles); cfg.blockDim = dim3(THREADS);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    const double ms = time_ms([&] { CK(cudaLaunchKernelEx(&cfg, copy_lsu2<W, ROWS, THREADS, MINB, MODE, CS>, g_buf, g, sink)); });
    std::printf("%-14s lsu2 W=%-2d piece=%-4dB rows/tile=%-4d thr=%-4d minb=%d %s cluster=%d : %8.3f ms  %7.1f GB/s\n", c.name, W, W * 8,
                ROWS, THREADS, MINB, MODE == 0 ? "copy " : MODE == 1 ? "loads" : "store", CS, ms, bytes / ms * 1e-6);
}

template <int W, int ROWS, int S, int D>
static void run_tma(const Case& c, int ctas_per_sm, CUtensorMapL2promotion promo, const char* pname) {
    static encode_fn enc = get_encode();
    if (ROWS > c.nrows) return;
    const Geo g = make_geo(c, W, ROWS);
    const double bytes = 2.0 * 8.0 * c.cols * c.nrows * c.outer;
    CUtensorMap tm;
    const cuuint64_t dims[3] = {(cuuint64_t)c.pitch * 2, (cuuint64_t)c.nrows, (cuuint64_t)c.outer};
    const cuuint64_t strides[2] = {(cuuint64_t)c.pitch * 8, (cuuint64_t)c.pitch * 8 * c.nrows};
    const cuuint32_t box[3] = {(cuuint32_t)W * 2, (cuuint32_t)(ROWS < 256 ? ROWS : 256), 1};
    const cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, g_buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { std::printf("encode failed %d\n", (int)r); return; }
    const size_t smem = (size_t)S * ROWS * W * 8;
    CK(cudaFuncSetAttribute(copy_tma<W, ROWS, S, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long grid = (long long)ctas_per_sm * g_sms;
    const double ms = time_ms([&] { copy_tma<W, ROWS, S, D><<<(unsigned)grid, 32, smem>>>(tm, g); });
    std::printf("%-14s tma  W=%-2d piece=%-4dB rows/tile=%-4d slots=%d ahead=%d ctas/SM=%d promo=%-5s smem=%3zuK : %8.3f ms  %7.1f GB/s\n",
                c.name, W, W * 8, ROWS, S, D, ctas_per_sm, pname, smem >> 10, ms, bytes / ms * 1e-6);
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    g_sms = prop.multiProcessorCount;
    std::printf("%s, %d SMs\n", prop.name, g_sms);
    const size_t total = (size_t)2048 * 4194304 * 8;   // 64 GiB: one 2048^3 complex64 array
    CK(cudaMalloc(&g_buf, total));
    CK(cudaMemset(g_buf, 0, (size_t)1 << 32));
    // name, pitch (elements), rows, outer blocks, columns used
    const Case cases[] = {
        {"z32MiB", 4194304, 2048, 1, 262144},       // Z pass of 2048^3: 4 GiB of traffic each way
        {"y16KiB", 2048, 2048, 128, 2048},          // Y pass of 2048^3 (128 planes)
        {"z8MiB", 1048576, 1024, 1, 524288},        // Z pass of 1024^3
        {"slabz2KiB", 256, 2048, 1024, 256},        // Z pass of the x-slab layout [Y][Z][X/8]
    };
    for (const Case& c : cases) {
        run_lsu<4, 2048, 512, 1>(c);
        run_lsu<8, 2048, 1024, 1>(c);
        run_lsu<8, 1024, 512, 2>(c);
        run_lsu<16, 1024, 1024, 1>(c);
        run_lsu<16, 512, 512, 2>(c);
        run_lsu<32, 512, 1024, 1>(c);
        run_lsu<32, 256, 512, 2>(c);
        run_lsu2<8, 2048, 1024, 1, 1, 1>(c);
        run_lsu2<8, 2048, 1024, 1, 2, 1>(c);
        run_lsu2<16, 1024, 1024, 1, 1, 1>(c);
        run_lsu2<16, 1024, 1024, 1, 2, 1>(c);
        run_lsu2<8, 2048, 1024, 1, 0, 1>(c);
        run_lsu2<8, 2048, 1024, 1, 0, 2>(c);
        run_lsu2<8, 2048, 1024, 1, 0, 4>(c);
        run_lsu2<8, 1024, 512, 2, 0, 2>(c);
        run_lsu2<8, 1024, 512, 2, 0, 4>(c);
        run_lsu2<4, 2048, 512, 1, 0, 2>(c);
        run_lsu2<4, 2048, 512, 1, 0, 4>(c);
        run_lsu2<4, 2048, 512, 1, 0, 8>(c);
        run_lsu2<8, 2048, 1024, 1, 1, 2>(c);
        run_lsu2<8, 2048, 1024, 1, 2, 2>(c);
        const CUtensorMapL2promotion P0 = CU_TENSOR_MAP_L2_PROMOTION_NONE, P128 = CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                     P256 = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
        run_tma<4, 2048, 3, 2>(c, 1, P0, "none");
        run_tma<4, 2048, 3, 2>(c, 1, P128, "128B");
        run_tma<4, 1024, 6, 4>(c, 1, P128, "128B");
        run_tma<8, 2048, 1, 0>(c, 1, P0, "none");
        run_tma<8, 2048, 1, 0>(c, 1, P128, "128B");
        run_tma<8, 1024, 1, 0>(c, 1, P128, "128B");   // 64 KiB slot, one in flight: what one ring slot of a W=8,N=1024 kernel sees
        run_tma<8, 512, 6, 4>(c, 1, P0, "none");
        run_tma<8, 512, 6, 4>(c, 1, P128, "128B");
        run_tma<8, 512, 6, 4>(c, 1, P256, "256B");
        run_tma<8, 512, 3, 2>(c, 2, P128, "128B");
        run_tma<16, 256, 6, 4>(c, 1, P0, "none");
        run_tma<16, 256, 6, 4>(c, 1, P128, "128B");
        run_tma<16, 256, 6, 4>(c, 1, P256, "256B");
        run_tma<32, 128, 6, 4>(c, 1, P128, "128B");
        run_tma<32, 128, 6, 4>(c, 1, P256, "256B");
        std::printf("\n");
    }
    return 0;
}
