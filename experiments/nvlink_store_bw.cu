// Microbenchmark for the slab exchange: how fast can ONE GPU write a
// stream of contiguous pieces into a PEER GPU's memory over NVLink, as a function of how the stores
// are issued?  Not part of the product; single process, two GPUs with peer access.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o nvlink_store_bw nvlink_store_bw.cu
//   ./nvlink_store_bw [GiB=2] [piece_bytes=2048]
//
// Modes: st.v2 (8 B / thread, 256 B / warp instruction -- what TileThread::store<BLK> does),
//        st.v4 (16 B / thread), bulk (cp.async.bulk from shared memory, one copy per piece -- what
//        store_blocked_bulk does), and cudaMemcpyPeerAsync as the copy-engine reference.
// Each mode is run with the grid capped at 1, 2, 3, 4, 8 CTAs per SM and uncapped.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) { std::printf("%s: %s\n", #x, cudaGetErrorString(e_)); std::exit(1); } \
    } while (0)

template <int VEC>   // VEC = 2 or 4 floats per thread and instruction
__global__ void copy_st(const float* __restrict__ src, float* __restrict__ dst, size_t n_vec) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
        if constexpr (VEC == 2) {
            float2 v = reinterpret_cast<const float2*>(src)[i];
            reinterpret_cast<float2*>(dst)[i] = v;
        } else {
            float4 v = reinterpret_cast<const float4*>(src)[i];
            reinterpret_cast<float4*>(dst)[i] = v;
        }
    }
}

// every CTA stages `pieces_per_cta` pieces of `piece` bytes in shared memory (plain loads + st.shared,
// as an FFT pass would after its last stage) and sends each with one cp.async.bulk
__global__ void copy_bulk(const float4* __restrict__ src, unsigned char* __restrict__ dst, size_t n_pieces, unsigned piece,
                          int pieces_per_cta) {
    extern __shared__ __align__(128) unsigned char sm[];
    const size_t groups = (n_pieces + pieces_per_cta - 1) / pieces_per_cta;
    const unsigned vec_per_group = piece / 16 * pieces_per_cta;
    for (size_t g = blockIdx.x; g < groups; g += gridDim.x) {
        const size_t first = g * pieces_per_cta;
        for (unsigned i = threadIdx.x; i < vec_per_group; i += blockDim.x)
            reinterpret_cast<float4*>(sm)[i] = src[first * (piece / 16) + i];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        for (int i = threadIdx.x; i < pieces_per_cta; i += blockDim.x)
            if (first + i < n_pieces)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + (first + i) * piece),
                             "r"((unsigned)__cvta_generic_to_shared(sm + (size_t)i * piece)), "r"(piece)
                             : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncthreads();
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main(int argc, char** argv) {
    const double gib = argc > 1 ? atof(argv[1]) : 2.0;
    const unsigned piece = argc > 2 ? (unsigned)atoi(argv[2]) : 2048;
    const size_t bytes = (size_t)(gib * (1ull << 30)) / piece * piece;
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev < 2) { std::printf("needs 2 GPUs\n"); return 0; }
    int can = 0;
    CK(cudaDeviceCanAccessPeer(&can, 0, 1));
    if (!can) { std::printf("no peer access 0 -> 1\n"); return 0; }
    float *src = nullptr, *dst = nullptr, *loc = nullptr;
    CK(cudaSetDevice(1));
    CK(cudaMalloc(&dst, bytes));
    CK(cudaSetDevice(0));
    CK(cudaDeviceEnablePeerAccess(1, 0));
    CK(cudaMalloc(&src, bytes));
    CK(cudaMalloc(&loc, bytes));
    CK(cudaMemset(src, 1, bytes));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    auto time = [&](const char* name, int cap, auto&& launch) {
        for (int i = 0; i < 2; ++i) launch();
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        const int reps = 5;
        for (int i = 0; i < reps; ++i) launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        std::printf("%-30s ctas/sm=%-3d %8.1f GB/s\n", name, cap, bytes * (double)reps / (ms * 1e-3) / 1e9);
    };
    time("cudaMemcpyPeerAsync", 0, [&] { CK(cudaMemcpyPeerAsync(dst, 1, src, 0, bytes)); });
    time("local copy st.v4 (HBM)", 0, [&] { copy_st<4><<<sms * 8, 256>>>(src, loc, bytes / 16); });
    const int caps[] = {1, 2, 3, 4, 8, 0};
    for (int cap : caps) {
        const int grid = cap ? sms * cap : sms * 16;
        time("peer st.v2 (8 B/thread)", cap, [&] { copy_st<2><<<grid, 256>>>(src, dst, bytes / 8); });
        time("peer st.v4 (16 B/thread)", cap, [&] { copy_st<4><<<grid, 256>>>(src, dst, bytes / 16); });
        time("peer ld.v2 (pull, 8 B/thread)", cap, [&] { copy_st<2><<<grid, 256>>>(dst, loc, bytes / 8); });
        time("peer ld.v4 (pull, 16 B/thr)", cap, [&] { copy_st<4><<<grid, 256>>>(dst, loc, bytes / 16); });
        const int ppc = 8;                                   // pieces per CTA and round
        const size_t smem = (size_t)ppc * piece;
        CK(cudaFuncSetAttribute(copy_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        time("peer cp.async.bulk", cap, [&] {
            copy_bulk<<<grid, 256, smem>>>(reinterpret_cast<const float4*>(src), reinterpret_cast<unsigned char*>(dst),
                                           bytes / piece, piece, ppc);
        });
    }
    CK(cudaGetLastError());
    return 0;
}
