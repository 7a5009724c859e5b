// __global__ wrappers around the tile thread program + the variant registry.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "fft_core.cuh"

namespace b2 {

// resident-CTA counts (SMs x occupancy) are properties of a (kernel, device) pair
#define B2_MAX_DEVICES 64
inline int b2_current_device() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= B2_MAX_DEVICES) d = 0;
    return d;
}

template <class Cfg, bool SPLIT, bool INV, int s, class TH>
__device__ __forceinline__ void run_stages(TH& th, const PassParams<typename Cfg::T>& p,
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

// Destination-blocked output of a contiguous-axis pass as TMA bulk copies (SASS UBLKCP): the finished
// lines are laid out densely in shared memory and every (line, block) piece -- N/nblocks contiguous
// elements -- leaves as ONE cp.async.bulk to the block's buffer, which in the x-slab exchange is a peer
// GPU's memory over NVLink.  The copy engine of the SM moves kilobyte pieces instead of the LSU moving
// 256 bytes per warp instruction, and no thread stalls on a remote store.
template <class Cfg, bool INV>
__device__ __forceinline__ void store_blocked_bulk(TileThread<Cfg, false, INV, false>& th,
                                                   const PassParams<typename Cfg::T>& p, vec2<typename Cfg::T>* smem,
                                                   long long bid) {
    using T = typename Cfg::T;
    using T2 = vec2<T>;
    constexpr int N = Cfg::N, TPC = Cfg::TPC, S = Cfg::S;
    constexpr int R = Cfg::R(S - 1), LG = ilog2(R), BPT = Cfg::BPT(S - 1);
    // blk_bulk == 2: the lines are staged in a buffer of their own behind the exchange buffer, so the copies of one group of
    // lines drain (over NVLink) while the CTA already loads and transforms the next group; only the staging buffer has
    // to wait for them.  blk_bulk == 1: staged in the exchange buffer itself, the CTA waits for the copies before it goes on.
    const bool own = p.blk_bulk == 2;
    if (own) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous group's copies have read the staging buffer
    __syncthreads();                                   // ... for every issuing thread / the last exchange has been read by everyone
    smem += own ? (long long)Cfg::COL_SMEM * Cfg::W * Cfg::G : 0;
    th.apply_scale(p);
    if (th.active) {
        T2* dst = smem + (long long)th.g * N + th.t;
        static_for<0, BPT>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            static_for<0, R>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                st_c(dst + (i + k * BPT) * TPC, th.v[i * R + brev(k, LG)]);
            });
        });
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async-proxy reads
    __syncthreads();
    const int lg = p.out_blk_log2;
    const int nb = N >> lg;
    const unsigned bytes = (unsigned)(sizeof(T2) << lg);
    for (int i = (int)threadIdx.x; i < Cfg::G * nb; i += Cfg::THREADS) {
        const int g = i / nb, h = i - g * nb;
        const long long tile = bid * Cfg::G + g;
        if (tile < p.n_tiles) {
            T2* gdst = reinterpret_cast<T2*>(p.out_blk0[h]) + TileThread<Cfg, false, INV, false>::line_out_base(tile, p);
            const T2* ssrc = smem + (long long)g * N + ((long long)h << lg);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                         "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes)
                         : "memory");
        }
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    if (!own) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory may be reused (after a barrier)
}

// Split-layout rows: the two output planes of the finished lines are laid out densely in the (now idle) exchange buffer and
// every plane of the group leaves as ONE cp.async.bulk -- instead of 4-byte global stores, of which the split layout needs
// twice as many instructions as the interleaved one needs 8-byte ones (the L1/LSU data pipe is the co-limiter of the row
// kernels).
template <class Cfg, class TH>
__device__ __forceinline__ void store_split_bulk(TH& th, const PassParams<typename Cfg::T>& p, typename Cfg::T* smem, long long grp) {
    using T = typename Cfg::T;
    constexpr int N = Cfg::N, TPC = Cfg::TPC, S = Cfg::S, G = Cfg::G;
    constexpr int R = Cfg::R(S - 1), LG = ilog2(R), BPT = Cfg::BPT(S - 1);
    __syncthreads();                                   // the last exchange has been read by everyone
    th.apply_scale(p);
    if (th.active) {
        T* dre = smem + (long long)th.g * N + th.t;
        T* dim = dre + (long long)G * N;
        static_for<0, BPT>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            static_for<0, R>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                T xr, xi;
                csplit(th.v[i * R + brev(k, LG)], xr, xi);
                dre[(i + k * BPT) * TPC] = xr;
                dim[(i + k * BPT) * TPC] = xi;
            });
        });
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async-proxy reads
    __syncthreads();
    if (threadIdx.x < 2) {
        long long tiles = p.n_tiles - grp * G;
        if (tiles > G) tiles = G;
        const unsigned bytes = (unsigned)(tiles * N * sizeof(T));
        T* gdst = (threadIdx.x == 0 ? p.out0 : p.out1) + grp * G * (long long)N;
        const T* ssrc = smem + (long long)threadIdx.x * G * N;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                     "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the buffer may be reused (after the next barrier)
    }
    __syncthreads();
}

// Four-step pass A on a contiguous axis (inner0 == 1): column w of the tile becomes the N1 contiguous
// outputs [n2][k1 = 0..N1).  The tile is staged TRANSPOSED in shared memory (row pitch N1 + 16 bytes:
// conflict-free for the column-major writes, 16-byte aligned for the copy) and every row leaves as ONE
// cp.async.bulk of N1 complex elements, instead of 32-byte pieces per warp store.
template <class Cfg, bool INV>
__device__ __forceinline__ void store_fs_bulk(TileThread<Cfg, false, INV, true>& th, const PassParams<typename Cfg::T>& p,
                                              vec2<typename Cfg::T>* smem, long long first_tile) {
    using T = typename Cfg::T;
    using T2 = vec2<T>;
    constexpr int N = Cfg::N, W = Cfg::W, TPC = Cfg::TPC, S = Cfg::S;
    constexpr int R = Cfg::R(S - 1), LG = ilog2(R), BPT = Cfg::BPT(S - 1);
    constexpr int P = N + Cfg::FS_PADE;                // staged row pitch in complex elements
    static_assert((long long)P * W <= (long long)Cfg::COL_SMEM * W, "staging fits the exchange buffer");
    __syncthreads();                                   // the last exchange has been read by everyone
    th.apply_fs_twiddle(p);
    if (th.active) {
        T2* dst = smem + ((long long)th.g * W + th.w) * P + th.t;
        static_for<0, BPT>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            static_for<0, R>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                st_c(dst + (i + k * BPT) * TPC, th.v[i * R + brev(k, LG)]);
            });
        });
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    for (int i = (int)threadIdx.x; i < Cfg::G * W; i += Cfg::THREADS) {
        const int g = i / W, w = i - g * W;
        const long long tile = first_tile + g;
        if (tile < p.n_tiles) {
            const long long o = tile / p.inner_blocks, ib = tile - o * p.inner_blocks;
            T2* gdst = reinterpret_cast<T2*>(p.out0) + o * p.out_outer_stride + (ib * W + w) * p.fs_col_stride;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                         "r"((uint32_t)__cvta_generic_to_shared(smem + ((long long)g * W + w) * P)),
                         "r"((unsigned)(N * sizeof(T2)))
                         : "memory");
        }
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory may be reused (after a barrier)
}

// BLK: destination-blocked stores (slab exchange); FS: four-step "A" pass (transposed store +
// inter-pass twiddle, see PassParams).  The two are never combined.
template <class Cfg, bool SPLIT, bool INV, int MINB, bool BLK = false, bool FS = false>
__global__ void __launch_bounds__(Cfg::THREADS, MINB)
tile_fft_kernel(const __grid_constant__ PassParams<typename Cfg::T> p) {
    static_assert(!(BLK && FS), "blocked and transposed stores are exclusive");
    extern __shared__ __align__(16) unsigned char b2_smem_raw[];
    auto* smem = reinterpret_cast<vec2<typename Cfg::T>*>(b2_smem_raw);
    TileThread<Cfg, SPLIT, INV, FS> th;
    // one CTA per G tiles, unless the launch capped the grid (PassParams::max_ctas: exchange passes that
    // are NVLink-bound leave most of each SM to the kernels they overlap with) -- then CTAs stride
    const long long n_groups = (p.n_tiles + Cfg::G - 1) / Cfg::G;
    for (long long bid = blockIdx.x; bid < n_groups; bid += gridDim.x) {
        th.setup((int)threadIdx.x, bid, p);
        th.load(p);
        run_stages<Cfg, SPLIT, INV, 0>(th, p, smem);
        bool bulk = false;
        if constexpr (BLK && Cfg::W == 1 && Cfg::S > 1 && !SPLIT) {
            bulk = p.blk_bulk != 0;
            if (bulk) store_blocked_bulk<Cfg, INV>(th, p, smem, bid);
        }
        if constexpr (FS && Cfg::S > 1 && !SPLIT) {
            bulk = p.fs_bulk != 0;
            if (bulk) store_fs_bulk<Cfg, INV>(th, p, smem, bid * Cfg::G);
        }
        if (!bulk) th.template store<BLK>(p);
        if (bid + (long long)gridDim.x < n_groups) __syncthreads();   // exchange buffer fully read before it is reused
    }
    if constexpr (BLK || FS) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // this thread's bulk stores are done
    if constexpr (BLK) __threadfence_system();      // peer (NVLink) stores visible before the kernel retires
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
    static constexpr size_t IN_BYTES = (size_t)Cfg::G * Cfg::N * Cfg::W * 2 * sizeof(T);  // one ring slot (re+im)
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
        if constexpr (SPLIT && Cfg::S > 1) {
            if (p.split_bulk) { store_split_bulk<Cfg>(th, p, reinterpret_cast<T*>(xbuf), grp); continue; }
        }
        th.store(p);
    }
    if constexpr (SPLIT && Cfg::S > 1) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------ persistent TMA-fed kernel for LONG rows: staging slot = exchange buffer
// Rows of 64-128 KiB (N = 8192 / 16384 complex64, 4096 / 8192 complex128) leave room for one or two CTAs per SM and none for
// a separate staging ring, so the plain kernel's load / transform / store phases run back to back (0.46-0.56 of the copy
// bandwidth at one CTA per SM).  Here ONE shared-memory buffer is both: the next group of rows is fetched into it by
// cp.async.bulk as soon as the LAST exchange of the current group has been read (every thread fences its reads against the
// async proxy and arrives on the "empty" mbarrier; the elected thread waits for that phase and issues the copy), so the
// HBM reads of group j+1 run beside the last radix stage and the global stores of group j.
template <class Cfg>
struct TmaRowAliasLayout {
    using T = typename Cfg::T;
    static constexpr size_t IN_BYTES = (size_t)Cfg::G * Cfg::N * 2 * sizeof(T);
    static constexpr size_t X_BYTES = (size_t)Cfg::SMEM_BYTES;
    static constexpr size_t BUF_BYTES = ((IN_BYTES > X_BYTES ? IN_BYTES : X_BYTES) + 127) / 128 * 128;
    static constexpr size_t BAR_OFF = BUF_BYTES;               // full, empty
    static constexpr size_t TOTAL = BAR_OFF + 16;
};

template <class Cfg, bool SPLIT, bool INV, int s, class TH, class F>
__device__ __forceinline__ void run_stages_hook(TH& th, const PassParams<typename Cfg::T>& p, vec2<typename Cfg::T>* smem,
                                                F&& after_last_read) {
    th.template compute<s>(p);
    if constexpr (s + 1 < Cfg::S) {
        if constexpr (s > 0) __syncthreads();   // previous exchange fully read before it is overwritten
        th.template xwrite<s>(smem);
        __syncthreads();
        th.template xread<s>(smem);
        if constexpr (s + 2 == Cfg::S) after_last_read();
        run_stages_hook<Cfg, SPLIT, INV, s + 1>(th, p, smem, static_cast<F&&>(after_last_read));
    }
}

template <class Cfg, bool SPLIT, bool INV, int MINB>
__global__ void __launch_bounds__(Cfg::THREADS, MINB)
tile_fft_kernel_tma_row_alias(const __grid_constant__ PassParams<typename Cfg::T> p) {
    using T = typename Cfg::T;
    using T2 = vec2<T>;
    using L = TmaRowAliasLayout<Cfg>;
    static_assert(Cfg::W == 1 && Cfg::S >= 2, "contiguous axis, at least one exchange");
    extern __shared__ __align__(128) unsigned char b2_smem_raw[];
    T2* buf = reinterpret_cast<T2*>(b2_smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(b2_smem_raw + L::BAR_OFF);
    uint64_t* empty = full + 1;
    const int tid = (int)threadIdx.x;
    const long long n_groups = (p.n_tiles + Cfg::G - 1) / Cfg::G;
    const long long stride = gridDim.x;
    auto issue = [&](long long grp) {
        long long tiles = p.n_tiles - grp * Cfg::G;
        if (tiles > Cfg::G) tiles = Cfg::G;
        const long long first = grp * Cfg::G * Cfg::N;
        if constexpr (SPLIT) {
            const uint32_t bytes = (uint32_t)(tiles * Cfg::N * sizeof(T));
            mbar_expect_tx(full, 2 * bytes);
            bulk_load(b2_smem_raw, p.in0 + first, bytes, full);
            bulk_load(b2_smem_raw + L::IN_BYTES / 2, p.in1 + first, bytes, full);
        } else {
            const uint32_t bytes = (uint32_t)(tiles * Cfg::N * 2 * sizeof(T));
            mbar_expect_tx(full, bytes);
            bulk_load(b2_smem_raw, reinterpret_cast<const T2*>(p.in0) + first, bytes, full);
        }
    };
    if (tid == 0) {
        mbar_init(full, 1);
        mbar_init(empty, Cfg::THREADS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    long long grp = blockIdx.x;
    if (tid == 0 && grp < n_groups) issue(grp);
    TileThread<Cfg, SPLIT, INV> th;
    for (unsigned it = 0; grp < n_groups; ++it, grp += stride) {
        th.setup(tid, grp, p);
        mbar_wait(full, it & 1);
        th.load_smem(b2_smem_raw, b2_smem_raw + L::IN_BYTES / 2);
        __syncthreads();                       // every thread holds its elements: the buffer becomes the exchange buffer
        run_stages_hook<Cfg, SPLIT, INV, 0>(th, p, buf, [&] {
            // the last exchange has been read by this thread: hand the buffer back to the async proxy
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(empty);
            if (tid == 0 && grp + stride < n_groups) {
                mbar_wait(empty, it & 1);
                issue(grp + stride);
            }
        });
        th.store(p);
    }
}

// ------------------------------------------------------------------ persistent TMA-staged kernel (strided axes)
// Same pipeline as the row version, for tiles of W columns of an [outer][N][inner] array: the tile is
// fetched with cp.async.bulk.tensor (TMA tiled loads, SASS UTMALDG) through a 3-D tensor map
// {inner, N, outer} in boxes of {W, NB, 1}; the boxes land densely as [N][W], which is exactly the
// layout TileThread::load_smem() reads.  Address generation for the strided gather is done by the TMA
// unit and the next tile is in flight while the current one is in registers.
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(dst)),
        "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tm),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

template <class Cfg>
struct TmaColBox {
    static constexpr int NB = Cfg::N < 256 ? Cfg::N : 256;     // rows per TMA box (box dims are <= 256)
    static constexpr int NLOAD = Cfg::N / NB;
};

template <class Cfg, bool SPLIT, bool INV, int MINB, int NBUF, bool BLK, bool FS = false>
__global__ void __launch_bounds__(Cfg::THREADS, MINB)
tile_fft_kernel_tma_col(const __grid_constant__ PassParams<typename Cfg::T> p, const __grid_constant__ CUtensorMap tm0,
                        const __grid_constant__ CUtensorMap tm1, const __grid_constant__ CUtensorMap tmo0,
                        const __grid_constant__ CUtensorMap tmo1) {
    using T = typename Cfg::T;
    using T2 = vec2<T>;
    using L = TmaRowLayout<Cfg, NBUF>;                 // same ring / exchange / barrier layout (IN_BYTES = G*N*W complex)
    constexpr int NB = TmaColBox<Cfg>::NB, NLOAD = TmaColBox<Cfg>::NLOAD;
    extern __shared__ __align__(128) unsigned char b2_smem_raw[];
    T2* xbuf = reinterpret_cast<T2*>(b2_smem_raw + L::X_OFF);
    uint64_t* bars = reinterpret_cast<uint64_t*>(b2_smem_raw + L::BAR_OFF);
    uint64_t* empty = bars + NBUF;
    const int tid = (int)threadIdx.x;
    const long long n_groups = (p.n_tiles + Cfg::G - 1) / Cfg::G;
    const long long stride = gridDim.x;

    auto issue = [&](long long grp, int slot) {
        if (grp >= n_groups) return;
        long long tiles = p.n_tiles - grp * Cfg::G;
        if (tiles > Cfg::G) tiles = Cfg::G;
        unsigned char* dst = b2_smem_raw + (size_t)slot * L::IN_BYTES;
        constexpr uint32_t plane_bytes = (uint32_t)(Cfg::N * Cfg::W * sizeof(T));      // one tile, one real plane
        mbar_expect_tx(&bars[slot], (uint32_t)tiles * 2u * plane_bytes);
        for (int g = 0; g < (int)tiles; ++g) {
            const long long tile = grp * Cfg::G + g;
            const long long o = tile / p.inner_blocks, ib = tile - o * p.inner_blocks;
#pragma unroll
            for (int nb = 0; nb < NLOAD; ++nb) {
                if constexpr (SPLIT) {
                    unsigned char* d0 = dst + ((size_t)g * Cfg::N + (size_t)nb * NB) * Cfg::W * sizeof(T);
                    tma_load_3d(d0, &tm0, (int)(ib * Cfg::W), nb * NB, (int)o, &bars[slot]);
                    tma_load_3d(d0 + L::IN_BYTES / 2, &tm1, (int)(ib * Cfg::W), nb * NB, (int)o, &bars[slot]);
                } else {
                    unsigned char* d0 = dst + ((size_t)g * Cfg::N + (size_t)nb * NB) * Cfg::W * 2 * sizeof(T);
                    tma_load_3d(d0, &tm0, (int)(ib * Cfg::W * 2), nb * NB, (int)o, &bars[slot]);
                }
            }
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
    TileThread<Cfg, SPLIT, INV, FS> th;
    for (unsigned it = 0; grp < n_groups; ++it, grp += stride) {
        const int slot = (int)(it % NBUF);
        th.setup(tid, grp, p);
        mbar_wait(&bars[slot], (it / NBUF) & 1);
        const unsigned char* src = b2_smem_raw + (size_t)slot * L::IN_BYTES;
        th.load_smem(src, src + L::IN_BYTES / 2);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // see tile_fft_kernel_tma_row
        mbar_arrive(&empty[slot]);
        if constexpr (!BLK && !FS) {
            // the previous tile's tensor stores have finished reading the exchange buffer
            if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0 && grp + (long long)NBUF * stride < n_groups) {
            mbar_wait(&empty[slot], (it / NBUF) & 1);
            issue(grp + (long long)NBUF * stride, slot);
        }
        run_stages<Cfg, SPLIT, INV, 0>(th, p, xbuf);
        if constexpr (!BLK && !FS) {
            if (p.tma_store) {
                // Output through the TMA unit as well: the finished tiles are laid out densely ([N][W], the box
                // layout) in the exchange buffer and leave as cp.async.bulk.tensor stores (SASS UTMASTG).  A
                // warp-wide st.global of a W-wide tile touches 4-8 different 128-byte lines (one L1 wavefront
                // each); the shared-memory store is 2 wavefronts and the strided scatter costs no LSU time.
                __syncthreads();                                   // the last exchange has been read by everyone
                th.apply_scale(p);
                if (th.active) {
                    constexpr int S_ = Cfg::S, R_ = Cfg::R(S_ - 1), LG_ = ilog2(R_), BPT_ = Cfg::BPT(S_ - 1);
                    const long long off = ((long long)th.g * Cfg::N + th.t) * Cfg::W + th.w;
                    if constexpr (SPLIT) {
                        T* sre = reinterpret_cast<T*>(xbuf) + off;
                        T* sim = sre + (size_t)Cfg::G * Cfg::N * Cfg::W;
                        static_for<0, BPT_>([&](auto ic) {
                            constexpr int i = decltype(ic)::value;
                            static_for<0, R_>([&](auto kc) {
                                constexpr int k = decltype(kc)::value;
                                T xr, xi;
                                csplit(th.v[i * R_ + brev(k, LG_)], xr, xi);
                                sre[(i + k * BPT_) * Cfg::TPC * Cfg::W] = xr;
                                sim[(i + k * BPT_) * Cfg::TPC * Cfg::W] = xi;
                            });
                        });
                    } else {
                        T2* sc = xbuf + off;
                        static_for<0, BPT_>([&](auto ic) {
                            constexpr int i = decltype(ic)::value;
                            static_for<0, R_>([&](auto kc) {
                                constexpr int k = decltype(kc)::value;
                                st_c(sc + (i + k * BPT_) * Cfg::TPC * Cfg::W, th.v[i * R_ + brev(k, LG_)]);
                            });
                        });
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncthreads();
                if (tid == 0) {
                    long long tiles = p.n_tiles - grp * Cfg::G;
                    if (tiles > Cfg::G) tiles = Cfg::G;
                    for (int g = 0; g < (int)tiles; ++g) {
                        const long long tile = grp * Cfg::G + g;
                        const long long o = tile / p.inner_blocks, ib = tile - o * p.inner_blocks;
#pragma unroll
                        for (int nb = 0; nb < NLOAD; ++nb) {
                            if constexpr (SPLIT) {
                                const T* s0 = reinterpret_cast<const T*>(xbuf) + ((size_t)g * Cfg::N + (size_t)nb * NB) * Cfg::W;
                                tma_store_3d(&tmo0, s0, (int)(ib * Cfg::W), nb * NB, (int)o);
                                tma_store_3d(&tmo1, s0 + (size_t)Cfg::G * Cfg::N * Cfg::W, (int)(ib * Cfg::W), nb * NB, (int)o);
                            } else {
                                const T2* s0 = xbuf + ((size_t)g * Cfg::N + (size_t)nb * NB) * Cfg::W;
                                tma_store_3d(&tmo0, s0, (int)(ib * Cfg::W * 2), nb * NB, (int)o);
                            }
                        }
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                continue;
            }
        }
        if constexpr (FS && Cfg::S > 1 && !SPLIT) {
            if (p.fs_bulk) {
                store_fs_bulk<Cfg, INV>(th, p, xbuf, grp * Cfg::G);
                continue;
            }
        }
        th.template store<BLK>(p);
    }
    if constexpr (!BLK && !FS) {
        if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // tensor stores have completed
    }
    if constexpr (FS) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if constexpr (BLK) __threadfence_system();
}

// ------------------------------------------------------------------ TMA-staged strided kernel, ring slot == exchange buffer
// The big strided tiles (N = 1024/2048) leave room for only ONE CTA per SM when the TMA ring and the
// exchange buffer are separate, so that CTA's load, compute and store phases never overlap with anything
// but its own prefetch.  Here the single staging slot IS the exchange buffer: half the shared memory, two
// CTAs per SM, and the phases of one CTA overlap with those of the other.  The next tile's TMA load is
// issued as soon as every thread has finished the last exchange read (generic -> async proxy fence +
// "empty" mbarrier), i.e. it runs under this CTA's last butterflies and stores.
template <class Cfg>
struct TmaAliasLayout {
    using T = typename Cfg::T;
    static constexpr size_t IN_BYTES = (size_t)Cfg::G * Cfg::N * Cfg::W * 2 * sizeof(T);
    static constexpr size_t X_BYTES = ((size_t)Cfg::SMEM_BYTES + 127) / 128 * 128;
    static constexpr size_t BUF_BYTES = IN_BYTES > X_BYTES ? IN_BYTES : X_BYTES;
    static constexpr size_t BAR_OFF = BUF_BYTES;               // full, empty
    static constexpr size_t TOTAL = BAR_OFF + 16;
};

template <class Cfg, bool SPLIT, bool INV, int MINB, bool BLK>
__global__ void __launch_bounds__(Cfg::THREADS, MINB)
tile_fft_kernel_tma_col_alias(const __grid_constant__ PassParams<typename Cfg::T> p, const __grid_constant__ CUtensorMap tm0,
                              const __grid_constant__ CUtensorMap tm1) {
    using T = typename Cfg::T;
    using T2 = vec2<T>;
    using L = TmaAliasLayout<Cfg>;
    constexpr int NB = TmaColBox<Cfg>::NB, NLOAD = TmaColBox<Cfg>::NLOAD;
    extern __shared__ __align__(128) unsigned char b2_smem_raw[];
    T2* xbuf = reinterpret_cast<T2*>(b2_smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(b2_smem_raw + L::BAR_OFF);
    uint64_t* empty = full + 1;
    const int tid = (int)threadIdx.x;
    const long long n_groups = (p.n_tiles + Cfg::G - 1) / Cfg::G;
    const long long stride = gridDim.x;

    auto issue = [&](long long grp) {
        long long tiles = p.n_tiles - grp * Cfg::G;
        if (tiles > Cfg::G) tiles = Cfg::G;
        constexpr uint32_t plane_bytes = (uint32_t)(Cfg::N * Cfg::W * sizeof(T));
        mbar_expect_tx(full, (uint32_t)tiles * 2u * plane_bytes);
        for (int g = 0; g < (int)tiles; ++g) {
            const long long tile = grp * Cfg::G + g;
            const long long o = tile / p.inner_blocks, ib = tile - o * p.inner_blocks;
#pragma unroll
            for (int nb = 0; nb < NLOAD; ++nb) {
                if constexpr (SPLIT) {
                    unsigned char* d0 = b2_smem_raw + ((size_t)g * Cfg::N + (size_t)nb * NB) * Cfg::W * sizeof(T);
                    tma_load_3d(d0, &tm0, (int)(ib * Cfg::W), nb * NB, (int)o, full);
                    tma_load_3d(d0 + L::IN_BYTES / 2, &tm1, (int)(ib * Cfg::W), nb * NB, (int)o, full);
                } else {
                    unsigned char* d0 = b2_smem_raw + ((size_t)g * Cfg::N + (size_t)nb * NB) * Cfg::W * 2 * sizeof(T);
                    tma_load_3d(d0, &tm0, (int)(ib * Cfg::W * 2), nb * NB, (int)o, full);
                }
            }
        }
    };

    if (tid == 0) {
        mbar_init(full, 1);
        mbar_init(empty, Cfg::THREADS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    long long grp = blockIdx.x;
    if (tid == 0 && grp < n_groups) issue(grp);
    TileThread<Cfg, SPLIT, INV, false> th;
    for (unsigned it = 0; grp < n_groups; ++it, grp += stride) {
        th.setup(tid, grp, p);
        mbar_wait(full, it & 1);
        th.load_smem(b2_smem_raw, b2_smem_raw + L::IN_BYTES / 2);
        __syncthreads();                       // the staged tile is in registers everywhere: the buffer may be overwritten
        run_stages<Cfg, SPLIT, INV, 0>(th, p, xbuf);
        // hand the buffer to the async proxy for the next tile (see tile_fft_kernel_tma_row for the fence)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(empty);
        if (tid == 0 && grp + stride < n_groups) {
            mbar_wait(empty, it & 1);
            issue(grp + stride);
        }
        th.template store<BLK>(p);
    }
    if constexpr (BLK) __threadfence_system();
}


// ------------------------------------------------------------------ fused two-step strided kernel
// Strided axes of length N >= 1024: a tile must be W = 128 bytes / sizeof(complex) columns wide for the memory
// system to run near its peak (measured, profiles/r02_strided_copy_bw.txt: at >= 8 MiB row pitch HBM serves ~45 G
// row pieces per second whatever their size up to 128 bytes, so 64-byte pieces cap a pass at 0.44 of the copy
// bandwidth and 128-byte pieces reach 0.87; at 2-16 KiB pitch it is 0.70 against 1.0), but N x 128 bytes does not
// fit the shared memory of one SM.  This kernel does the axis as N = N1*N2 in two steps inside ONE launch (see
// fused2_setup_a/b in fft_core.cuh): each persistent CTA owns one super-tile at a time, every global access of both
// steps is a full 128-byte line, and the [n2][k1][W] intermediate goes through a per-CTA scratch slot that is
// rewritten every iteration and therefore lives in the 126 MB L2 -- one DRAM read and one DRAM write per element,
// like the single-pass kernels.  While step B runs out of L2 the next super-tile is prefetched into L2.
template <class CfgA, class CfgB, bool INV, int MINB>
__global__ void __launch_bounds__(CfgA::THREADS, MINB)
fused2_fft_kernel(const __grid_constant__ PassParams<typename CfgA::T> pa, const __grid_constant__ PassParams<typename CfgA::T> pb,
                  const long long inner_in, const long long inner_out) {
    using T = typename CfgA::T;
    using T2 = vec2<T>;
    static_assert(CfgA::THREADS == CfgB::THREADS && CfgA::W == CfgB::W, "both steps run on the same CTA shape");
    constexpr int W = CfgA::W, N1 = CfgA::N, N2 = CfgB::N;
    static_assert(N2 % CfgA::G == 0 && N1 % CfgB::G == 0, "sub-tiles must divide the super-tile");
    constexpr int NSA = N2 / CfgA::G, NSB = N1 / CfgB::G;
    extern __shared__ __align__(16) unsigned char b2_smem_raw[];
    T2* smem = reinterpret_cast<T2*>(b2_smem_raw);
    const int tid = (int)threadIdx.x;
    const long long slot = (long long)blockIdx.x * ((long long)N1 * N2 * W);
    // L2 policies: the input and output streams are touched once (evict first), the scratch slot is rewritten every
    // iteration and must stay resident (evict last); step B drops every scratch line it has consumed (discard.L2) so that
    // the dirty intermediate is never written back to DRAM
    unsigned long long pol_stream, pol_scratch;
    {
        unsigned long long pf, pl, pn;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pf));
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pl));
        asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pn));
        pol_stream = (pa.fused_flags & 4) ? pf : pn;
        pol_scratch = (pa.fused_flags & 8) ? pl : pn;
    }
    for (long long s = blockIdx.x; s < pa.n_tiles; s += gridDim.x) {
        const long long o = s / pa.inner_blocks, ib = s - o * pa.inner_blocks;
        {
            TileThread<CfgA, false, INV, true, true> th;
            th.pol_in = pol_stream; th.pol_out = pol_scratch;
            const long long in_base = o * pa.outer_stride + ib * W;
#pragma unroll 1
            for (int c = 0; c < NSA; ++c) {
                fused2_setup_a<CfgA, CfgB>(th, tid, c, in_base, inner_in, slot);
                th.load(pa);
                run_stages<CfgA, false, INV, 0>(th, pa, smem);
                th.store(pa);
                if constexpr (CfgA::S > 1) __syncthreads();      // exchange buffer fully read before the next sub-tile
            }
        }
        __syncthreads();                                         // the whole intermediate is in the scratch slot
        {
            const long long s2 = s + gridDim.x;                  // next super-tile of this CTA -> L2
            if ((pa.fused_flags & 2) && s2 < pa.n_tiles) {
                const long long o2 = s2 / pa.inner_blocks, ib2 = s2 - o2 * pa.inner_blocks;
                const T2* nb = reinterpret_cast<const T2*>(pa.in0) + o2 * pa.outer_stride + ib2 * W;
                for (int r = tid; r < N1 * N2; r += CfgA::THREADS)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(nb + (long long)r * inner_in));
            }
        }
        {
            TileThread<CfgB, false, INV, false, true> th;
            th.pol_in = pol_scratch; th.pol_out = pol_stream;
            const long long out_base = o * pb.out_outer_stride + ib * W;
#pragma unroll 1
            for (int c = 0; c < NSB; ++c) {
                fused2_setup_b<CfgA, CfgB>(th, tid, c, out_base, inner_out, slot);
                th.load(pb);
                run_stages<CfgB, false, INV, 0>(th, pb, smem);
                th.store(pb);
                if (pa.fused_flags & 1) {
                    // the W lanes that share a scratch row have consumed it (their loads fed the butterflies above): drop the
                    // 128-byte lines of this thread row, lane w taking rows j = w, w + W, ..
                    __syncwarp();
                    const T2* row0 = reinterpret_cast<const T2*>(pb.in0) + (th.base - th.w) + (long long)th.t * pb.inner;
                    for (int j = th.w; j < CfgB::E; j += W)
                        asm volatile("discard.global.L2 [%0], 128;" ::"l"(row0 + (long long)j * CfgB::TPC * pb.inner) : "memory");
                }
                if constexpr (CfgB::S > 1) __syncthreads();
            }
        }
        __syncthreads();                                         // slot fully read before the next super-tile overwrites it
    }
}


// ------------------------------------------------------------------ fused two-step strided kernel, intermediate in shared memory
// Same decomposition as fused2_fft_kernel, for N = E_A * E_B with both steps done as single register FFTs (radix-32 and
// radix-64 for N = 2048): no exchange buffer is needed, so the [k1][n2][W] intermediate of the 128-byte-wide super-tile
// occupies the CTA's shared memory (KS of the N1 rows k1; the rest -- 1/8 of it for N = 2048 -- goes through a 32 KiB
// global scratch slot per CTA, 5 MB in total, which does stay in L2).  The L2-scratch version above loses 35-45 % of its
// DRAM bandwidth to write-backs and re-reads of the 38-76 MB scratch (profiles/r02_fused2.md); here DRAM sees one read and
// one write per element.  One CTA per SM: the loads of the next step-A sub-tile are issued into a second register set
// before the current one is transformed, so HBM latency hides behind the butterflies.
template <class CfgA, class CfgB, int KS, bool INV>
__global__ void __launch_bounds__(CfgA::THREADS, 1)
fused2s_fft_kernel(const __grid_constant__ PassParams<typename CfgA::T> pa, const __grid_constant__ PassParams<typename CfgA::T> pb,
                   const long long inner_in, const long long inner_out) {
    using T = typename CfgA::T;
    using T2 = vec2<T>;
    static_assert(CfgA::THREADS == CfgB::THREADS && CfgA::W == CfgB::W, "both steps run on the same CTA shape");
    constexpr int W = CfgA::W, N1 = CfgA::N, N2 = CfgB::N;
    static_assert(N2 % CfgA::G == 0 && N1 % CfgB::G == 0 && KS <= N1 && (KS % 2) == 0, "sub-tiles must divide the super-tile");
    constexpr int NSA = N2 / CfgA::G, NSB = N1 / CfgB::G;
    extern __shared__ __align__(16) unsigned char b2_smem_raw[];
    T2* smem_i = reinterpret_cast<T2*>(b2_smem_raw);
    const int tid = (int)threadIdx.x;
    T2* scratch_slot = reinterpret_cast<T2*>(pa.out0) + (long long)blockIdx.x * ((long long)(N1 - KS) * N2 * W);
    unsigned long long pol_stream, pol_scratch;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_scratch));
    using THA = TileThread<CfgA, false, INV, true, true>;
    for (long long s = blockIdx.x; s < pa.n_tiles; s += gridDim.x) {
        const long long o = s / pa.inner_blocks, ib = s - o * pa.inner_blocks;
        const long long in_base = o * pa.outer_stride + ib * W;
        {
            THA ta, tb;
            ta.pol_in = tb.pol_in = pol_stream;
            fused2_setup_a<CfgA, CfgB>(ta, tid, 0, in_base, inner_in, 0);
            ta.load(pa);
            static_for<0, NSA>([&](auto cc) {
                constexpr int c = decltype(cc)::value;
                THA& cur = (c % 2 == 0) ? ta : tb;
                THA& nxt = (c % 2 == 0) ? tb : ta;
                if constexpr (c + 1 < NSA) {
                    fused2_setup_a<CfgA, CfgB>(nxt, tid, c + 1, in_base, inner_in, 0);
                    nxt.load(pa);
                }
                cur.template compute<0>(pa);
                cur.apply_fs_twiddle(pa);
                fused2s_store_a<CfgA, CfgB, KS>(cur, smem_i, scratch_slot, pol_scratch);
            });
        }
        __syncthreads();                                         // the whole intermediate is in place
        {
            const long long s2 = s + gridDim.x;                  // first step-A sub-tile of this CTA's next super-tile -> L2
            if (s2 < pa.n_tiles) {
                const long long o2 = s2 / pa.inner_blocks, ib2 = s2 - o2 * pa.inner_blocks;
                const T2* nb = reinterpret_cast<const T2*>(pa.in0) + o2 * pa.outer_stride + ib2 * W;
                for (int r = tid; r < N1 * CfgA::G; r += CfgA::THREADS)      // rows n1*N2 + n2, n2 < GA
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(nb + ((long long)(r / CfgA::G) * N2 + (r % CfgA::G)) * inner_in));
            }
        }
        {
            TileThread<CfgB, false, INV, false, true> th;
            th.pol_out = pol_stream;
            const long long out_base = o * pb.out_outer_stride + ib * W;
#pragma unroll 1
            for (int c = 0; c < NSB; ++c) {
                fused2_setup_b<CfgA, CfgB>(th, tid, c, out_base, inner_out, 0);
                fused2s_load_b<CfgA, CfgB, KS>(th, c * CfgB::G + th.g, smem_i, scratch_slot, pol_scratch);
                th.template compute<0>(pb);
                th.store(pb);
            }
        }
        __syncthreads();                                         // intermediate fully read before the next super-tile overwrites it
    }
}


// ------------------------------------------------------------------ fused two-step strided kernel, streamed in place through shared memory
// fused2s keeps the [k1][n2][W] intermediate in shared memory but loads step A through registers and has no global load in
// flight during step B: with one CTA of 8 warps per SM the DRAM read stream of an SM stops for half of every super-tile
// (profiles/r02_fused2.md: 0.9 IPC, long-scoreboard bound).  Here the shared-memory tile is also the landing zone of the
// NEXT super-tile: as soon as a step-B sub-tile has pulled its GB rows k1 into registers (and the CTA has passed a
// barrier) those rows are refilled with input rows n1 = k1 of the CTA's next super-tile by asynchronous 16-byte copies
// (cp.async.cg, SASS LDGSTS: eight lanes per 128-byte row piece, no registers held), which run beside the radix-64
// butterflies and the global stores of the current one.  Step A then finds its input in shared memory and transforms every
// column in place.  The N1-KS rows that do not fit (1/8 of a 2048-point tile) are prefetched into L2 during step B and read
// one sub-tile ahead during step A, as in fused2s.  DRAM sees one read and one write per element; both streams stay busy.
__device__ __forceinline__ void b2_cp_async16(void* dst_smem, const void* src, unsigned long long pol) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "l"(pol) : "memory");
}
__device__ __forceinline__ void b2_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void b2_progress_add(unsigned* counter) {
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
}
template <int PENDING>
__device__ __forceinline__ void b2_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory"); }

// Refill of the rows one WARP owns in step-B sub-tile c (the rows its own lanes have just read into registers: k1 =
// c*GB + warp*RPW + i, so no CTA barrier is needed, only warp convergence): input rows n = k1*N2 + n2 of the super-tile
// at `tile_in`, copied segment by segment (segment = the GA values of n2 of one step-A sub-tile) with one commit group per
// segment, so that step A of the next super-tile can start on segment 0 while the later ones are still in flight.
template <class CfgA, class CfgB, int KS>
__device__ __forceinline__ void fused2p_refill(const vec2<typename CfgA::T>* tile_in, long long inner_in,
                                               vec2<typename CfgA::T>* smem_i, int lane, int k0, bool on,
                                               unsigned long long pol) {
    using CH = Fused2PChunk<CfgA>;
    using M = Fused2PRows<CfgA, CfgB>;
    const int q = CH::elem(lane), pl = (int)CH::row(lane);      // chunk column, piece within one warp iteration
    static_for<0, M::NSEG>([&](auto sc) {
        constexpr int sg = decltype(sc)::value;
        if (on) {
            static_for<0, M::ITERS>([&](auto ic) {
                constexpr int it = decltype(ic)::value;
                const int n = M::row(k0, sg, it * M::PPI + pl);
                if (M::k_of(k0, it * M::PPI + pl) < KS) b2_cp_async16(smem_i + (long long)n * CfgA::W + q, tile_in + (long long)n * inner_in + q, pol);
            });
        }
        b2_cp_async_commit();
    });
}

template <class CfgA, class CfgB, int KS, bool INV>
__global__ void __launch_bounds__(CfgA::THREADS, 1)
fused2p_fft_kernel(const __grid_constant__ PassParams<typename CfgA::T> pa, const __grid_constant__ PassParams<typename CfgA::T> pb,
                   const long long inner_in, const long long inner_out) {
    using T = typename CfgA::T;
    using T2 = vec2<T>;
    using C = cpx<T>;
    using M = Fused2PRows<CfgA, CfgB>;
    static_assert(CfgA::THREADS == CfgB::THREADS && CfgA::W == CfgB::W, "both steps run on the same CTA shape");
    constexpr int W = CfgA::W, N1 = CfgA::N, N2 = CfgB::N, GA = CfgA::G, GB = CfgB::G, THREADS = CfgA::THREADS;
    static_assert(N2 % GA == 0 && N1 % GB == 0 && KS <= N1, "sub-tiles must divide the super-tile");
    constexpr int NSA = N2 / GA, NSB = N1 / GB, NX = N1 - KS;
    static_assert(GB * W == THREADS && M::RPW * (THREADS / 32) == GB, "a step-B sub-tile is RPW rows per warp");
    extern __shared__ __align__(16) unsigned char b2_smem_raw[];
    T2* smem_i = reinterpret_cast<T2*>(b2_smem_raw);
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    T2* scratch_slot = reinterpret_cast<T2*>(pa.out0) + (long long)blockIdx.x * ((long long)(NX > 0 ? NX : 1) * N2 * W);
    const T2* in = reinterpret_cast<const T2*>(pa.in0);
    unsigned long long pol_stream, pol_scratch;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_scratch));
    using THA = TileThread<CfgA, false, INV, true, true>;
    auto tile_in_base = [&](long long s) {
        const long long o = s / pa.inner_blocks, ib = s - o * pa.inner_blocks;
        return o * pa.outer_stride + ib * W;
    };
    long long s = blockIdx.x;
    {   // first super-tile of this CTA: the same NSB x NSEG commit groups a step B would have issued
        const T2* first_in = in + (s < pa.n_tiles ? tile_in_base(s) : 0);
        for (int c = 0; c < NSB; ++c)
            fused2p_refill<CfgA, CfgB, KS>(first_in, inner_in, smem_i, lane, c * GB + warp * M::RPW, s < pa.n_tiles, pol_stream);
    }
    for (; s < pa.n_tiles; s += gridDim.x) {
        const long long o = s / pa.inner_blocks, ib = s - o * pa.inner_blocks;
        const long long in_base = o * pa.outer_stride + ib * W;
        {
            THA th;
            C xa[NX > 0 ? NX : 1], xb[NX > 0 ? NX : 1];
            th.pol_in = pol_stream;
            // rows n1 >= KS of sub-tile c: (n1*N2 + c*GA + g) * inner_in, one sub-tile ahead of the butterflies
            auto load_extra = [&](C* x, int c) {
                if constexpr (NX > 0) {
                    const T2* p = in + in_base + ((long long)KS * N2 + c * GA + tid / W) * inner_in + tid % W;
                    static_for<0, NX>([&](auto ic) {
                        x[decltype(ic)::value] = ld_stream_c_pol(p + (long long)decltype(ic)::value * N2 * inner_in, pol_stream);
                    });
                }
            };
            // inter-step twiddles of sub-tile c (fs_base_load), also one sub-tile ahead: the table reads are L1 hits, but with
            // two warps per scheduler nothing hides even that latency behind the butterflies unless they are issued early
            C la[4], ha[N1 / 4], lb[4], hb[N1 / 4];
            th.fs_n2i = tid / W;
            th.fs_base_load(pa, la, ha);
            load_extra(xa, 0);
            static_for<0, NSA>([&](auto cc) {
                constexpr int c = decltype(cc)::value;
                C* cur = (c % 2 == 0) ? xa : xb;
                C* nxt = (c % 2 == 0) ? xb : xa;
                C(&lc)[4] = (c % 2 == 0) ? la : lb;
                C(&hc)[N1 / 4] = (c % 2 == 0) ? ha : hb;
                C(&ln)[4] = (c % 2 == 0) ? lb : la;
                C(&hn)[N1 / 4] = (c % 2 == 0) ? hb : ha;
                if constexpr (c + 1 < NSA) {
                    load_extra(nxt, c + 1);
                    th.fs_n2i = (c + 1) * GA + tid / W;
                    th.fs_base_load(pa, ln, hn);
                }
                // segment c of every row has landed once this thread's groups up to (last step-B sub-tile, c) are complete
                // and every other thread has seen the same for its own
                b2_cp_async_wait<NSA - 1 - c>();
                __syncthreads();
                if constexpr (c == 0) {
                    // every thread of the CTA has issued the stores of the previous super-tile: publish it (the barrier orders
                    // their stores before this thread's fence, the fence is cumulative, the add is the release)
                    if (pa.progress != nullptr && tid == 0 && s != (long long)blockIdx.x)
                        b2_progress_add(pa.progress + (s - gridDim.x) / pa.progress_tiles);
                }
                fused2_setup_a<CfgA, CfgB>(th, tid, c, in_base, inner_in, 0);
                fused2p_load_a<CfgA, CfgB, KS>(th, smem_i, cur);
                th.template compute<0>(pa);
                th.fs_base_apply(lc, hc);
                fused2s_store_a<CfgA, CfgB, KS>(th, smem_i, scratch_slot, pol_scratch);   // same column, same thread: in place
            });
        }
        __syncthreads();                                         // the whole intermediate is in place
        {
            const long long s2 = s + gridDim.x;
            const bool has_next = s2 < pa.n_tiles;
            const T2* next_in = in + (has_next ? tile_in_base(s2) : 0);
            if (NX > 0 && has_next)                              // the rows that are not staged -> L2
                for (int r = tid; r < NX * N2; r += THREADS)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(next_in + (long long)(KS * N2 + r) * inner_in));
            TileThread<CfgB, false, INV, false, true> th;
            th.pol_out = pol_stream;
            const long long out_base = o * pb.out_outer_stride + ib * W;
#pragma unroll 1
            for (int c = 0; c < NSB; ++c) {
                fused2_setup_b<CfgA, CfgB>(th, tid, c, out_base, inner_out, 0);
                fused2s_load_b<CfgA, CfgB, KS>(th, c * GB + th.g, smem_i, scratch_slot, pol_scratch);
                __syncwarp();                                    // this warp's rows are in registers: refill them
                fused2p_refill<CfgA, CfgB, KS>(next_in, inner_in, smem_i, lane, c * GB + warp * M::RPW, has_next, pol_stream);
                th.template compute<0>(pb);
                th.store(pb);
            }
        }
    }
    b2_cp_async_wait<0>();
    if (pa.progress != nullptr && (long long)blockIdx.x < pa.n_tiles) {      // this CTA's last super-tile is s - gridDim.x
        __syncthreads();
        if (tid == 0) b2_progress_add(pa.progress + (s - gridDim.x) / pa.progress_tiles);
    }
}


// ------------------------------------------------------------------ fused two-step strided kernel, lane-pair steps + warp shuffles
// fused2s is bound by issue latency: its register FFTs (32 and 64 elements per thread) leave room for only 8 warps per SM
// (profiles/r02_fused2.md: 0.9 IPC, exact DRAM traffic).  Here every step is shared by a lane pair (16 and 32 elements per
// thread, the radix-2 stage exchanged with shfl.sync -- see PairFFT), so a CTA has 512 threads at <= 128 registers: twice the
// warps for the same shared-memory-resident intermediate, and room for a second register set that holds the next step-A
// sub-tile while the current one is transformed.
template <int LOG2A, int LOG2B, int KS, bool INV>
__global__ void __launch_bounds__(512, 1)
fused2w_fft_kernel(const __grid_constant__ PassParams<float> pa, const __grid_constant__ PassParams<float> pb,
                   const long long inner_in, const long long inner_out) {
    using F = Fused2W<LOG2A, LOG2B, 16, KS, INV>;
    using T2 = vec2<float>;
    using C = cpx<float>;
    constexpr int W = F::W, N1 = F::N1, N2 = F::N2, G = 512 / (2 * W);
    static_assert(N2 % G == 0 && N1 % G == 0, "sub-tiles must divide the super-tile");
    constexpr int NSA = N2 / G, NSB = N1 / G;
    extern __shared__ __align__(16) unsigned char b2_smem_raw[];
    T2* smem_i = reinterpret_cast<T2*>(b2_smem_raw);
    const int tid = (int)threadIdx.x;
    const int w = tid % W, t = (tid / W) & 1, g = tid / (2 * W);
    T2* scratch_slot = reinterpret_cast<T2*>(pa.out0) + (long long)blockIdx.x * ((long long)(N1 - KS) * N2 * W);
    const T2* fs_tab = reinterpret_cast<const T2*>(pa.fs_t2);
    unsigned long long pol_stream, pol_scratch;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_scratch));
    for (long long s = blockIdx.x; s < pa.n_tiles; s += gridDim.x) {
        const long long o = s / pa.inner_blocks, ib = s - o * pa.inner_blocks;
        const T2* in_col = reinterpret_cast<const T2*>(pa.in0) + o * pa.outer_stride + ib * W + w;
        {
            C va[F::EA], vb[F::EA], send[F::HA], recv[F::HA];
            F::a_load(va, in_col + (long long)g * inner_in, pa.inner, t, pol_stream);
            static_for<0, NSA>([&](auto cc) {
                constexpr int c = decltype(cc)::value;
                C* cur = (c % 2 == 0) ? va : vb;
                C* nxt = (c % 2 == 0) ? vb : va;
                if constexpr (c + 1 < NSA) F::a_load(nxt, in_col + (long long)((c + 1) * G + g) * inner_in, pa.inner, t, pol_stream);
                F::PA::pre(cur, t, send);
                F::PA::exchange(send, recv);
                F::PA::post(cur, t, recv);
                F::a_store(cur, t, c * G + g, w, fs_tab, smem_i, scratch_slot, pol_scratch);
            });
        }
        __syncthreads();                                         // the whole intermediate is in place
        {
            C v[F::EB], send[F::HB], recv[F::HB];
            T2* out_col = reinterpret_cast<T2*>(pb.out0) + o * pb.out_outer_stride + ib * W + w;
#pragma unroll 1
            for (int c = 0; c < NSB; ++c) {
                const int k1 = c * G + g;
                F::b_load(v, t, k1, w, smem_i, scratch_slot, pol_scratch);
                F::PB::pre(v, t, send);
                F::PB::exchange(send, recv);
                F::PB::post(v, t, recv);
                F::b_store(v, t, out_col + (long long)k1 * inner_out, pb.out_inner, pb.scale, pb.scale_mode, pol_stream);
            }
        }
        __syncthreads();                                         // intermediate fully read before the next super-tile overwrites it
    }
}

// ------------------------------------------------------------------ short contiguous rows: 16-byte loads + warp-shuffle exchanges
// N = 4 .. 32 complex64 (interleaved or split): see ShflRow in fft_core.cuh.  256 threads per CTA, N/2 lanes per row, every
// thread walks rows with a grid stride (its twiddles are set up once).  No shared memory, ~40 registers.
__device__ __forceinline__ cpx<float> b2_shfl_xor_c(const cpx<float>& a, int mask) {
#if defined(__CUDA_ARCH__)
    cpx<float> r;
    r.v = __shfl_xor_sync(0xffffffffu, a.v, mask);        // one 64-bit register pair = one complex value
    return r;
#else
    (void)mask;
    return a;
#endif
}

template <int LOG2N, bool SPLIT, bool INV>
__global__ void __launch_bounds__(256)
row_shfl_kernel(const __grid_constant__ PassParams<float> p) {
    using R = ShflRow<LOG2N, INV>;
    using C = cpx<float>;
    constexpr int N = R::N, LP = R::LP, RPW = 32 / LP;                  // rows per warp and step
    constexpr int U = 4;                                                // independent rows per thread in flight (latency hiding)
    const int lane = (int)threadIdx.x & 31, l = lane % LP;
    R r;
    r.init(l);
    const int k0 = R::out_index(l);
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (long long row0 = warp * (RPW * U); row0 < p.n_tiles; row0 += warps * (RPW * U)) {   // warp-uniform: the shuffles need every lane
        C v[U][2];
        static_for<0, U>([&](auto uc) {
            constexpr int u = decltype(uc)::value;
            const long long row = row0 + u * RPW + lane / LP;
            v[u][0] = cmake<float>(0.f, 0.f); v[u][1] = v[u][0];
            if (row < p.n_tiles) {
                if constexpr (SPLIT) {
                    float2 re, im;
                    const float2* pr = reinterpret_cast<const float2*>(p.in0 + row * p.outer_stride) + l;
                    const float2* pi = reinterpret_cast<const float2*>(p.in1 + row * p.outer_stride) + l;
                    asm volatile("ld.global.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(re.x), "=f"(re.y) : "l"(pr));
                    asm volatile("ld.global.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(im.x), "=f"(im.y) : "l"(pi));
                    v[u][0] = cmake<float>(re.x, im.x); v[u][1] = cmake<float>(re.y, im.y);
                } else {
                    float4 q;
                    const float4* ps = reinterpret_cast<const float4*>(reinterpret_cast<const vec2<float>*>(p.in0) + row * p.outer_stride) + l;
                    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "l"(ps));
                    v[u][0] = cmake<float>(q.x, q.y); v[u][1] = cmake<float>(q.z, q.w);
                }
            }
        });
        static_for<0, U>([&](auto uc) {
            constexpr int u = decltype(uc)::value;
            const long long row = row0 + u * RPW + lane / LP;
            static_for<0, R::NST>([&](auto sc) {
                constexpr int s = decltype(sc)::value;
                constexpr int mask = (N >> (s + 1)) >> 1;
                C o[2];
                o[0] = b2_shfl_xor_c(v[u][0], mask);
                o[1] = b2_shfl_xor_c(v[u][1], mask);
                r.template stage<s>(v[u], o);
            });
            R::last(v[u]);
            R::scale(v[u], p.scale, p.scale_mode);
            if (row < p.n_tiles) {
                if constexpr (SPLIT) {
                    float* orow = p.out0 + row * p.out_outer_stride + k0;
                    float* irow = p.out1 + row * p.out_outer_stride + k0;
                    float a, b;
                    csplit(v[u][0], a, b); orow[0] = a; irow[0] = b;
                    csplit(v[u][1], a, b); orow[N / 2] = a; irow[N / 2] = b;
                } else {
                    vec2<float>* orow = reinterpret_cast<vec2<float>*>(p.out0) + row * p.out_outer_stride + k0;
                    st_c(orow, v[u][0]);
                    st_c(orow + N / 2, v[u][1]);
                }
            }
        });
    }
}

// ... four elements per lane (ShflRow4): 16-byte loads AND 16-byte stores, one shuffle stage fewer.  N = 4 .. 128.
template <int LOG2N, bool SPLIT, bool INV>
__global__ void __launch_bounds__(256)
row_shfl4_kernel(const __grid_constant__ PassParams<float> p) {
    using R = ShflRow4<LOG2N, INV>;
    using C = cpx<float>;
    constexpr int N = R::N, LP = R::LP, RPW = 32 / LP;                  // rows per warp and step
    constexpr int U = 2;                                                // independent rows per thread in flight
    const int lane = (int)threadIdx.x & 31, l = lane % LP;
    R r;
    r.init(l);
    const int k0 = 2 * R::out_index(l);
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (long long row0 = warp * (RPW * U); row0 < p.n_tiles; row0 += warps * (RPW * U)) {   // warp-uniform: the shuffles need every lane
        C v[U][4];
        static_for<0, U>([&](auto uc) {
            constexpr int u = decltype(uc)::value;
            const long long row = row0 + u * RPW + lane / LP;
            static_for<0, 4>([&](auto rc) { v[u][decltype(rc)::value] = cmake<float>(0.f, 0.f); });
            if (row < p.n_tiles) {
                static_for<0, 2>([&](auto hc) {
                    constexpr int h = decltype(hc)::value;
                    if constexpr (SPLIT) {
                        float2 re, im;
                        const float2* pr = reinterpret_cast<const float2*>(p.in0 + row * p.outer_stride + (N / 2) * h) + l;
                        const float2* pi = reinterpret_cast<const float2*>(p.in1 + row * p.outer_stride + (N / 2) * h) + l;
                        asm volatile("ld.global.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(re.x), "=f"(re.y) : "l"(pr));
                        asm volatile("ld.global.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(im.x), "=f"(im.y) : "l"(pi));
                        v[u][2 * h] = cmake<float>(re.x, im.x); v[u][2 * h + 1] = cmake<float>(re.y, im.y);
                    } else {
                        float4 q;
                        const float4* ps = reinterpret_cast<const float4*>(reinterpret_cast<const vec2<float>*>(p.in0) + row * p.outer_stride + (N / 2) * h) + l;
                        asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "l"(ps));
                        v[u][2 * h] = cmake<float>(q.x, q.y); v[u][2 * h + 1] = cmake<float>(q.z, q.w);
                    }
                });
            }
        });
        static_for<0, U>([&](auto uc) {
            constexpr int u = decltype(uc)::value;
            const long long row = row0 + u * RPW + lane / LP;
            r.first(v[u]);
            static_for<0, R::NSX>([&](auto sc) {
                constexpr int sx = decltype(sc)::value;
                constexpr int mask = (N >> (sx + 2)) >> 1;
                C o[4];
                static_for<0, 4>([&](auto rc) { o[decltype(rc)::value] = b2_shfl_xor_c(v[u][decltype(rc)::value], mask); });
                r.template stage<sx>(v[u], o);
            });
            R::last(v[u]);
            R::scale(v[u], p.scale, p.scale_mode);
            if (row < p.n_tiles) {
                static_for<0, 2>([&](auto bc) {
                    constexpr int b = decltype(bc)::value;                  // {X[k0 + (N/2) b], X[k0 + 1 + (N/2) b]} = registers b, 2 + b
                    float ar, ai, br, bi;
                    csplit(v[u][b], ar, ai);
                    csplit(v[u][2 + b], br, bi);
                    if constexpr (SPLIT) {
                        float2* orow = reinterpret_cast<float2*>(p.out0 + row * p.out_outer_stride + (N / 2) * b + k0);
                        float2* irow = reinterpret_cast<float2*>(p.out1 + row * p.out_outer_stride + (N / 2) * b + k0);
                        *orow = make_float2(ar, br);
                        *irow = make_float2(ai, bi);
                    } else {
                        float4* orow = reinterpret_cast<float4*>(reinterpret_cast<vec2<float>*>(p.out0) + row * p.out_outer_stride + (N / 2) * b + k0);
                        *orow = make_float4(ar, ai, br, bi);
                    }
                });
            }
        });
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
    int kind;        // 0 = direct global loads, 1 = persistent + TMA bulk staging (contiguous axis), 2 = persistent + TMA tensor staging (strided axes),
                     // 3 = fused two-step strided kernel (N = N1*N2 through an L2-resident scratch slot per CTA),
                     // 4 = short-row kernel (16-byte accesses + warp shuffles; plain contiguous passes only)
    int nbuf;        // ring depth for kind 1
    int blk;         // 1: also compiled with destination-blocked stores (slab exchange passes)
    int fs;          // 1: also compiled as a four-step "A" pass (transposed store + inter-pass twiddle)
    // launches ceil(n_tiles / G) CTAs; params points at a PassParams<T> of the right T
    cudaError_t (*launch)(int split, int inv, const void* params, cudaStream_t stream);
    cudaError_t (*prepare)();   // one-time function attributes (dynamic smem opt-in)
    // occupancy (CTAs/SM) of the interleaved forward kernel, for the tuning report
    int (*occupancy)();
    // kind 3 only: step A is log2n1 long with radices radix[]/S/E, step B has radix_b[]/S_b/E_b; the scratch buffer
    // needs slot_elems complex elements for each of the grid_slots() CTAs the launch may use
    int log2n1;
    int S_b, E_b;
    int radix_b[4];
    long long slot_elems;
    int (*grid_slots)();
    int progress;    // 1: the kernel publishes per-chunk progress counters (PassParams::progress)
};

template <class Cfg, int MINB, bool BLKCAP = false, bool FSCAP = false>
struct VariantOps {
    using T = typename Cfg::T;
    template <bool SPLIT, bool INV, bool BLK, bool FS>
    static cudaError_t attr() {
        return cudaFuncSetAttribute(tile_fft_kernel<Cfg, SPLIT, INV, MINB, BLK, FS>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((BLK && BLK_OWN_OK) ? BLK_SMEM : Cfg::SMEM_BYTES));
    }
    // blocked bulk stores with their own staging buffer (PassParams::blk_bulk == 2): G dense lines behind the exchange buffer
    static constexpr size_t STAGE_BYTES = (size_t)Cfg::G * Cfg::N * Cfg::W * 2 * sizeof(T);
    static constexpr size_t BLK_SMEM = (size_t)Cfg::SMEM_BYTES + STAGE_BYTES;
    static constexpr bool BLK_OWN_OK = BLK_SMEM * MINB <= 200 * 1024;
    template <bool BLK, bool FS>
    static cudaError_t attr3() {
        cudaError_t e;
        if ((e = attr<false, false, BLK, FS>()) != cudaSuccess) return e;
        if ((e = attr<false, true, BLK, FS>()) != cudaSuccess) return e;
        return attr<true, false, BLK, FS>();
    }
    static cudaError_t prepare() {
        cudaError_t e = cudaSuccess;
        if constexpr (BLKCAP) {
            if (BLK_SMEM > 48 * 1024 && BLK_OWN_OK) { if ((e = attr3<true, false>()) != cudaSuccess) return e; }
        }
        if (Cfg::SMEM_BYTES > 48 * 1024) {
            if constexpr (BLKCAP) { if (!BLK_OWN_OK && (e = attr3<true, false>()) != cudaSuccess) return e; }
            if constexpr (FSCAP) { if ((e = attr3<false, true>()) != cudaSuccess) return e; }
            e = attr3<false, false>();
        }
        return e;
    }
    template <bool BLK, bool FS>
    static void go(int split, int inv, dim3 grid, dim3 block, size_t sm, cudaStream_t stream, const PassParams<T>& p) {
        if (split) tile_fft_kernel<Cfg, true, false, MINB, BLK, FS><<<grid, block, sm, stream>>>(p);
        else if (inv) tile_fft_kernel<Cfg, false, true, MINB, BLK, FS><<<grid, block, sm, stream>>>(p);
        else tile_fft_kernel<Cfg, false, false, MINB, BLK, FS><<<grid, block, sm, stream>>>(p);
    }
    static cudaError_t launch(int split, int inv, const void* params, cudaStream_t stream) {
        const PassParams<T>& p = *static_cast<const PassParams<T>*>(params);
        const long long ctas = (p.n_tiles + Cfg::G - 1) / Cfg::G;
        if (ctas <= 0) return cudaSuccess;
        if (ctas > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
        const dim3 grid((unsigned)((p.max_ctas > 0 && ctas > p.max_ctas) ? p.max_ctas : ctas)), block(Cfg::THREADS);
        const size_t sm = (size_t)Cfg::SMEM_BYTES;
        if (p.out_blk_log2 >= 0) {
            if constexpr (BLKCAP) {
                if (p.blk_bulk == 2 && BLK_OWN_OK) {
                    go<true, false>(split, inv, grid, block, BLK_SMEM, stream, p);
                } else if (p.blk_bulk == 2) {
                    PassParams<T> q = p;
                    q.blk_bulk = 1;
                    go<true, false>(split, inv, grid, block, sm, stream, q);
                } else {
                    go<true, false>(split, inv, grid, block, sm, stream, p);
                }
            } else return cudaErrorNotSupported;
        } else if (p.fs_t1 != nullptr) {
            if constexpr (FSCAP) go<false, true>(split, inv, grid, block, sm, stream, p);
            else return cudaErrorNotSupported;
        } else {
            go<false, false>(split, inv, grid, block, sm, stream, p);
        }
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
        v.fs = FSCAP ? 1 : 0;
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
    static int& slots() { static int s[B2_MAX_DEVICES] = {}; return s[b2_current_device()]; }   // resident CTAs (SMs x occupancy), per device   // resident CTAs on the device (SMs x occupancy)
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
        if (split) {
            // whole output planes of a group leave as bulk copies when the plain dense layout and the 16-byte rules allow it
            static const bool sb_ok = [] { const char* e = getenv("B2FFT_SPLIT_BULK"); return !e || atoi(e) != 0; }();
            PassParams<T> q = p;
            q.split_bulk = sb_ok && Cfg::S > 1 && p.out_blk_log2 < 0 && p.outer_div <= 0 && p.fs_t1 == nullptr && p.out_inner == 1 &&
                           p.out_outer_stride == Cfg::N && ((uintptr_t)p.out0 % 16) == 0 && ((uintptr_t)p.out1 % 16) == 0 &&
                           ((size_t)Cfg::N * sizeof(T)) % 16 == 0;
            tile_fft_kernel_tma_row<Cfg, true, false, MINB, NBUF><<<grid, block, sm, stream>>>(q);
        }
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

template <class Cfg, int MINB>
struct VariantOpsTmaRowAlias {
    using T = typename Cfg::T;
    using L = TmaRowAliasLayout<Cfg>;
    static int& slots() { static int s[B2_MAX_DEVICES] = {}; return s[b2_current_device()]; }
    static cudaError_t prepare() {
        const int b = (int)L::TOTAL;
        cudaError_t e;
        if ((e = cudaFuncSetAttribute(tile_fft_kernel_tma_row_alias<Cfg, false, false, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, b)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(tile_fft_kernel_tma_row_alias<Cfg, false, true, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, b)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(tile_fft_kernel_tma_row_alias<Cfg, true, false, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, b)) != cudaSuccess) return e;
        int dev = 0, sms = 0, occ = 0;
        if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
        if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tile_fft_kernel_tma_row_alias<Cfg, false, false, MINB>, Cfg::THREADS, L::TOTAL);
        if (e != cudaSuccess) return e;
        slots() = sms * (occ > 0 ? occ : 1);
        return cudaSuccess;
    }
    static cudaError_t launch(int split, int inv, const void* params, cudaStream_t stream) {
        const PassParams<T>& p = *static_cast<const PassParams<T>*>(params);
        if (p.out_blk_log2 >= 0 || p.outer_div > 0 || p.fs_t1 != nullptr || p.progress != nullptr ||
            (p.in_blk_log2 >= 0 && p.in_blk[0] != nullptr))
            return cudaErrorNotSupported;
        long long ctas = (p.n_tiles + Cfg::G - 1) / Cfg::G;
        if (ctas <= 0) return cudaSuccess;
        if (slots() <= 0) { cudaError_t e = prepare(); if (e != cudaSuccess) return e; }
        if (ctas > slots()) ctas = slots();
        const dim3 grid((unsigned)ctas), block(Cfg::THREADS);
        const size_t sm = L::TOTAL;
        if (split) tile_fft_kernel_tma_row_alias<Cfg, true, false, MINB><<<grid, block, sm, stream>>>(p);
        else if (inv) tile_fft_kernel_tma_row_alias<Cfg, false, true, MINB><<<grid, block, sm, stream>>>(p);
        else tile_fft_kernel_tma_row_alias<Cfg, false, false, MINB><<<grid, block, sm, stream>>>(p);
        return cudaGetLastError();
    }
    static int occupancy() {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, tile_fft_kernel_tma_row_alias<Cfg, false, false, MINB>, Cfg::THREADS, L::TOTAL) != cudaSuccess) return -1;
        return n;
    }
    static KernelVariant make(const char* name) {
        KernelVariant v = VariantOps<Cfg, MINB, false>::make(name);
        v.smem_bytes = (long long)L::TOTAL;
        v.kind = 1;
        v.nbuf = 0;
        v.launch = &launch;
        v.prepare = &prepare;
        v.occupancy = &occupancy;
        return v;
    }
};

// ---- host side of the tensor-map variants
typedef CUresult (*b2_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                       const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                       CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline b2_encode_tiled_fn b2_get_encode_tiled() {
    static b2_encode_tiled_fn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return (b2_encode_tiled_fn)f;
    }();
    return fn;
}

// {inner (x elems_per_complex), N, outer} view of one plane pointer, boxes {W (x epc), NB, 1}
// Encoded maps are cached per (pointer, geometry): a plan that is executed repeatedly on the same buffers -- the normal
// case -- pays for cuTensorMapEncodeTiled once, not once per launch.
struct B2MapKey {
    const void* base; long long inner, outer; int epc, N, W, NB, esz, dev;
    bool operator<(const B2MapKey& o) const {
        return std::tie(base, inner, outer, epc, N, W, NB, esz, dev) < std::tie(o.base, o.inner, o.outer, o.epc, o.N, o.W, o.NB, o.esz, o.dev);
    }
};
inline std::map<B2MapKey, CUtensorMap>& b2_map_cache() { static std::map<B2MapKey, CUtensorMap> m; return m; }
inline std::mutex& b2_map_mutex() { static std::mutex m; return m; }

template <typename T>
inline cudaError_t b2_make_col_map(CUtensorMap* tm, const void* base, int epc, long long inner, int N, long long outer,
                                   int W, int NB) {
    b2_encode_tiled_fn enc = b2_get_encode_tiled();
    if (!enc) return cudaErrorNotSupported;
    const B2MapKey key{base, inner, outer, epc, N, W, NB, (int)sizeof(T), b2_current_device()};
    {
        std::lock_guard<std::mutex> lk(b2_map_mutex());
        auto it = b2_map_cache().find(key);
        if (it != b2_map_cache().end()) { *tm = it->second; return cudaSuccess; }
    }
    const cuuint64_t dims[3] = {(cuuint64_t)(inner * epc), (cuuint64_t)N, (cuuint64_t)outer};
    const cuuint64_t strides[2] = {(cuuint64_t)(inner * epc) * sizeof(T), (cuuint64_t)(inner * epc) * N * sizeof(T)};
    const cuuint32_t box[3] = {(cuuint32_t)(W * epc), (cuuint32_t)NB, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3,
                     const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
    std::lock_guard<std::mutex> lk(b2_map_mutex());
    if (b2_map_cache().size() >= 4096) b2_map_cache().clear();      // bounded: callers that stream through fresh buffers
    b2_map_cache()[key] = *tm;
    return cudaSuccess;
}

template <class Cfg, int MINB, int NBUF, bool FSCAP = false>
struct VariantOpsTmaCol {
    using T = typename Cfg::T;
    using L = TmaRowLayout<Cfg, NBUF>;
    static int& slots() { static int s[B2_MAX_DEVICES] = {}; return s[b2_current_device()]; }   // resident CTAs (SMs x occupancy), per device
    template <bool SPLIT, bool INV, bool BLK, bool FS = false>
    static cudaError_t attr() {
        return cudaFuncSetAttribute(tile_fft_kernel_tma_col<Cfg, SPLIT, INV, MINB, NBUF, BLK, FS>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::TOTAL);
    }
    static cudaError_t prepare() {
        cudaError_t e;
        if ((e = attr<false, false, false>()) != cudaSuccess) return e;
        if ((e = attr<false, true, false>()) != cudaSuccess) return e;
        if ((e = attr<true, false, false>()) != cudaSuccess) return e;
        if ((e = attr<false, false, true>()) != cudaSuccess) return e;
        if ((e = attr<false, true, true>()) != cudaSuccess) return e;
        if ((e = attr<true, false, true>()) != cudaSuccess) return e;
        if constexpr (FSCAP) {
            if ((e = attr<false, false, false, true>()) != cudaSuccess) return e;
            if ((e = attr<false, true, false, true>()) != cudaSuccess) return e;
            if ((e = attr<true, false, false, true>()) != cudaSuccess) return e;
        }
        int dev = 0, sms = 0, occ = 0;
        if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
        if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tile_fft_kernel_tma_col<Cfg, false, false, MINB, NBUF, false>,
                                                          Cfg::THREADS, L::TOTAL);
        if (e != cudaSuccess) return e;
        slots() = sms * (occ > 0 ? occ : 1);
        return b2_get_encode_tiled() ? cudaSuccess : cudaErrorNotSupported;
    }
    template <bool SPLIT, bool INV>
    static void go(bool blk, dim3 grid, dim3 block, size_t sm, cudaStream_t st, const PassParams<T>& p, const CUtensorMap& a,
                   const CUtensorMap& b, const CUtensorMap& oa, const CUtensorMap& ob) {
        if (blk) tile_fft_kernel_tma_col<Cfg, SPLIT, INV, MINB, NBUF, true><<<grid, block, sm, st>>>(p, a, b, oa, ob);
        else if (p.fs_t1 != nullptr) {
            if constexpr (FSCAP) tile_fft_kernel_tma_col<Cfg, SPLIT, INV, MINB, NBUF, false, true><<<grid, block, sm, st>>>(p, a, b, oa, ob);
        } else tile_fft_kernel_tma_col<Cfg, SPLIT, INV, MINB, NBUF, false><<<grid, block, sm, st>>>(p, a, b, oa, ob);
    }
    static cudaError_t launch(int split, int inv, const void* params, cudaStream_t stream) {
        const PassParams<T>& p = *static_cast<const PassParams<T>*>(params);
        long long ctas = (p.n_tiles + Cfg::G - 1) / Cfg::G;
        if (ctas <= 0) return cudaSuccess;
        if (slots() <= 0) { cudaError_t e = prepare(); if (e != cudaSuccess) return e; }
        if (ctas > slots()) ctas = slots();
        const long long outer = p.n_tiles / p.inner_blocks;
        constexpr int NB = TmaColBox<Cfg>::NB;
        alignas(64) CUtensorMap tm0, tm1;
        cudaError_t e = b2_make_col_map<T>(&tm0, p.in0, split ? 1 : 2, p.inner, Cfg::N, outer, Cfg::W, NB);
        if (e != cudaSuccess) return e;
        tm1 = tm0;
        if (split) {
            e = b2_make_col_map<T>(&tm1, p.in1, 1, p.inner, Cfg::N, outer, Cfg::W, NB);
            if (e != cudaSuccess) return e;
        }
        const dim3 grid((unsigned)ctas), block(Cfg::THREADS);
        const bool blk = p.out_blk_log2 >= 0;
        if (!FSCAP && p.fs_t1 != nullptr) return cudaErrorNotSupported;
        // output maps for the tensor-store path (plain passes only; same geometry as the input)
        alignas(64) CUtensorMap tmo0 = tm0, tmo1 = tm1;
        PassParams<T> q = p;
        q.tma_store = 0;
        // measured on B200 (profiles/r01_tma_store_fs_bulk.md): -3 % on N=1024 x W=8 tiles, +1 % on N=2048 x W=4 --
        // the extra barrier and shared-memory pass cancel the saved L1 wavefronts, so it is opt-in
        static const bool ts_ok = [] { const char* e = getenv("B2FFT_TMA_STORE"); return e && atoi(e) != 0; }();
        if (ts_ok && !blk && p.fs_t1 == nullptr && p.outer_div == 0 && p.out_inner == p.inner &&
            p.out_outer_stride == p.outer_stride && ((uintptr_t)p.out0 % 16) == 0 && (!split || ((uintptr_t)p.out1 % 16) == 0)) {
            q.tma_store = 1;
            if (p.out0 != p.in0) {
                e = b2_make_col_map<T>(&tmo0, p.out0, split ? 1 : 2, p.inner, Cfg::N, outer, Cfg::W, NB);
                if (e != cudaSuccess) return e;
            }
            if (split && p.out1 != p.in1) {
                e = b2_make_col_map<T>(&tmo1, p.out1, 1, p.inner, Cfg::N, outer, Cfg::W, NB);
                if (e != cudaSuccess) return e;
            }
        }
        if (split) go<true, false>(blk, grid, block, L::TOTAL, stream, q, tm0, tm1, tmo0, tmo1);
        else if (inv) go<false, true>(blk, grid, block, L::TOTAL, stream, q, tm0, tm1, tmo0, tmo1);
        else go<false, false>(blk, grid, block, L::TOTAL, stream, q, tm0, tm1, tmo0, tmo1);
        return cudaGetLastError();
    }
    static int occupancy() {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, tile_fft_kernel_tma_col<Cfg, false, false, MINB, NBUF, false>,
                                                          Cfg::THREADS, L::TOTAL) != cudaSuccess)
            return -1;
        return n;
    }
    static KernelVariant make(const char* name) {
        KernelVariant v = VariantOps<Cfg, MINB, false>::make(name);
        v.smem_bytes = (long long)L::TOTAL;
        v.kind = 2;
        v.nbuf = NBUF;
        v.blk = 1;
        v.fs = FSCAP ? 1 : 0;
        v.launch = &launch;
        v.prepare = &prepare;
        v.occupancy = &occupancy;
        return v;
    }
};

template <class Cfg, int MINB>
struct VariantOpsTmaColAlias {
    using T = typename Cfg::T;
    using L = TmaAliasLayout<Cfg>;
    static int& slots() { static int s[B2_MAX_DEVICES] = {}; return s[b2_current_device()]; }   // resident CTAs (SMs x occupancy), per device
    template <bool SPLIT, bool INV, bool BLK>
    static cudaError_t attr() {
        return cudaFuncSetAttribute(tile_fft_kernel_tma_col_alias<Cfg, SPLIT, INV, MINB, BLK>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::TOTAL);
    }
    static cudaError_t prepare() {
        cudaError_t e;
        if ((e = attr<false, false, false>()) != cudaSuccess) return e;
        if ((e = attr<false, true, false>()) != cudaSuccess) return e;
        if ((e = attr<true, false, false>()) != cudaSuccess) return e;
        if ((e = attr<false, false, true>()) != cudaSuccess) return e;
        if ((e = attr<false, true, true>()) != cudaSuccess) return e;
        if ((e = attr<true, false, true>()) != cudaSuccess) return e;
        int dev = 0, sms = 0, occ = 0;
        if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
        if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tile_fft_kernel_tma_col_alias<Cfg, false, false, MINB, false>,
                                                          Cfg::THREADS, L::TOTAL);
        if (e != cudaSuccess) return e;
        slots() = sms * (occ > 0 ? occ : 1);
        return b2_get_encode_tiled() ? cudaSuccess : cudaErrorNotSupported;
    }
    template <bool SPLIT, bool INV>
    static void go(bool blk, dim3 grid, dim3 block, cudaStream_t st, const PassParams<T>& p, const CUtensorMap& a,
                   const CUtensorMap& b) {
        if (blk) tile_fft_kernel_tma_col_alias<Cfg, SPLIT, INV, MINB, true><<<grid, block, L::TOTAL, st>>>(p, a, b);
        else tile_fft_kernel_tma_col_alias<Cfg, SPLIT, INV, MINB, false><<<grid, block, L::TOTAL, st>>>(p, a, b);
    }
    static cudaError_t launch(int split, int inv, const void* params, cudaStream_t stream) {
        const PassParams<T>& p = *static_cast<const PassParams<T>*>(params);
        long long ctas = (p.n_tiles + Cfg::G - 1) / Cfg::G;
        if (ctas <= 0) return cudaSuccess;
        if (p.fs_t1 != nullptr) return cudaErrorNotSupported;
        if (slots() <= 0) { cudaError_t e = prepare(); if (e != cudaSuccess) return e; }
        if (ctas > slots()) ctas = slots();
        const long long outer = p.n_tiles / p.inner_blocks;
        constexpr int NB = TmaColBox<Cfg>::NB;
        alignas(64) CUtensorMap tm0, tm1;
        cudaError_t e = b2_make_col_map<T>(&tm0, p.in0, split ? 1 : 2, p.inner, Cfg::N, outer, Cfg::W, NB);
        if (e != cudaSuccess) return e;
        tm1 = tm0;
        if (split) {
            e = b2_make_col_map<T>(&tm1, p.in1, 1, p.inner, Cfg::N, outer, Cfg::W, NB);
            if (e != cudaSuccess) return e;
        }
        const dim3 grid((unsigned)ctas), block(Cfg::THREADS);
        const bool blk = p.out_blk_log2 >= 0;
        if (split) go<true, false>(blk, grid, block, stream, p, tm0, tm1);
        else if (inv) go<false, true>(blk, grid, block, stream, p, tm0, tm1);
        else go<false, false>(blk, grid, block, stream, p, tm0, tm1);
        return cudaGetLastError();
    }
    static int occupancy() {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, tile_fft_kernel_tma_col_alias<Cfg, false, false, MINB, false>,
                                                          Cfg::THREADS, L::TOTAL) != cudaSuccess)
            return -1;
        return n;
    }
    static KernelVariant make(const char* name) {
        KernelVariant v = VariantOps<Cfg, MINB, false>::make(name);
        v.smem_bytes = (long long)L::TOTAL;
        v.kind = 2;
        v.nbuf = 0;
        v.blk = 1;
        v.fs = 0;
        v.launch = &launch;
        v.prepare = &prepare;
        v.occupancy = &occupancy;
        return v;
    }
};


template <class CfgA, class CfgB, int MINB>
struct VariantOpsFused2 {
    using T = typename CfgA::T;
    static constexpr size_t SMEM = (size_t)(CfgA::SMEM_BYTES > CfgB::SMEM_BYTES ? CfgA::SMEM_BYTES : CfgB::SMEM_BYTES);
    static int& slots() { static int s[B2_MAX_DEVICES] = {}; return s[b2_current_device()]; }   // resident CTAs (SMs x occupancy), per device
    static cudaError_t prepare() {
        cudaError_t e;
        if ((e = cudaFuncSetAttribute(fused2_fft_kernel<CfgA, CfgB, false, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)SMEM)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(fused2_fft_kernel<CfgA, CfgB, true, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)SMEM)) != cudaSuccess) return e;
        int dev = 0, sms = 0, occ = 0;
        if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
        if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fused2_fft_kernel<CfgA, CfgB, false, MINB>, CfgA::THREADS, SMEM);
        if (e != cudaSuccess) return e;
        slots() = sms * (occ > 0 ? occ : 1);
        return cudaSuccess;
    }
    static int grid_slots() {
        if (slots() <= 0 && prepare() != cudaSuccess) return -1;
        return slots();
    }
    static cudaError_t launch(int split, int inv, const void* params, cudaStream_t stream) {
        const PassParams<T>& p = *static_cast<const PassParams<T>*>(params);
        if (split || p.out_blk_log2 >= 0 || p.outer_div > 0 || !p.scratch || !p.fs_t2) return cudaErrorNotSupported;
        if (p.progress != nullptr) return cudaErrorNotSupported;
        if (p.n_tiles <= 0) return cudaSuccess;
        if (slots() <= 0) { cudaError_t e = prepare(); if (e != cudaSuccess) return e; }
        long long ctas = p.n_tiles;
        if (ctas > slots()) ctas = slots();
        if (ctas > p.scratch_slots) ctas = p.scratch_slots;
        if (p.max_ctas > 0 && ctas > p.max_ctas) ctas = p.max_ctas;
        if (ctas <= 0) return cudaErrorInvalidValue;
        PassParams<T> pa = p, pb = p;
        pa.inner = (long long)CfgB::N * p.inner;          // n1 stride
        pa.out0 = p.scratch;
        pa.out_inner = CfgA::W;
        pa.scale_mode = 0;
        pa.fs_n2 = CfgB::N;
        pb.in0 = p.scratch;
        pb.inner = (long long)CfgA::N * CfgA::W;          // n2 stride inside the scratch slot
        pb.out_inner = (long long)CfgA::N * p.out_inner;  // k2 stride in the output
        for (int s = 0; s < 3; ++s) pb.tw[s] = p.tw_b[s];
        pb.fs_t1 = pb.fs_t2 = nullptr;
        static const int flags = [] { const char* e = getenv("B2FFT_FUSED_FLAGS"); return e ? atoi(e) : 13; }();
        // discard needs 128-byte aligned scratch rows (W * sizeof(complex) == 128 and a 128-byte aligned buffer)
        pa.fused_flags = pb.fused_flags = (((uintptr_t)p.scratch % 128) == 0 && CfgA::W * 2 * sizeof(T) == 128) ? flags : (flags & ~1);
        const dim3 grid((unsigned)ctas), block(CfgA::THREADS);
        if (inv) fused2_fft_kernel<CfgA, CfgB, true, MINB><<<grid, block, SMEM, stream>>>(pa, pb, p.inner, p.out_inner);
        else fused2_fft_kernel<CfgA, CfgB, false, MINB><<<grid, block, SMEM, stream>>>(pa, pb, p.inner, p.out_inner);
        return cudaGetLastError();
    }
    static int occupancy() {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fused2_fft_kernel<CfgA, CfgB, false, MINB>, CfgA::THREADS, SMEM) !=
            cudaSuccess)
            return -1;
        return n;
    }
    static KernelVariant make(const char* name) {
        KernelVariant v{};
        v.name = name;
        v.prec = sizeof(T) == 4 ? 0 : 1;
        v.log2n = CfgA::LOG2N + CfgB::LOG2N;
        v.log2n1 = CfgA::LOG2N;
        v.W = CfgA::W; v.G = CfgA::G; v.E = CfgA::E; v.S = CfgA::S;
        v.S_b = CfgB::S; v.E_b = CfgB::E;
        for (int s = 0; s < 4; ++s) { v.radix[s] = s < CfgA::S ? CfgA::R(s) : 1; v.radix_b[s] = s < CfgB::S ? CfgB::R(s) : 1; }
        v.threads = CfgA::THREADS;
        v.smem_bytes = (long long)SMEM;
        v.minb = MINB;
        v.kind = 3;
        v.slot_elems = (long long)CfgA::N * CfgB::N * CfgA::W;
        v.launch = &launch;
        v.prepare = &prepare;
        v.occupancy = &occupancy;
        v.grid_slots = &grid_slots;
        return v;
    }
};


// P: the streamed in-place kernel (fused2p_fft_kernel) instead of fused2s_fft_kernel; same parameters, same scratch slot
template <class CfgA, class CfgB, int KS, bool P = false>
struct VariantOpsFused2S {
    using T = typename CfgA::T;
    using KernFn = void (*)(const PassParams<T>, const PassParams<T>, const long long, const long long);
    template <bool INV>
    static KernFn kern() {
        if constexpr (P) return &fused2p_fft_kernel<CfgA, CfgB, KS, INV>;
        else return &fused2s_fft_kernel<CfgA, CfgB, KS, INV>;
    }
    static constexpr size_t SMEM = (size_t)KS * CfgB::N * CfgA::W * 2 * sizeof(T);
    static_assert(SMEM <= 227 * 1024, "intermediate rows kept in shared memory must fit");
    static int& slots() { static int s[B2_MAX_DEVICES] = {}; return s[b2_current_device()]; }   // resident CTAs (SMs x occupancy), per device
    static cudaError_t prepare() {
        cudaError_t e;
        if ((e = cudaFuncSetAttribute(kern<false>(), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)SMEM)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(kern<true>(), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)SMEM)) != cudaSuccess) return e;
        int dev = 0, sms = 0, occ = 0;
        if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
        if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern<false>(), CfgA::THREADS, SMEM);
        if (e != cudaSuccess) return e;
        slots() = sms * (occ > 0 ? occ : 1);
        return cudaSuccess;
    }
    static int grid_slots() {
        if (slots() <= 0 && prepare() != cudaSuccess) return -1;
        return slots();
    }
    static cudaError_t launch(int split, int inv, const void* params, cudaStream_t stream) {
        const PassParams<T>& p = *static_cast<const PassParams<T>*>(params);
        if (split || p.out_blk_log2 >= 0 || p.outer_div > 0 || !p.scratch || !p.fs_t2) return cudaErrorNotSupported;
        if (p.progress != nullptr && (!P || p.progress_tiles <= 0)) return cudaErrorNotSupported;
        if (p.n_tiles <= 0) return cudaSuccess;
        if (slots() <= 0) { cudaError_t e = prepare(); if (e != cudaSuccess) return e; }
        long long ctas = p.n_tiles;
        if (ctas > slots()) ctas = slots();
        if (ctas > p.scratch_slots) ctas = p.scratch_slots;
        if (p.max_ctas > 0 && ctas > p.max_ctas) ctas = p.max_ctas;
        if (ctas <= 0) return cudaErrorInvalidValue;
        PassParams<T> pa = p, pb = p;
        pa.inner = (long long)CfgB::N * p.inner;          // n1 stride
        pa.out0 = p.scratch;
        pa.scale_mode = 0;
        pa.fs_n2 = CfgB::N;
        pb.in0 = p.scratch;
        pb.out_inner = (long long)CfgA::N * p.out_inner;  // k2 stride in the output
        pb.fs_t1 = pb.fs_t2 = nullptr;
        const dim3 grid((unsigned)ctas), block(CfgA::THREADS);
        if (inv) (*kern<true>())<<<grid, block, SMEM, stream>>>(pa, pb, p.inner, p.out_inner);
        else (*kern<false>())<<<grid, block, SMEM, stream>>>(pa, pb, p.inner, p.out_inner);
        return cudaGetLastError();
    }
    static int occupancy() {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern<false>(), CfgA::THREADS, SMEM) !=
            cudaSuccess)
            return -1;
        return n;
    }
    static KernelVariant make(const char* name) {
        KernelVariant v{};
        v.name = name;
        v.prec = sizeof(T) == 4 ? 0 : 1;
        v.log2n = CfgA::LOG2N + CfgB::LOG2N;
        v.log2n1 = CfgA::LOG2N;
        v.W = CfgA::W; v.G = CfgA::G; v.E = CfgA::E; v.S = 1;
        v.S_b = 1; v.E_b = CfgB::E;
        for (int s = 0; s < 4; ++s) { v.radix[s] = s == 0 ? CfgA::E : 1; v.radix_b[s] = s == 0 ? CfgB::E : 1; }
        v.threads = CfgA::THREADS;
        v.smem_bytes = (long long)SMEM;
        v.minb = 1;
        v.kind = 3;
        v.slot_elems = (long long)(CfgA::N - KS) * CfgB::N * CfgA::W;
        if (v.slot_elems == 0) v.slot_elems = CfgA::W;          // keep a (dummy) scratch allocation so that launch() runs
        v.progress = P ? 1 : 0;
        v.launch = &launch;
        v.prepare = &prepare;
        v.occupancy = &occupancy;
        v.grid_slots = &grid_slots;
        return v;
    }
};


template <int LOG2A, int LOG2B, int KS>
struct VariantOpsFused2W {
    using T = float;
    static constexpr int N1 = 1 << LOG2A, N2 = 1 << LOG2B, W = 16;
    static constexpr size_t SMEM = (size_t)KS * N2 * W * 2 * sizeof(T);
    static_assert(SMEM <= 227 * 1024, "intermediate rows kept in shared memory must fit");
    static int& slots() { static int s[B2_MAX_DEVICES] = {}; return s[b2_current_device()]; }
    static cudaError_t prepare() {
        cudaError_t e;
        if ((e = cudaFuncSetAttribute(fused2w_fft_kernel<LOG2A, LOG2B, KS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)SMEM)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(fused2w_fft_kernel<LOG2A, LOG2B, KS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)SMEM)) != cudaSuccess) return e;
        int dev = 0, sms = 0, occ = 0;
        if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
        if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fused2w_fft_kernel<LOG2A, LOG2B, KS, false>, 512, SMEM);
        if (e != cudaSuccess) return e;
        slots() = sms * (occ > 0 ? occ : 1);
        return cudaSuccess;
    }
    static int grid_slots() {
        if (slots() <= 0 && prepare() != cudaSuccess) return -1;
        return slots();
    }
    static cudaError_t launch(int split, int inv, const void* params, cudaStream_t stream) {
        const PassParams<T>& p = *static_cast<const PassParams<T>*>(params);
        if (split || p.out_blk_log2 >= 0 || p.outer_div > 0 || !p.scratch || !p.fs_t2) return cudaErrorNotSupported;
        if (p.n_tiles <= 0) return cudaSuccess;
        if (slots() <= 0) { cudaError_t e = prepare(); if (e != cudaSuccess) return e; }
        long long ctas = p.n_tiles;
        if (ctas > slots()) ctas = slots();
        if (ctas > p.scratch_slots) ctas = p.scratch_slots;
        if (p.max_ctas > 0 && ctas > p.max_ctas) ctas = p.max_ctas;
        if (ctas <= 0) return cudaErrorInvalidValue;
        PassParams<T> pa = p, pb = p;
        pa.inner = (long long)N2 * p.inner;               // n1 stride
        pa.out0 = p.scratch;
        pb.out_inner = (long long)N1 * p.out_inner;       // k2 stride in the output
        const dim3 grid((unsigned)ctas), block(512);
        if (inv) fused2w_fft_kernel<LOG2A, LOG2B, KS, true><<<grid, block, SMEM, stream>>>(pa, pb, p.inner, p.out_inner);
        else fused2w_fft_kernel<LOG2A, LOG2B, KS, false><<<grid, block, SMEM, stream>>>(pa, pb, p.inner, p.out_inner);
        return cudaGetLastError();
    }
    static int occupancy() {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fused2w_fft_kernel<LOG2A, LOG2B, KS, false>, 512, SMEM) != cudaSuccess) return -1;
        return n;
    }
    static KernelVariant make(const char* name) {
        KernelVariant v{};
        v.name = name;
        v.prec = 0;
        v.log2n = LOG2A + LOG2B;
        v.log2n1 = LOG2A;
        v.W = W; v.G = 512 / (2 * W); v.E = N1; v.S = 1;       // E = N1: the inter-step twiddle table is [k1][N2]
        v.S_b = 1; v.E_b = N2;
        for (int s = 0; s < 4; ++s) { v.radix[s] = s == 0 ? N1 : 1; v.radix_b[s] = s == 0 ? N2 : 1; }
        v.threads = 512;
        v.smem_bytes = (long long)SMEM;
        v.minb = 1;
        v.kind = 3;
        v.slot_elems = (long long)(N1 - KS) * N2 * W;
        if (v.slot_elems == 0) v.slot_elems = W;
        v.launch = &launch;
        v.prepare = &prepare;
        v.occupancy = &occupancy;
        v.grid_slots = &grid_slots;
        return v;
    }
};

template <int LOG2N>
struct VariantOpsShfl {
    static int& sms() { static int s[B2_MAX_DEVICES] = {}; return s[b2_current_device()]; }
    static cudaError_t prepare() {
        int dev = 0, n = 0;
        cudaError_t e;
        if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
        if ((e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
        sms() = n;
        return cudaSuccess;
    }
    static cudaError_t launch(int split, int inv, const void* params, cudaStream_t stream) {
        const PassParams<float>& p = *static_cast<const PassParams<float>*>(params);
        if (p.out_blk_log2 >= 0 || p.outer_div > 0 || p.fs_t1 != nullptr || p.inner != 1 || p.progress != nullptr ||
            (p.in_blk_log2 >= 0 && p.in_blk[0] != nullptr))
            return cudaErrorNotSupported;
        if (p.n_tiles <= 0) return cudaSuccess;
        if (sms() <= 0) { cudaError_t e = prepare(); if (e != cudaSuccess) return e; }
        constexpr int LP = (1 << LOG2N) / 2;
        long long ctas = (p.n_tiles * LP + 4 * 256 - 1) / (4 * 256);      // four rows per thread and step
        const long long cap = (long long)sms() * 32;
        if (ctas > cap) ctas = cap;
        const dim3 grid((unsigned)ctas), block(256);
        if (split) row_shfl_kernel<LOG2N, true, false><<<grid, block, 0, stream>>>(p);
        else if (inv) row_shfl_kernel<LOG2N, false, true><<<grid, block, 0, stream>>>(p);
        else row_shfl_kernel<LOG2N, false, false><<<grid, block, 0, stream>>>(p);
        return cudaGetLastError();
    }
    static int occupancy() {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, row_shfl_kernel<LOG2N, false, false>, 256, 0) != cudaSuccess) return -1;
        return n;
    }
    static KernelVariant make(const char* name) {
        KernelVariant v{};
        v.name = name;
        v.prec = 0;
        v.log2n = LOG2N;
        v.W = 1; v.G = 512 / (1 << LOG2N); v.E = 2; v.S = 1;
        for (int s = 0; s < 4; ++s) v.radix[s] = s == 0 ? (1 << LOG2N) : 1;
        v.threads = 256;
        v.smem_bytes = 0;
        v.minb = 1;
        v.kind = 4;
        v.launch = &launch;
        v.prepare = &prepare;
        v.occupancy = &occupancy;
        return v;
    }
};

template <int LOG2N>
struct VariantOpsShfl4 {
    static int& sms() { static int s[B2_MAX_DEVICES] = {}; return s[b2_current_device()]; }
    static cudaError_t prepare() { return VariantOpsShfl<2>::prepare(); }
    static cudaError_t launch(int split, int inv, const void* params, cudaStream_t stream) {
        const PassParams<float>& p = *static_cast<const PassParams<float>*>(params);
        if (p.out_blk_log2 >= 0 || p.outer_div > 0 || p.fs_t1 != nullptr || p.inner != 1 || p.progress != nullptr ||
            (p.in_blk_log2 >= 0 && p.in_blk[0] != nullptr))
            return cudaErrorNotSupported;
        // 16-byte vectors on both sides: the output planes / array must be 16-byte aligned as well (the planner's alignment rule
        // covers the input only)
        if (((uintptr_t)p.out0 % 16) != 0 || (split && ((uintptr_t)p.out1 % 16) != 0)) return cudaErrorNotSupported;
        if (p.n_tiles <= 0) return cudaSuccess;
        if (VariantOpsShfl<2>::sms() <= 0) { cudaError_t e = prepare(); if (e != cudaSuccess) return e; }
        constexpr int LP = (1 << LOG2N) / 4;
        long long ctas = (p.n_tiles * LP + 2 * 256 - 1) / (2 * 256);      // two rows per thread and step
        const long long cap = (long long)VariantOpsShfl<2>::sms() * 32;
        if (ctas > cap) ctas = cap;
        const dim3 grid((unsigned)ctas), block(256);
        if (split) row_shfl4_kernel<LOG2N, true, false><<<grid, block, 0, stream>>>(p);
        else if (inv) row_shfl4_kernel<LOG2N, false, true><<<grid, block, 0, stream>>>(p);
        else row_shfl4_kernel<LOG2N, false, false><<<grid, block, 0, stream>>>(p);
        return cudaGetLastError();
    }
    static int occupancy() {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, row_shfl4_kernel<LOG2N, false, false>, 256, 0) != cudaSuccess) return -1;
        return n;
    }
    static KernelVariant make(const char* name) {
        KernelVariant v = VariantOpsShfl<LOG2N < 6 ? LOG2N : 6>::make(name);
        v.log2n = LOG2N;
        v.G = 1024 / (1 << LOG2N); v.E = 4;
        v.radix[0] = 1 << LOG2N;
        v.launch = &launch;
        v.prepare = &prepare;
        v.occupancy = &occupancy;
        return v;
    }
};
#define B2_VR4(L) out.push_back(::b2::VariantOpsShfl4<L>::make("float_n" #L "_w1_shfl4"));

// B2_VR(log2n): short-row kernel, complex64 / split float32, 16-byte accesses + warp-shuffle exchanges
#define B2_VR(L) out.push_back(::b2::VariantOpsShfl<L>::make("float_n" #L "_w1_shfl"));

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

// B2_VF: like B2_VB, additionally compiled as a four-step "A" pass (transposed store + twiddle)
#define B2_VF(T, L, W, G, MB, R0, R1, R2, R3)                                                        \
    out.push_back(::b2::VariantOps<::b2::TileCfg<T, L, W, G, R0, R1, R2, R3>, MB, true, true>::make(   \
        #T "_n" B2_STR(L) "_w" B2_STR(W) "_g" B2_STR(G) "_b" B2_STR(MB) "_r" B2_STR(R0) "x" B2_STR(R1) "x" B2_STR(R2) "x" B2_STR(R3)));

// B2_VT(type, log2n, G, minblocks, ring depth, R0, R1, R2, R3): persistent TMA-staged contiguous-axis variant
#define B2_VT(T, L, G, MB, NB, R0, R1, R2, R3)                                                       \
    out.push_back(::b2::VariantOpsTma<::b2::TileCfg<T, L, 1, G, R0, R1, R2, R3>, MB, NB>::make(        \
        #T "_n" B2_STR(L) "_w1_g" B2_STR(G) "_b" B2_STR(MB) "_r" B2_STR(R0) "x" B2_STR(R1) "x" B2_STR(R2) "x" B2_STR(R3) "_tma" B2_STR(NB)));

// B2_VTA(type, log2n, G, minblocks, R0..R3): persistent TMA-staged contiguous-axis variant whose staging slot is the exchange buffer
#define B2_VTA(T, L, G, MB, R0, R1, R2, R3)                                                          \
    out.push_back(::b2::VariantOpsTmaRowAlias<::b2::TileCfg<T, L, 1, G, R0, R1, R2, R3>, MB>::make(    \
        #T "_n" B2_STR(L) "_w1_g" B2_STR(G) "_b" B2_STR(MB) "_r" B2_STR(R0) "x" B2_STR(R1) "x" B2_STR(R2) "x" B2_STR(R3) "_tmara"));

// B2_VC(type, log2n, W, G, minblocks, ring depth, R0..R3): persistent TMA tensor-staged strided-axis variant
#define B2_VC(T, L, W, G, MB, NB, R0, R1, R2, R3)                                                    \
    out.push_back(::b2::VariantOpsTmaCol<::b2::TileCfg<T, L, W, G, R0, R1, R2, R3>, MB, NB>::make(     \
        #T "_n" B2_STR(L) "_w" B2_STR(W) "_g" B2_STR(G) "_b" B2_STR(MB) "_r" B2_STR(R0) "x" B2_STR(R1) "x" B2_STR(R2) "x" B2_STR(R3) "_tmac" B2_STR(NB)));

// B2_VCF: like B2_VC, additionally compiled as a four-step "A" pass
#define B2_VCF(T, L, W, G, MB, NB, R0, R1, R2, R3)                                                   \
    out.push_back(::b2::VariantOpsTmaCol<::b2::TileCfg<T, L, W, G, R0, R1, R2, R3>, MB, NB, true>::make( \
        #T "_n" B2_STR(L) "_w" B2_STR(W) "_g" B2_STR(G) "_b" B2_STR(MB) "_r" B2_STR(R0) "x" B2_STR(R1) "x" B2_STR(R2) "x" B2_STR(R3) "_tmac" B2_STR(NB)));

// B2_V0 / B2_VT0 / B2_VC0: tuning variants with the stage twiddles loaded in full (TileCfg::TWP = 0)
#define B2_V0(T, L, W, G, MB, R0, R1, R2, R3)                                                        \
    out.push_back(::b2::VariantOps<::b2::TileCfg<T, L, W, G, R0, R1, R2, R3, 0>, MB>::make(            \
        #T "_n" B2_STR(L) "_w" B2_STR(W) "_g" B2_STR(G) "_b" B2_STR(MB) "_r" B2_STR(R0) "x" B2_STR(R1) "x" B2_STR(R2) "x" B2_STR(R3) "_tw0"));
#define B2_VT0(T, L, G, MB, NB, R0, R1, R2, R3)                                                      \
    out.push_back(::b2::VariantOpsTma<::b2::TileCfg<T, L, 1, G, R0, R1, R2, R3, 0>, MB, NB>::make(     \
        #T "_n" B2_STR(L) "_w1_g" B2_STR(G) "_b" B2_STR(MB) "_r" B2_STR(R0) "x" B2_STR(R1) "x" B2_STR(R2) "x" B2_STR(R3) "_tma" B2_STR(NB) "_tw0"));
#define B2_VC0(T, L, W, G, MB, NB, R0, R1, R2, R3)                                                   \
    out.push_back(::b2::VariantOpsTmaCol<::b2::TileCfg<T, L, W, G, R0, R1, R2, R3, 0>, MB, NB>::make(  \
        #T "_n" B2_STR(L) "_w" B2_STR(W) "_g" B2_STR(G) "_b" B2_STR(MB) "_r" B2_STR(R0) "x" B2_STR(R1) "x" B2_STR(R2) "x" B2_STR(R3) "_tmac" B2_STR(NB) "_tw0"));

// B2_VCA(type, log2n, W, G, minblocks, R0..R3): TMA-staged strided variant whose staging slot is the exchange buffer
#define B2_VCA(T, L, W, G, MB, R0, R1, R2, R3)                                                       \
    out.push_back(::b2::VariantOpsTmaColAlias<::b2::TileCfg<T, L, W, G, R0, R1, R2, R3>, MB>::make(    \
        #T "_n" B2_STR(L) "_w" B2_STR(W) "_g" B2_STR(G) "_b" B2_STR(MB) "_r" B2_STR(R0) "x" B2_STR(R1) "x" B2_STR(R2) "x" B2_STR(R3) "_tmaca"));

// B2_VU(type, log2n1, log2n2, W, GA, GB, minblocks, A radices R0..R3, B radices R0..R3): fused two-step strided variant
#define B2_VU(T, L1, L2, W, GA, GB, MB, A0, A1, A2, A3, B0, B1, B2_, B3)                              \
    out.push_back(::b2::VariantOpsFused2<::b2::TileCfg<T, L1, W, GA, A0, A1, A2, A3>,                  \
                                         ::b2::TileCfg<T, L2, W, GB, B0, B1, B2_, B3>, MB>::make(      \
        #T "_n" B2_STR(L1) "+" B2_STR(L2) "_w" B2_STR(W) "_g" B2_STR(GA) "+" B2_STR(GB) "_b" B2_STR(MB) "_r" B2_STR(A0) "x" B2_STR(A1) "x" B2_STR(A2) "+" B2_STR(B0) "x" B2_STR(B1) "x" B2_STR(B2_) "_fused2"));

// B2_VS(type, log2n1, log2n2, W, GA, GB, KS): fused two-step strided variant with the intermediate in shared memory; step A
// is a radix-2^log2n1 register FFT per thread (GA values of n2 per sub-tile), step B a radix-2^log2n2 one (GB values of k1)
#define B2_VS(T, L1, L2, W, GA, GB, KS)                                                              \
    out.push_back(::b2::VariantOpsFused2S<::b2::TileCfg<T, L1, W, GA, (1 << L1)>,                     \
                                          ::b2::TileCfg<T, L2, W, GB, (1 << L2)>, KS>::make(          \
        #T "_n" B2_STR(L1) "+" B2_STR(L2) "_w" B2_STR(W) "_g" B2_STR(GA) "+" B2_STR(GB) "_ks" B2_STR(KS) "_fused2s"));

// B2_VP: like B2_VS with the streamed in-place kernel (asynchronous refill of the shared-memory tile during step B)
#define B2_VP(T, L1, L2, W, GA, GB, KS)                                                              \
    out.push_back(::b2::VariantOpsFused2S<::b2::TileCfg<T, L1, W, GA, (1 << L1)>,                     \
                                          ::b2::TileCfg<T, L2, W, GB, (1 << L2)>, KS, true>::make(    \
        #T "_n" B2_STR(L1) "+" B2_STR(L2) "_w" B2_STR(W) "_g" B2_STR(GA) "+" B2_STR(GB) "_ks" B2_STR(KS) "_fused2p"));

// B2_VW(log2n1, log2n2, KS): fused two-step strided variant (complex64, W = 16, 512 threads) whose steps are lane-pair FFTs
// with a warp-shuffle exchange; KS rows of the intermediate in shared memory
#define B2_VW(L1, L2, KS)                                                                            \
    out.push_back(::b2::VariantOpsFused2W<L1, L2, KS>::make(                                          \
        "float_n" B2_STR(L1) "+" B2_STR(L2) "_w16_pair_ks" B2_STR(KS) "_fused2w"));

void register_f32_row(std::vector<KernelVariant>& out);
void register_f32_col(std::vector<KernelVariant>& out);
void register_f64_row(std::vector<KernelVariant>& out);
void register_f64_col(std::vector<KernelVariant>& out);
void register_exp(std::vector<KernelVariant>& out);
void register_exp2(std::vector<KernelVariant>& out);
void register_fused(std::vector<KernelVariant>& out);
void register_fused_exp(std::vector<KernelVariant>& out);

}  // namespace b2
