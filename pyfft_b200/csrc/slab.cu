// Slab-decomposed 3-D C2C transform over several GPUs: one rank's schedule, behind the C ABI
// (include/b2fft.h, b2fft_slab_*).  No reference counterpart: the reference is single-device
// (pyfft/plan.py:204-245 runs every kernel of a plan on one stream of one context).
//
// Layout.  Rank g holds z in [g*Zl, (g+1)*Zl) of a (Z, Y, X) array as the z-slab [Zl][Y][X].  After the
// forward transform rank h holds x in [h*Xb, (h+1)*Xb) as the x-slab [Y][Z][Xb] ("transposed out": one
// exchange, the result stays distributed).
//
// Forward schedule (all on the device, no host synchronisation, no NCCL on the data path):
//   main stream   : "my x-slab may be overwritten" -> ready flag of this rank on every peer
//                   for each z-chunk k:  Y pass of chunk k (local, in place)
//   exchange strm : for each z-chunk k (after its Y pass), for each y-chunk c:
//                       X pass of rows {z in chunk k} x {y in chunk c}; its stores are destination-blocked:
//                       x-block h of every output row goes, as one contiguous X/G-element piece, straight into
//                       rank h's x-slab over NVLink (cp.async.bulk from shared memory or 256-byte warp stores)
//                   after the last z-chunk of y-chunk c: flag (rank, c) on every peer (st.release.sys)
//   Z stream      : for each y-chunk c: spin on the flags (all ranks, c) in local memory (ld.acquire.sys), then
//                   the Z pass of that chunk's rows (row pitch Xb) -- it runs beside the NVLink stores of the
//                   following chunks.
// The cross-rank synchronisation is a word per (rank, chunk) in peer-mapped device memory carrying a
// monotonically increasing epoch; nothing waits on the host and no collective library is involved.
//
// Inverse schedule (x-slabs -> z-slabs): inverse Z pass per y-chunk in place, flag per chunk, then the X pass
// of every (z, y) row PULLS its G pieces from the peers' x-slabs with TMA bulk loads (source-blocked loads,
// the mirror image of the forward stores), then the local inverse Y pass applies the scale.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/b2fft.h"

namespace {

int slab_fail(int code, const char* fmt, ...);

__global__ void slab_signal_kernel(unsigned* const* peer_flags, int n_peers, int word, unsigned value) {
    const int h = (int)threadIdx.x;
    if (h < n_peers) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer_flags[h] + word), "r"(value) : "memory");
    }
}

// thread i waits until flags[first + i*stride] has reached `value` (epochs only grow; wrap-safe compare)
__global__ void slab_wait_kernel(const unsigned* flags, int first, int stride, int n, unsigned value, unsigned* err,
                                 long long timeout_cycles) {
    const int i = (int)threadIdx.x;
    if (i < n) {
        const unsigned* p = flags + first + (long long)i * stride;
        const long long t0 = clock64();
        for (;;) {
            unsigned v;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
            if ((int)(v - value) >= 0) break;
            if (clock64() - t0 > timeout_cycles) { atomicExch(err, 1u + (unsigned)i); break; }   // never hang the GPU
            __nanosleep(100);
        }
    }
    __syncthreads();
    __threadfence_system();
}

thread_local std::string g_slab_err;

int slab_fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_slab_err = buf;
    return code;
}

#define SLAB_CUDA(expr)                                                                                      \
    do {                                                                                                     \
        cudaError_t e__ = (expr);                                                                            \
        if (e__ != cudaSuccess) return slab_fail(B2FFT_E_CUDA, "%s: %s", #expr, cudaGetErrorString(e__));    \
    } while (0)
#define SLAB_TRY(expr)                                                \
    do {                                                              \
        int rc__ = (expr);                                            \
        if (rc__ != B2FFT_OK) return slab_fail(rc__, "%s: %s", #expr, b2fft_last_error()); \
    } while (0)

struct DevGuard {
    int prev = -1;
    bool changed = false;
    explicit DevGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) changed = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DevGuard() { if (changed) cudaSetDevice(prev); }
};

}  // namespace

// ---- the order of the X-pass launches (pure host logic, also exported as b2fft_slab_schedule_preview for the CPU tests)
// A cell = one launch of the X pass: (k, c) rows {z in chunk k} x {y in chunk c}; (k, -n) rows {z in chunk k} x {y in chunks
// 0..n-1}; (-1, c) rows {all local z} x {y in chunk c}.
static int slab_default_columns(int G, int C) {
    // the z-chunk-major part has to last about as long as the Y launch: measured at 2048^3 on B200 the Y launch takes 41.7/G ms
    // on 100 SMs and the exchange 114.5*(G-1)/G^2 ms at 600 GB/s, i.e. a fraction 0.364*G/(G-1) of the columns (8 GPUs: 3 of 8,
    // where 2 and 4 of 8 measured 0.6-1.2 ms slower; 4 GPUs: 4 of 8)
    int c1 = (int)((double)C * 0.364 * G / (G > 1 ? G - 1 : 1) + 0.5);
    if (c1 < 1) c1 = 1;
    if (c1 > C) c1 = C;
    return c1;
}
static void slab_schedule(int K, int C, int C1, bool overlap, std::vector<std::pair<int, int>>& cells) {
    cells.clear();
    if (overlap) {
        for (int k = 0; k < K; ++k) cells.emplace_back(k, -C1);
        for (int c = C1; c < C; ++c) cells.emplace_back(-1, c);
    } else {
        for (int k = 0; k < K; ++k)
            for (int c = 0; c < C; ++c) cells.emplace_back(k, c);
    }
}

struct b2fft_slab_plan {
    long long X, Y, Z, Zl, Xb, Yc, Zk;
    int G, rank, device, prec, C, K;
    size_t esz;
    b2fft_plan *fwd_y = nullptr, *fwd_x = nullptr, *fwd_z = nullptr;      // forward sub-plans
    // overlap mode (b2fft_slab_plan_set_overlap): ONE Y launch over the whole slab on all but `overlap_sms` SMs with a
    // progress counter per z-chunk; the X passes of chunk k wait for counter k instead of an event
    b2fft_plan* fwd_y_all = nullptr;
    b2fft_plan* fwd_x_all = nullptr;      // X pass of {all local z} x {one y-chunk}: the columns sent after the Y pass has finished
    int exchange_ctas_per_sm = 0;
    b2fft_plan* fwd_x_p1 = nullptr;       // X pass of {one z-chunk} x {the first p1_cols y-chunks}: sent while the Y launch runs
    int p1_cols = 0;
    int overlap_columns = 0;              // y-chunks sent z-chunk by z-chunk while the Y launch runs (0 = default, 3/8 of them)
    void* ws_y_all = nullptr;
    unsigned* d_progress = nullptr;
    int overlap_sms = 0;
    int z_max_ctas = 0;
    unsigned progress_target = 0;
    int normalize = 1, fast_math = 1;
    double scale = 1.0;
    b2fft_plan *inv_z = nullptr, *inv_x = nullptr, *inv_y = nullptr;      // inverse sub-plans
    void* ws[6] = {};                                                     // plan-owned workspaces of sub-plans that need one
    char* slab = nullptr;
    std::vector<char*> xslab;             // xslab of every rank, as addressable from this device
    std::vector<unsigned*> flags;         // flag words of every rank
    unsigned** d_flag_ptrs = nullptr;     // device copy of `flags`
    unsigned* d_err = nullptr;
    cudaStream_t sx = nullptr, sz = nullptr;
    std::vector<cudaEvent_t> ev_y;        // Y pass of z-chunk k done
    cudaEvent_t ev_start = nullptr, ev_sx = nullptr, ev_sz = nullptr;
    unsigned epoch = 0;                   // one per forward / inverse call
    unsigned last_inverse_epoch = 0;      // peers pulled from my x-slab during that call
    bool attached = false;
    // optional timeline of the last forward call (b2fft_slab_plan_set_trace): timing events at the phase boundaries
    bool trace = false;
    std::vector<std::pair<std::string, cudaEvent_t>> marks;
    void mark(const char* what, int a, int b, cudaStream_t st) {
        if (!trace) return;
        cudaEvent_t ev;
        if (cudaEventCreate(&ev) != cudaSuccess) return;
        cudaEventRecord(ev, st);
        char nm[64];
        snprintf(nm, sizeof nm, b >= 0 ? "%s[%d,%d]" : a >= 0 ? "%s[%d]" : "%s", what, a, b);
        marks.emplace_back(nm, ev);
    }
    void clear_marks() {
        for (auto& m : marks) cudaEventDestroy(m.second);
        marks.clear();
    }
    long long timeout_cycles = 20LL * 1000 * 1000 * 1000;   // ~10 s at 2 GHz
    // flag word layout (each word written by exactly one remote rank)
    int w_ready(int src) const { return src; }
    int w_chunk(int src, int c) const { return G + src * C + c; }
    int w_inv_chunk(int src, int c) const { return G + G * C + src * C + c; }
    int w_inv_done(int src) const { return G + 2 * G * C + src; }
    int n_words() const { return 2 * G + 2 * G * C; }
};

extern "C" {

const char* b2fft_slab_last_error(void) { return g_slab_err.c_str(); }

int b2fft_slab_plan_create(b2fft_slab_plan** out, const int64_t dims_xyz[3], int precision, int normalize, double scale,
                           int fast_math, int device, int rank, int nranks, int y_chunks, int z_chunks,
                           int exchange_ctas_per_sm) {
    if (!out || !dims_xyz) return slab_fail(B2FFT_E_INVALID, "null argument");
    *out = nullptr;
    const long long X = dims_xyz[0], Y = dims_xyz[1], Z = dims_xyz[2];
    const int G = nranks;
    if (G < 1 || G > 16 || (G & (G - 1))) return slab_fail(B2FFT_E_INVALID, "number of ranks must be a power of two <= 16");
    if (rank < 0 || rank >= G) return slab_fail(B2FFT_E_INVALID, "bad rank %d of %d", rank, G);
    if (X < 1 || Y < 1 || Z < 1 || (X & (X - 1)) || (Y & (Y - 1)) || (Z & (Z - 1)))
        return slab_fail(B2FFT_E_INVALID, "Array dimensions must be powers of two");
    if (Z % G || X % G) return slab_fail(B2FFT_E_INVALID, "Z and X must be divisible by the number of ranks");
    if (precision != B2FFT_F32 && precision != B2FFT_F64) return slab_fail(B2FFT_E_INVALID, "bad precision %d", precision);
    DevGuard guard(device);
    b2fft_slab_plan* sp = new b2fft_slab_plan();
    sp->X = X; sp->Y = Y; sp->Z = Z; sp->G = G; sp->rank = rank; sp->device = device; sp->prec = precision;
    sp->Zl = Z / G; sp->Xb = X / G;
    sp->esz = precision ? 16 : 8;
    sp->normalize = normalize; sp->fast_math = fast_math; sp->scale = scale;
    sp->exchange_ctas_per_sm = G > 1 ? exchange_ctas_per_sm : 0;
    // default pipeline: 8 y-chunks; 8 z-chunks with the Y pass hidden under the exchange when that is possible and pays
    // (>= 4 ranks: the exchange, not HBM, bounds the transform), else one z-chunk
    int want_overlap = -1;
    if (const char* e = getenv("B2FFT_SLAB_OVERLAP_SMS")) want_overlap = atoi(e);
    const bool try_overlap = G > 1 && (want_overlap > 0 || (want_overlap < 0 && G >= 4));
    long long C = y_chunks > 0 ? y_chunks : 8, K = z_chunks > 0 ? z_chunks : (try_overlap ? 8 : 1);
    if (C > Y) C = Y;
    while (Y % C) --C;
    if (K > sp->Zl) K = sp->Zl;
    while (sp->Zl % K) --K;
    sp->C = (int)C; sp->K = (int)K;
    sp->Yc = Y / C; sp->Zk = sp->Zl / K;
    const double nsize = (double)X * (double)Y * (double)Z;
    auto fail_free = [&](int rc) { b2fft_slab_plan_destroy(sp); return rc; };
    auto mk = [&](b2fft_plan** p, long long x, long long y, long long z, int axes, int apply_scale) {
        const int64_t d[3] = {x, y, z};
        return b2fft_plan_create_ex(p, d, axes, precision, B2FFT_INTERLEAVED, normalize, scale, fast_math, device, nsize, apply_scale);
    };
    int rc;
    // forward: Y pass per z-chunk, X pass per (z-chunk, y-chunk) with blocked stores, Z pass per y-chunk with the scale
    if ((rc = mk(&sp->fwd_y, X, Y, sp->Zk, B2FFT_AXIS_Y, 0)) != B2FFT_OK) return fail_free(slab_fail(rc, "%s", b2fft_last_error()));
    if ((rc = mk(&sp->fwd_x, X, sp->Yc, sp->Zk, B2FFT_AXIS_X, 0)) != B2FFT_OK) return fail_free(slab_fail(rc, "%s", b2fft_last_error()));
    // the Z pass runs beside the exchange kernels of the following y-chunks: $B2FFT_SLAB_Z_VARIANT names the kernel variant it
    // should prefer (tuning: a kernel that leaves room for the exchange CTAs on its SMs against the fastest one)
    const char* zv = getenv("B2FFT_SLAB_Z_VARIANT");
    if (zv && *zv) b2fft_set_preferred_variants(zv);
    rc = mk(&sp->fwd_z, sp->Xb, Z, sp->Yc, B2FFT_AXIS_Y, 1);
    if (zv && *zv) b2fft_set_preferred_variants(getenv("B2FFT_PREFER") ? getenv("B2FFT_PREFER") : "");
    if (rc != B2FFT_OK) return fail_free(slab_fail(rc, "%s", b2fft_last_error()));
    // inverse: Z pass per y-chunk, X pass per y-chunk pulling from the peers, Y pass over the whole slab with the scale
    if ((rc = mk(&sp->inv_z, sp->Xb, Z, sp->Yc, B2FFT_AXIS_Y, 0)) != B2FFT_OK) return fail_free(slab_fail(rc, "%s", b2fft_last_error()));
    if ((rc = mk(&sp->inv_x, X, sp->Yc, sp->Zl, B2FFT_AXIS_X, 0)) != B2FFT_OK) return fail_free(slab_fail(rc, "%s", b2fft_last_error()));
    if ((rc = mk(&sp->inv_y, X, Y, sp->Zl, B2FFT_AXIS_Y, 1)) != B2FFT_OK) return fail_free(slab_fail(rc, "%s", b2fft_last_error()));
    if (G > 1 && exchange_ctas_per_sm > 0 &&
        (rc = b2fft_plan_set_exchange_ctas(sp->fwd_x, exchange_ctas_per_sm)) != B2FFT_OK)
        return fail_free(slab_fail(rc, "%s", b2fft_last_error()));
    // sub-plans whose local axis is too long for one pass need scratch memory: owned here (in-place executes)
    b2fft_plan* subs[6] = {sp->fwd_y, sp->fwd_x, sp->fwd_z, sp->inv_z, sp->inv_x, sp->inv_y};
    for (int i = 0; i < 6; ++i) {
        size_t need = 0;
        if ((rc = b2fft_plan_workspace_bytes_ex(subs[i], 1, 1, &need)) != B2FFT_OK) return fail_free(slab_fail(rc, "%s", b2fft_last_error()));
        if (i == 1 || i == 4) {
            if (need) return fail_free(slab_fail(B2FFT_E_UNSUPPORTED, "X = %lld is too long for the single-pass exchange kernel", X));
            continue;
        }
        if (need) {
            cudaError_t e = cudaMalloc(&sp->ws[i], need);
            if (e != cudaSuccess) return fail_free(slab_fail(B2FFT_E_CUDA, "cudaMalloc of a %zu byte workspace: %s", need, cudaGetErrorString(e)));
            if ((rc = b2fft_plan_set_workspace(subs[i], sp->ws[i], need)) != B2FFT_OK) return fail_free(slab_fail(rc, "%s", b2fft_last_error()));
        }
    }
    cudaError_t e;
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if ((e = cudaStreamCreateWithFlags(&sp->sx, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithPriority(&sp->sz, cudaStreamNonBlocking, hi)) != cudaSuccess)
        return fail_free(slab_fail(B2FFT_E_CUDA, "stream creation: %s", cudaGetErrorString(e)));
    sp->ev_y.resize(sp->K);
    for (auto& ev : sp->ev_y)
        if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess)
            return fail_free(slab_fail(B2FFT_E_CUDA, "event creation: %s", cudaGetErrorString(e)));
    if ((e = cudaEventCreateWithFlags(&sp->ev_start, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&sp->ev_sx, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&sp->ev_sz, cudaEventDisableTiming)) != cudaSuccess)
        return fail_free(slab_fail(B2FFT_E_CUDA, "event creation: %s", cudaGetErrorString(e)));
    if ((e = cudaMalloc(&sp->d_err, sizeof(unsigned))) != cudaSuccess || (e = cudaMemset(sp->d_err, 0, sizeof(unsigned))) != cudaSuccess)
        return fail_free(slab_fail(B2FFT_E_CUDA, "cudaMalloc: %s", cudaGetErrorString(e)));
    if (try_overlap && sp->K > 1) {
        rc = b2fft_slab_plan_set_overlap(sp, want_overlap > 0 ? want_overlap : -1);
        if (rc != B2FFT_OK && rc != B2FFT_E_UNSUPPORTED) return fail_free(rc);
        if (rc == B2FFT_E_UNSUPPORTED && z_chunks <= 0) {
            // no hidden Y pass for these dimensions: z-chunks only add launches (profiles/r02_slab_pipeline.md)
            sp->K = 1; sp->Zk = sp->Zl;
            b2fft_plan_destroy(sp->fwd_y); b2fft_plan_destroy(sp->fwd_x);
            sp->fwd_y = sp->fwd_x = nullptr;
            if ((rc = mk(&sp->fwd_y, X, Y, sp->Zk, B2FFT_AXIS_Y, 0)) != B2FFT_OK) return fail_free(slab_fail(rc, "%s", b2fft_last_error()));
            if ((rc = mk(&sp->fwd_x, X, sp->Yc, sp->Zk, B2FFT_AXIS_X, 0)) != B2FFT_OK) return fail_free(slab_fail(rc, "%s", b2fft_last_error()));
            if (exchange_ctas_per_sm > 0 && (rc = b2fft_plan_set_exchange_ctas(sp->fwd_x, exchange_ctas_per_sm)) != B2FFT_OK)
                return fail_free(slab_fail(rc, "%s", b2fft_last_error()));
            size_t need = 0;
            if ((rc = b2fft_plan_workspace_bytes_ex(sp->fwd_y, 1, 1, &need)) != B2FFT_OK) return fail_free(slab_fail(rc, "%s", b2fft_last_error()));
            if (need) {
                if (sp->ws[0]) { cudaFree(sp->ws[0]); sp->ws[0] = nullptr; }
                if ((e = cudaMalloc(&sp->ws[0], need)) != cudaSuccess) return fail_free(slab_fail(B2FFT_E_CUDA, "cudaMalloc: %s", cudaGetErrorString(e)));
                if ((rc = b2fft_plan_set_workspace(sp->fwd_y, sp->ws[0], need)) != B2FFT_OK) return fail_free(slab_fail(rc, "%s", b2fft_last_error()));
            }
            for (size_t i = 1; i < sp->ev_y.size(); ++i) cudaEventDestroy(sp->ev_y[i]);
            sp->ev_y.resize(1);
        }
    }
    *out = sp;
    return B2FFT_OK;
}

int b2fft_slab_plan_set_overlap(b2fft_slab_plan* sp, int reserved_sms) {
    if (!sp) return slab_fail(B2FFT_E_INVALID, "null plan");
    DevGuard guard(sp->device);
    if (reserved_sms == 0 || sp->K < 2 || (reserved_sms < 0 && sp->G < 2)) {
        if (sp->fwd_y_all) { b2fft_plan_destroy(sp->fwd_y_all); sp->fwd_y_all = nullptr; }
        if (sp->fwd_x_all) { b2fft_plan_destroy(sp->fwd_x_all); sp->fwd_x_all = nullptr; }
        if (sp->fwd_z) b2fft_plan_set_max_ctas(sp->fwd_z, 0);
        sp->overlap_sms = 0;
        return reserved_sms == 0 ? B2FFT_OK : slab_fail(B2FFT_E_UNSUPPORTED, "hiding the Y pass needs >= 2 z-chunks");
    }
    int sms = 0;
    SLAB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, sp->device));
    // default: 48 of 148 SMs -- measured on 8 B200s at 2048^3 (profiles/r02_slab_pipeline.md): 40 SMs starve the exchange
    // (15.5-15.7 ms), 56-64 slow the Y and Z passes down (15.2-16.0 ms), 48: 14.9 ms
    int r = reserved_sms > 0 ? reserved_sms : (sms * 48) / 148;
    if (r < 1) r = 1;
    if (r > sms - 1) r = sms - 1;
    if (!sp->fwd_y_all) {
        const int64_t d[3] = {sp->X, sp->Y, sp->Zl};
        const double nsize = (double)sp->X * (double)sp->Y * (double)sp->Z;
        int rc = b2fft_plan_create_ex(&sp->fwd_y_all, d, B2FFT_AXIS_Y, sp->prec, B2FFT_INTERLEAVED, sp->normalize, sp->scale,
                                      sp->fast_math, sp->device, nsize, 0);
        if (rc != B2FFT_OK) return slab_fail(rc, "%s", b2fft_last_error());
    }
    if (!sp->d_progress) {
        SLAB_CUDA(cudaMalloc(&sp->d_progress, 64 * sizeof(unsigned)));
        SLAB_CUDA(cudaMemset(sp->d_progress, 0, 64 * sizeof(unsigned)));
    }
    int64_t target = 0;
    int rc = sp->K <= 64 ? b2fft_plan_set_progress(sp->fwd_y_all, sp->d_progress, sp->Zk, sms - r, &target) : B2FFT_E_UNSUPPORTED;
    if (rc != B2FFT_OK) {
        std::string why = rc == B2FFT_E_UNSUPPORTED && sp->K > 64 ? "more than 64 z-chunks" : b2fft_last_error();
        b2fft_plan_destroy(sp->fwd_y_all); sp->fwd_y_all = nullptr;
        sp->overlap_sms = 0;
        return slab_fail(rc, "%s", why.c_str());
    }
    sp->progress_target = (unsigned)target;                      // tiles per z-chunk
    // the columns that leave after the Y pass has finished go out as one launch over all local z
    if (!sp->fwd_x_all) {
        const int64_t d[3] = {sp->X, sp->Yc, sp->Zl};
        const double nsize = (double)sp->X * (double)sp->Y * (double)sp->Z;
        rc = b2fft_plan_create_ex(&sp->fwd_x_all, d, B2FFT_AXIS_X, sp->prec, B2FFT_INTERLEAVED, sp->normalize, sp->scale, sp->fast_math,
                                  sp->device, nsize, 0);
        if (rc == B2FFT_OK && sp->exchange_ctas_per_sm > 0) rc = b2fft_plan_set_exchange_ctas(sp->fwd_x_all, sp->exchange_ctas_per_sm);
        if (rc != B2FFT_OK) {
            std::string why = b2fft_last_error();
            if (sp->fwd_x_all) { b2fft_plan_destroy(sp->fwd_x_all); sp->fwd_x_all = nullptr; }
            b2fft_plan_destroy(sp->fwd_y_all); sp->fwd_y_all = nullptr;
            return slab_fail(rc, "%s", why.c_str());
        }
    }
    // the Z passes run beside the exchange of the following columns: they leave it the same SMs
    sp->z_max_ctas = sms - r;
    if ((rc = b2fft_plan_set_max_ctas(sp->fwd_z, sp->z_max_ctas)) != B2FFT_OK) return slab_fail(rc, "%s", b2fft_last_error());
    sp->overlap_sms = r;
    return B2FFT_OK;
}

int b2fft_slab_plan_set_option(b2fft_slab_plan* sp, const char* key, double value) {
    if (!sp || !key) return slab_fail(B2FFT_E_INVALID, "null argument");
    if (!strcmp(key, "overlap_columns")) {
        if (value < 0 || value > sp->C) return slab_fail(B2FFT_E_INVALID, "overlap_columns must be in [0, y_chunks]");
        sp->overlap_columns = (int)value;
        return B2FFT_OK;
    }
    return slab_fail(B2FFT_E_INVALID, "unknown slab option %s", key);
}

int b2fft_slab_schedule_preview(int nranks, int y_chunks, int z_chunks, int hidden_y, int overlap_columns, char* buf, size_t buflen) {
    if (!buf || !buflen || nranks < 1 || y_chunks < 1 || z_chunks < 1 || overlap_columns < 0)
        return slab_fail(B2FFT_E_INVALID, "bad argument");
    const int C1 = !hidden_y ? y_chunks : overlap_columns > 0 ? (overlap_columns < y_chunks ? overlap_columns : y_chunks)
                                                                : slab_default_columns(nranks, y_chunks);
    std::vector<std::pair<int, int>> cells;
    slab_schedule(z_chunks, y_chunks, C1, hidden_y != 0, cells);
    std::string out;
    for (const auto& c : cells) {
        char item[48];
        if (c.first < 0) snprintf(item, sizeof item, "all:%d;", c.second);
        else if (c.second < 0) snprintf(item, sizeof item, "%d:0-%d;", c.first, -c.second - 1);
        else snprintf(item, sizeof item, "%d:%d;", c.first, c.second);
        out += item;
    }
    if (out.size() + 1 > buflen) return slab_fail(B2FFT_E_INVALID, "buffer too small");
    snprintf(buf, buflen, "%s", out.c_str());
    return B2FFT_OK;
}

int b2fft_slab_plan_sizes(const b2fft_slab_plan* sp, size_t* slab_bytes, size_t* xslab_bytes, size_t* flag_bytes) {
    if (!sp) return slab_fail(B2FFT_E_INVALID, "null plan");
    if (slab_bytes) *slab_bytes = (size_t)(sp->Zl * sp->Y * sp->X) * sp->esz;
    if (xslab_bytes) *xslab_bytes = (size_t)(sp->Y * sp->Z * sp->Xb) * sp->esz;
    if (flag_bytes) *flag_bytes = (size_t)sp->n_words() * sizeof(unsigned);
    return B2FFT_OK;
}

int b2fft_slab_plan_geometry(const b2fft_slab_plan* sp, int64_t out[8]) {
    if (!sp || !out) return slab_fail(B2FFT_E_INVALID, "null argument");
    out[0] = sp->Zl; out[1] = sp->Xb; out[2] = sp->C; out[3] = sp->K; out[4] = sp->Yc; out[5] = sp->Zk; out[6] = sp->G; out[7] = sp->rank;
    return B2FFT_OK;
}

int b2fft_slab_plan_attach(b2fft_slab_plan* sp, void* slab, void* const* xslab_of_rank, void* const* flags_of_rank) {
    if (!sp || !slab || !xslab_of_rank || !flags_of_rank) return slab_fail(B2FFT_E_INVALID, "null argument");
    DevGuard guard(sp->device);
    sp->slab = (char*)slab;
    sp->xslab.assign(sp->G, nullptr);
    sp->flags.assign(sp->G, nullptr);
    for (int r = 0; r < sp->G; ++r) {
        if (!xslab_of_rank[r] || !flags_of_rank[r]) return slab_fail(B2FFT_E_INVALID, "null buffer of rank %d", r);
        if (((uintptr_t)xslab_of_rank[r] % 16) || ((uintptr_t)flags_of_rank[r] % 4)) return slab_fail(B2FFT_E_INVALID, "misaligned buffer of rank %d", r);
        sp->xslab[r] = (char*)xslab_of_rank[r];
        sp->flags[r] = (unsigned*)flags_of_rank[r];
    }
    if (((uintptr_t)slab % 16)) return slab_fail(B2FFT_E_INVALID, "misaligned slab");
    if (!sp->d_flag_ptrs) SLAB_CUDA(cudaMalloc(&sp->d_flag_ptrs, 16 * sizeof(unsigned*)));
    SLAB_CUDA(cudaMemcpy(sp->d_flag_ptrs, sp->flags.data(), sp->G * sizeof(unsigned*), cudaMemcpyHostToDevice));
    // my own flag words start at epoch 0 (the owner clears them; peers only ever write epochs >= 1)
    SLAB_CUDA(cudaMemset(sp->flags[sp->rank], 0, (size_t)sp->n_words() * sizeof(unsigned)));
    SLAB_CUDA(cudaDeviceSynchronize());
    sp->attached = true;
    return B2FFT_OK;
}

static int slab_signal(b2fft_slab_plan* sp, int word, cudaStream_t st) {
    slab_signal_kernel<<<1, 32, 0, st>>>(sp->d_flag_ptrs, sp->G, word, sp->epoch);
    SLAB_CUDA(cudaGetLastError());
    return B2FFT_OK;
}
static int slab_wait(b2fft_slab_plan* sp, int first, int stride, unsigned value, cudaStream_t st) {
    slab_wait_kernel<<<1, 32, 0, st>>>(sp->flags[sp->rank], first, stride, sp->G, value, sp->d_err, sp->timeout_cycles);
    SLAB_CUDA(cudaGetLastError());
    return B2FFT_OK;
}

int b2fft_slab_forward(b2fft_slab_plan* sp, void* cuda_stream) {
    if (!sp || !sp->attached) return slab_fail(B2FFT_E_INVALID, "slab plan has no buffers attached");
    DevGuard guard(sp->device);
    cudaStream_t s = (cudaStream_t)cuda_stream;
    const size_t esz = sp->esz;
    const int G = sp->G, C = sp->C, K = sp->K;
    ++sp->epoch;
    sp->clear_marks();
    sp->mark("start", -1, -1, s);
    if (G > 1) {
        // peers may still be pulling from my x-slab (previous inverse): wait for them, then tell everybody that my
        // x-slab may be overwritten (everything queued on `s` before this call has consumed it)
        if (sp->last_inverse_epoch) SLAB_TRY(slab_wait(sp, sp->w_inv_done(0), 1, sp->last_inverse_epoch, s));
        SLAB_TRY(slab_signal(sp, sp->w_ready(sp->rank), s));
    }
    SLAB_CUDA(cudaEventRecord(sp->ev_start, s));
    SLAB_CUDA(cudaStreamWaitEvent(sp->sz, sp->ev_start, 0));
    std::vector<void*> blk(G);
    const bool overlap = sp->fwd_y_all != nullptr && sp->overlap_sms > 0;
    if (overlap) {
        // one persistent Y launch over the whole slab on all but overlap_sms SMs; counter k completes with z-chunk k
        SLAB_CUDA(cudaMemsetAsync(sp->d_progress, 0, 64 * sizeof(unsigned), s));
        SLAB_CUDA(cudaEventRecord(sp->ev_y[0], s));
        SLAB_CUDA(cudaStreamWaitEvent(sp->sx, sp->ev_y[0], 0));
        SLAB_TRY(b2fft_execute(sp->fwd_y_all, sp->slab, nullptr, sp->slab, nullptr, 0, 1, s));
        sp->mark("Y_all", -1, -1, s);
    }
    // Order of the (z-chunk k, y-chunk c) cells of the X pass.  Z pass c needs column c complete, cell (k, c) needs the Y
    // pass of z-chunk k.  Event-ordered Y passes: k-major (the Y pass of chunk k+1 runs beside the cells of chunk k).
    // Hidden Y pass: the first C1 columns k-major while the Y launch works its way through the z-chunks (they take about as
    // long as the Y pass, so the cells rarely wait), then the other columns one after the other, each as one launch over
    // all z, so that the Z passes start early, spread out and hide under the stores of the following columns.
    // a cell = one launch: (k, c) rows {z in chunk k} x {y in chunk c}; (k, -n) rows {z in chunk k} x {y in chunks 0..n-1};
    // (-1, c) rows {all local z} x {y in chunk c}
    std::vector<std::pair<int, int>> cells;
    const int C1 = !overlap ? C : sp->overlap_columns > 0 ? (sp->overlap_columns < C ? sp->overlap_columns : C) : slab_default_columns(G, C);
    if (overlap && (!sp->fwd_x_p1 || sp->p1_cols != C1)) {
        if (sp->fwd_x_p1) { b2fft_plan_destroy(sp->fwd_x_p1); sp->fwd_x_p1 = nullptr; }
        const int64_t d[3] = {sp->X, (int64_t)C1 * sp->Yc, sp->Zk};
        const double nsize = (double)sp->X * (double)sp->Y * (double)sp->Z;
        SLAB_TRY(b2fft_plan_create_ex(&sp->fwd_x_p1, d, B2FFT_AXIS_X, sp->prec, B2FFT_INTERLEAVED, sp->normalize, sp->scale, sp->fast_math,
                                      sp->device, nsize, 0));
        if (sp->exchange_ctas_per_sm > 0) SLAB_TRY(b2fft_plan_set_exchange_ctas(sp->fwd_x_p1, sp->exchange_ctas_per_sm));
        sp->p1_cols = C1;
    }
    slab_schedule(K, C, C1, overlap, cells);
    std::vector<int> done_in_column(C, 0);
    int y_ready = -1;                                            // z-chunks whose Y pass the exchange stream has waited for
    for (const auto& cell : cells) {
        const bool whole = cell.first < 0, multi = cell.second < 0;
        const int k = whole ? 0 : cell.first, c = multi ? 0 : cell.second, ncols = multi ? -cell.second : 1;
        b2fft_plan* xp = whole ? sp->fwd_x_all : multi ? sp->fwd_x_p1 : sp->fwd_x;
        char* zk = sp->slab + (size_t)k * sp->Zk * sp->Y * sp->X * esz;
        while (y_ready < (whole ? K - 1 : k)) {
            ++y_ready;
            char* zy = sp->slab + (size_t)y_ready * sp->Zk * sp->Y * sp->X * esz;
            if (overlap) {
                slab_wait_kernel<<<1, 32, 0, sp->sx>>>(sp->d_progress, y_ready, 1, 1, sp->progress_target, sp->d_err, sp->timeout_cycles);
                SLAB_CUDA(cudaGetLastError());
                sp->mark("Y", y_ready, -1, sp->sx);
            } else {
                SLAB_TRY(b2fft_execute(sp->fwd_y, zy, nullptr, zy, nullptr, 0, 1, s));
                sp->mark("Y", y_ready, -1, s);
                SLAB_CUDA(cudaEventRecord(sp->ev_y[y_ready], s));
                SLAB_CUDA(cudaStreamWaitEvent(sp->sx, sp->ev_y[y_ready], 0));
            }
            if (y_ready == 0 && G > 1) { SLAB_TRY(slab_wait(sp, sp->w_ready(0), 1, sp->epoch, sp->sx)); sp->mark("peers_ready", -1, -1, sp->sx); }
        }
        // rows {z in chunk k} x {y in chunk c}: row (z, y) starts at slab[(z*Y + y)*X]; its x-block h goes to
        // xslab_h[(y*Z + rank*Zl + z)*Xb]
        char* src = zk + (size_t)c * sp->Yc * sp->X * esz;
        for (int h = 0; h < G; ++h)
            blk[h] = sp->xslab[h] + ((size_t)c * sp->Yc * sp->Z + (size_t)sp->rank * sp->Zl + (size_t)k * sp->Zk) * sp->Xb * esz;
        SLAB_TRY(b2fft_plan_set_output_blocks(xp, G, blk.data(), nullptr, 1, sp->Z * sp->Xb));
        SLAB_TRY(b2fft_plan_set_outer_split(xp, (int64_t)ncols * sp->Yc, sp->X, sp->Y * sp->X, sp->Z * sp->Xb, sp->Xb));
        SLAB_TRY(b2fft_execute(xp, src, nullptr, src, nullptr, 0, 1, sp->sx));
        sp->mark(whole ? "Xall" : multi ? "Xcols" : "X", whole ? c : k, whole || multi ? -1 : c, sp->sx);
        for (int cc = c; cc < c + ncols; ++cc) {
            if ((done_in_column[cc] += whole ? K : 1) != K) continue;
            const int c = cc;
            if (G > 1) {
                SLAB_TRY(slab_signal(sp, sp->w_chunk(sp->rank, c), sp->sx));
                SLAB_TRY(slab_wait(sp, sp->w_chunk(0, c), C, sp->epoch, sp->sz));
                sp->mark("chunk_arrived", c, -1, sp->sz);
            } else {
                SLAB_CUDA(cudaEventRecord(sp->ev_sx, sp->sx));
                SLAB_CUDA(cudaStreamWaitEvent(sp->sz, sp->ev_sx, 0));
            }
            char* zc = sp->xslab[sp->rank] + (size_t)c * sp->Yc * sp->Z * sp->Xb * esz;
            // the Z pass of the column that arrives last has nothing to share the GPU with: no grid cap
            int z_done = 0;
            for (int q = 0; q < C; ++q) z_done += done_in_column[q] == K ? 1 : 0;
            if (overlap && z_done == C) SLAB_TRY(b2fft_plan_set_max_ctas(sp->fwd_z, 0));
            SLAB_TRY(b2fft_execute(sp->fwd_z, zc, nullptr, zc, nullptr, 0, 1, sp->sz));
            if (overlap && z_done == C) SLAB_TRY(b2fft_plan_set_max_ctas(sp->fwd_z, sp->z_max_ctas));
            sp->mark("Z", c, -1, sp->sz);
        }
    }
    SLAB_CUDA(cudaEventRecord(sp->ev_sx, sp->sx));
    SLAB_CUDA(cudaEventRecord(sp->ev_sz, sp->sz));
    SLAB_CUDA(cudaStreamWaitEvent(s, sp->ev_sx, 0));
    SLAB_CUDA(cudaStreamWaitEvent(s, sp->ev_sz, 0));
    sp->mark("end", -1, -1, s);
    return B2FFT_OK;
}

int b2fft_slab_plan_set_trace(b2fft_slab_plan* sp, int on) {
    if (!sp) return slab_fail(B2FFT_E_INVALID, "null plan");
    sp->trace = on != 0;
    if (!sp->trace) sp->clear_marks();
    return B2FFT_OK;
}

/* "name:ms;name:ms;..." -- completion time of every phase of the last traced forward call, relative to its start.
 * Synchronises the device. */
int b2fft_slab_plan_trace(b2fft_slab_plan* sp, char* buf, size_t buflen) {
    if (!sp || !buf || !buflen) return slab_fail(B2FFT_E_INVALID, "bad argument");
    DevGuard guard(sp->device);
    SLAB_CUDA(cudaDeviceSynchronize());
    std::string out;
    for (size_t i = 0; i < sp->marks.size(); ++i) {
        float ms = 0;
        if (i > 0) SLAB_CUDA(cudaEventElapsedTime(&ms, sp->marks[0].second, sp->marks[i].second));
        char item[96];
        snprintf(item, sizeof item, "%s:%.4f;", sp->marks[i].first.c_str(), ms);
        out += item;
    }
    snprintf(buf, buflen, "%s", out.c_str());
    return B2FFT_OK;
}

int b2fft_slab_inverse(b2fft_slab_plan* sp, void* cuda_stream) {
    if (!sp || !sp->attached) return slab_fail(B2FFT_E_INVALID, "slab plan has no buffers attached");
    DevGuard guard(sp->device);
    cudaStream_t s = (cudaStream_t)cuda_stream;
    const size_t esz = sp->esz;
    const int G = sp->G, C = sp->C;
    ++sp->epoch;
    // the in-place Z pass below overwrites my x-slab: peers must have finished pulling from it (previous inverse)
    if (G > 1 && sp->last_inverse_epoch) SLAB_TRY(slab_wait(sp, sp->w_inv_done(0), 1, sp->last_inverse_epoch, s));
    SLAB_CUDA(cudaEventRecord(sp->ev_start, s));
    SLAB_CUDA(cudaStreamWaitEvent(sp->sx, sp->ev_start, 0));
    std::vector<const void*> src(G);
    for (int c = 0; c < C; ++c) {
        char* zc = sp->xslab[sp->rank] + (size_t)c * sp->Yc * sp->Z * sp->Xb * esz;
        SLAB_TRY(b2fft_execute(sp->inv_z, zc, nullptr, zc, nullptr, 1, 1, s));
        if (G > 1) {
            SLAB_TRY(slab_signal(sp, sp->w_inv_chunk(sp->rank, c), s));
            SLAB_TRY(slab_wait(sp, sp->w_inv_chunk(0, c), C, sp->epoch, sp->sx));
        } else {
            SLAB_CUDA(cudaEventRecord(sp->ev_y[0], s));
            SLAB_CUDA(cudaStreamWaitEvent(sp->sx, sp->ev_y[0], 0));
        }
        // rows {all local z} x {y in chunk c}: piece h of row (z, y) is xslab_h[(y*Z + rank*Zl + z)*Xb .. +Xb)
        for (int h = 0; h < G; ++h)
            src[h] = sp->xslab[h] + ((size_t)c * sp->Yc * sp->Z + (size_t)sp->rank * sp->Zl) * sp->Xb * esz;
        char* dst = sp->slab + (size_t)c * sp->Yc * sp->X * esz;
        SLAB_TRY(b2fft_plan_set_input_blocks(sp->inv_x, G, src.data()));
        SLAB_TRY(b2fft_plan_set_outer_split(sp->inv_x, sp->Yc, sp->Z * sp->Xb, sp->Xb, sp->X, sp->Y * sp->X));
        SLAB_TRY(b2fft_execute(sp->inv_x, dst, nullptr, dst, nullptr, 1, 1, sp->sx));
    }
    if (G > 1) SLAB_TRY(slab_signal(sp, sp->w_inv_done(sp->rank), sp->sx));
    sp->last_inverse_epoch = sp->epoch;
    SLAB_CUDA(cudaEventRecord(sp->ev_sx, sp->sx));
    SLAB_CUDA(cudaStreamWaitEvent(s, sp->ev_sx, 0));
    SLAB_TRY(b2fft_execute(sp->inv_y, sp->slab, nullptr, sp->slab, nullptr, 1, 1, s));
    return B2FFT_OK;
}

int b2fft_slab_plan_status(b2fft_slab_plan* sp, int* out) {
    if (!sp || !out) return slab_fail(B2FFT_E_INVALID, "null argument");
    DevGuard guard(sp->device);
    unsigned v = 0;
    SLAB_CUDA(cudaMemcpy(&v, sp->d_err, sizeof v, cudaMemcpyDeviceToHost));
    *out = (int)v;
    return B2FFT_OK;
}

int64_t b2fft_slab_plan_launch_count(const b2fft_slab_plan* sp) {
    if (!sp) return -1;
    int64_t n = 0;
    const b2fft_plan* subs[6] = {sp->fwd_y, sp->fwd_x, sp->fwd_z, sp->inv_z, sp->inv_x, sp->inv_y};
    for (auto* p : subs)
        if (p) n += b2fft_plan_launch_count(p);
    if (sp->fwd_y_all) n += b2fft_plan_launch_count(sp->fwd_y_all);
    if (sp->fwd_x_all) n += b2fft_plan_launch_count(sp->fwd_x_all);
    if (sp->fwd_x_p1) n += b2fft_plan_launch_count(sp->fwd_x_p1);
    return n;
}

int b2fft_slab_plan_describe(const b2fft_slab_plan* sp, char* buf, size_t buflen) {
    if (!sp || !buf || !buflen) return slab_fail(B2FFT_E_INVALID, "bad argument");
    char line[1024], a[256], b[256], c[256];
    b2fft_plan_describe(sp->fwd_y, a, sizeof a);
    b2fft_plan_describe(sp->fwd_x, b, sizeof b);
    b2fft_plan_describe(sp->fwd_z, c, sizeof c);
    for (char* s : {a, b, c})
        if (char* nl = strchr(s, '\n')) *nl = 0;
    snprintf(line, sizeof line,
             "rank %d/%d: z-slab [%lld][%lld][%lld] -> x-slab [%lld][%lld][%lld]; %d z-chunk(s) x %d y-chunk(s); "
             "Y pass {%s}%s | X pass + NVLink-blocked stores {%s} | peer flags | Z pass {%s}",
             sp->rank, sp->G, sp->Zl, sp->Y, sp->X, sp->Y, sp->Z, sp->Xb, sp->K, sp->C, a,
             sp->overlap_sms > 0 ? " as one launch with per-chunk progress counters, hidden under the exchange" : "", b, c);
    snprintf(buf, buflen, "%s", line);
    return B2FFT_OK;
}

int b2fft_slab_plan_destroy(b2fft_slab_plan* sp) {
    if (!sp) return B2FFT_OK;
    DevGuard guard(sp->device);
    b2fft_plan* subs[6] = {sp->fwd_y, sp->fwd_x, sp->fwd_z, sp->inv_z, sp->inv_x, sp->inv_y};
    for (auto* p : subs)
        if (p) b2fft_plan_destroy(p);
    if (sp->fwd_y_all) b2fft_plan_destroy(sp->fwd_y_all);
    if (sp->fwd_x_all) b2fft_plan_destroy(sp->fwd_x_all);
    if (sp->fwd_x_p1) b2fft_plan_destroy(sp->fwd_x_p1);
    if (sp->ws_y_all) cudaFree(sp->ws_y_all);
    if (sp->d_progress) cudaFree(sp->d_progress);
    for (void* w : sp->ws)
        if (w) cudaFree(w);
    if (sp->sx) cudaStreamDestroy(sp->sx);
    if (sp->sz) cudaStreamDestroy(sp->sz);
    for (auto ev : sp->ev_y)
        if (ev) cudaEventDestroy(ev);
    for (cudaEvent_t ev : {sp->ev_start, sp->ev_sx, sp->ev_sz})
        if (ev) cudaEventDestroy(ev);
    sp->clear_marks();
    if (sp->d_flag_ptrs) cudaFree(sp->d_flag_ptrs);
    if (sp->d_err) cudaFree(sp->d_err);
    delete sp;
    return B2FFT_OK;
}

}  // extern "C"
