// C ABI of libb2fft.so: plan objects, the host planner and the launch loop.
// See include/b2fft.h for the contract and the reference lines each entry point replaces.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <sstream>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/b2fft.h"
#include "kernels.cuh"
#include "twiddle.h"

namespace {

using b2::KernelVariant;

thread_local std::string g_err;

// L2-resident multi-pass execution: consecutive passes run chunk by chunk so that the data a
// pass wrote is still in the 126 MB L2 when the next pass reads it (0 disables).  Measured on
// B200 (profiles/r01_l2_chunking.md): with one kernel launch per chunk and pass the launch/drain
// overhead of the many short kernels outweighs the saved DRAM traffic, so the default is off.
// Destination-blocked stores of contiguous-axis passes (slab exchange) as TMA bulk copies: 0 = warp stores, 1 = bulk copies
// staged in the exchange buffer (the CTA waits for them), 2 = bulk copies from a staging buffer of their own (they drain
// while the CTA works on its next lines).  $B2FFT_BLK_BULK / b2fft_set_option("blk_bulk").
std::atomic<int> g_blk_bulk_mode{[] { const char* e = getenv("B2FFT_BLK_BULK"); return e ? atoi(e) : 1; }()};
std::atomic<long long> g_l2_chunk_bytes{[] {
    const char* e = getenv("B2FFT_L2_CHUNK_MB");
    return (long long)((e ? atof(e) : 0.0) * 1024.0 * 1024.0);
}()};

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t e__ = (expr);                                                                        \
        if (e__ != cudaSuccess) return fail(B2FFT_E_CUDA, "%s: %s", #expr, cudaGetErrorString(e__));     \
    } while (0)

// ------------------------------------------------------------------ registry
struct Registry {
    std::vector<KernelVariant> v;
    std::vector<bool> prepared;
    size_t n_default = 0;      // variants [0, n_default) are the production set, the rest tuning candidates
    std::vector<std::string> preferred;
    std::mutex mu;
    Registry() {
        b2::register_f32_row(v);
        b2::register_f32_col(v);
        b2::register_f64_row(v);
        b2::register_f64_col(v);
        b2::register_fused(v);
        n_default = v.size();
        b2::register_exp(v);   // tuning candidates, never picked by default (listed last)
        b2::register_exp2(v);
        b2::register_fused_exp(v);
        prepared.assign(v.size(), false);
        if (const char* e = getenv("B2FFT_PREFER")) set_preferred(e);
    }
    void set_preferred(const char* names) {
        preferred.clear();
        std::stringstream ss(names ? names : "");
        std::string tok;
        while (std::getline(ss, tok, ',')) {
            size_t a = tok.find_first_not_of(" \t"), b = tok.find_last_not_of(" \t");
            if (a != std::string::npos) preferred.push_back(tok.substr(a, b - a + 1));
        }
    }
    // contiguous axis -> W == 1; strided axis -> first variant (= measured preference order) whose W
    // divides `inner`.  Exception, measured on B200 (profiles/r01_strided_sweep.md): when consecutive
    // rows of the transformed axis are >= 256 KiB apart (Z passes of large 3-D arrays) every row
    // access pays a DRAM-page / TLB miss, throughput scales with the contiguous segment width, and the
    // widest tile wins regardless of staging (W=16: 59-74 %, W=8: 37-46 %, W=4: ~20 % of peak).
    // Split re/im rows prefer the TMA-staged kernels: the two planes arrive as two bulk copies instead of
    // twice the number of 4-byte global loads (measured, profiles/r01_sweep_split.txt).
    int pick(int prec, int log2n, bool contiguous, long long inner, bool split = false) {
        auto ok = [&](const KernelVariant& k) {
            if (k.prec != prec || k.log2n != log2n) return false;
            if (contiguous) return k.W == 1;
            if (k.kind == 3 && split) return false;          // fused two-step kernels: interleaved layout only
            return (k.W > 1 && k.kind != 1) ? (inner % k.W == 0) : false;
        };
        for (const auto& name : preferred)
            for (size_t i = 0; i < v.size(); ++i)
                if (name == v[i].name && ok(v[i])) return (int)i;
        // Rows of the transformed axis >= 256 KiB apart (Z passes of large 3-D arrays): the memory system serves ~45 G row
        // pieces per second whatever their size up to 128 bytes (profiles/r02_strided_copy_bw.txt), so only 128-byte pieces
        // get near the peak -- the fused two-step kernels (N >= 1024, interleaved).  Measured on 2048^3 / 1024^3 Z passes:
        // 3412 / 4121 GB/s against 2772 / 3707 GB/s for the widest single-pass tile (profiles/r02_fused2.md).  At smaller
        // pitch the single-pass TMA-staged tiles are faster (their neighbours share DRAM pages) and stay the default.
        static const int fused_mode = [] { const char* e = getenv("B2FFT_FUSED2"); return e ? atoi(e) : 1; }();   // 0 never, 1 large pitch, 2 always
        const long long pitch0 = inner * (prec ? 16 : 8);
        // complex64 N = 2048: the streamed fused kernel also beats the TMA-staged W = 4 tiles at small pitch (16 KiB: 4717 against
        // 4367 GB/s, profiles/r02_fused2p.md), so it is taken whenever 16 columns divide the inner stride
        if (!contiguous && fused_mode > 0 && (fused_mode > 1 || pitch0 >= (256 << 10) || (prec == 0 && log2n == 11)))
            for (size_t i = 0; i < n_default; ++i)
                if (v[i].kind == 3 && ok(v[i])) return (int)i;
        const long long pitch = inner * (prec ? 16 : 8);
        if (!contiguous && pitch >= (256 << 10)) {
            int best = -1;
            for (size_t i = 0; i < n_default; ++i)
                if (ok(v[i]) && v[i].kind == 0 && (best < 0 || v[i].W > v[best].W)) best = (int)i;
            if (best >= 0) return best;
        }
        if (contiguous && split)
            for (size_t i = 0; i < n_default; ++i)
                if (ok(v[i]) && v[i].kind == 1) return (int)i;
        for (size_t i = 0; i < v.size(); ++i)
            if (ok(v[i]) && v[i].kind != 3) return (int)i;
        if (!contiguous) return pick_direct_w1(prec, log2n);   // narrow inner dimension: one column per tile
        return -1;
    }
    // first plain strided-axis variant usable for `inner` (fallback of the tensor-map variants)
    int pick_direct_col(int prec, int log2n, long long inner, bool need_fs = false) {
        for (size_t i = 0; i < v.size(); ++i)
            if (v[i].prec == prec && v[i].log2n == log2n && v[i].W > 1 && v[i].kind == 0 && inner % v[i].W == 0 &&
                (!need_fs || v[i].fs))
                return (int)i;
        return need_fs ? -1 : pick_direct_w1(prec, log2n);
    }
    // four-step "A" pass: first FS-capable strided-axis variant whose W divides `inner`
    int pick_fs(int prec, int log2n, long long inner) {
        auto ok = [&](const KernelVariant& k) {
            return k.prec == prec && k.log2n == log2n && k.fs && k.W > 1 && inner % k.W == 0;
        };
        for (const auto& name : preferred)
            for (size_t i = 0; i < v.size(); ++i)
                if (name == v[i].name && ok(v[i])) return (int)i;
        const long long pitch = inner * (prec ? 16 : 8);
        if (pitch >= (256 << 10)) {   // see pick(): very long row pitch -> widest plain tile
            int best = -1;
            for (size_t i = 0; i < n_default; ++i)
                if (ok(v[i]) && v[i].kind == 0 && (best < 0 || v[i].W > v[best].W)) best = (int)i;
            if (best >= 0) return best;
        }
        for (size_t i = 0; i < v.size(); ++i)
            if (ok(v[i])) return (int)i;
        return -1;
    }
    // longest transform a single pass can do (contiguous axis: W == 1 variants, strided: W > 1)
    int max_log2(int prec, bool contiguous) {
        int m = 0;
        for (size_t i = 0; i < n_default; ++i)
            if (v[i].prec == prec && v[i].kind != 3 && (contiguous ? v[i].W == 1 : v[i].W > 1) && v[i].log2n > m) m = v[i].log2n;
        return m;
    }
    // first plain (non-TMA) W = 1 variant: strided fallback, and the fallback for unaligned pointers
    int pick_direct_w1(int prec, int log2n) {
        for (size_t i = 0; i < v.size(); ++i)
            if (v[i].prec == prec && v[i].log2n == log2n && v[i].W == 1 && v[i].kind == 0) return (int)i;
        return -1;
    }
    cudaError_t prepare(int idx) {
        std::lock_guard<std::mutex> lk(mu);
        // function attributes are per device context; cheap enough to redo per plan
        return v[idx].prepare();
    }
};
Registry& registry() {
    static Registry r;
    return r;
}

// ------------------------------------------------------------------ twiddle cache
struct TwiddleCache {
    std::mutex mu;
    std::map<std::tuple<int, int, int, int>, void*> tabs;   // (device, prec, NS, R) -> device pointer
    int get(int device, int prec, int NS, int R, const void** out) {
        std::lock_guard<std::mutex> lk(mu);
        auto key = std::make_tuple(device, prec, NS, R);
        auto it = tabs.find(key);
        if (it != tabs.end()) { *out = it->second; return 0; }
        void* d = nullptr;
        size_t bytes;
        if (prec == B2FFT_F32) {
            auto h = b2::make_stage_table<float>(NS, R);
            bytes = h.size() * sizeof(h[0]);
            CUDA_TRY(cudaMalloc(&d, bytes));
            CUDA_TRY(cudaMemcpy(d, h.data(), bytes, cudaMemcpyHostToDevice));
        } else {
            auto h = b2::make_stage_table<double>(NS, R);
            bytes = h.size() * sizeof(h[0]);
            CUDA_TRY(cudaMalloc(&d, bytes));
            CUDA_TRY(cudaMemcpy(d, h.data(), bytes, cudaMemcpyHostToDevice));
        }
        tabs[key] = d;
        *out = d;
        return 0;
    }
};
// four-step inter-pass twiddle tables (see PassParams::fs_t1 / fs_t2 in fft_core.cuh)
struct FsTwiddleCache {
    std::mutex mu;
    std::map<std::tuple<int, int, int, long long, long long, int>, void*> tabs;   // (device, prec, which, N, N2, rows)
    template <typename T>
    int make(long long N, long long N2, int rows, long long mult, void** out) {
        const auto h = b2::make_fs_table<T>(N, N2, rows, mult);
        void* d = nullptr;
        CUDA_TRY(cudaMalloc(&d, h.size() * sizeof(h[0])));
        CUDA_TRY(cudaMemcpy(d, h.data(), h.size() * sizeof(h[0]), cudaMemcpyHostToDevice));
        *out = d;
        return 0;
    }
    // which = 1: [TPC][N2] w_N^(t*n2);  which = 2: [E][N2] w_N^(TPC*c*n2)
    int get(int device, int prec, int which, long long N, long long N2, int TPC, int E, const void** out) {
        std::lock_guard<std::mutex> lk(mu);
        const int rows = which == 1 ? TPC : E;
        auto key = std::make_tuple(device, prec, which, N, N2, rows);
        auto it = tabs.find(key);
        if (it != tabs.end()) { *out = it->second; return 0; }
        void* d = nullptr;
        const long long mult = which == 1 ? 1 : TPC;
        int rc = prec == B2FFT_F32 ? make<float>(N, N2, rows, mult, &d) : make<double>(N, N2, rows, mult, &d);
        if (rc) return rc;
        tabs[key] = d;
        *out = d;
        return 0;
    }
};
FsTwiddleCache& fs_twiddles() {
    static FsTwiddleCache c;
    return c;
}

TwiddleCache& twiddles() {
    static TwiddleCache c;
    return c;
}

struct DeviceGuard {
    int prev = -1;
    bool changed = false;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int dev) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) {
            err = cudaSetDevice(dev);
            changed = (err == cudaSuccess);
        }
    }
    ~DeviceGuard() {
        if (changed) cudaSetDevice(prev);
    }
};

// plans without any pass (every transformed axis has length 1): out = in * f (mode 1) or in / f (mode 2)
template <typename T>
__global__ void scale_copy_kernel(const T* __restrict__ in, T* __restrict__ out, size_t n, T f, int mode) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = mode == 1 ? in[i] * f : in[i] / f;
}

bool is_pow2(long long v) { return v > 0 && (v & (v - 1)) == 0; }
int ilog2ll(long long v) {
    int l = 0;
    while ((1LL << l) < v) ++l;
    return l;
}

struct Pass {
    int variant;
    int fallback;          // plain variant used when a TMA variant's 16-byte alignment rule is not met
    int axis;              // 0 = x, 1 = y, 2 = z
    int log2n;
    long long n;
    long long inner;       // element stride of the transformed axis
    const void* tw[3];
    const void* tw_fb[3];  // twiddle tables of the fallback variant (its radices may differ)
    const void* tw_b[3];   // fused two-step variants: stage tables of step B (tw = step A, fs_t1/fs_t2 = inter-step twiddle)
    // four-step "A" pass (transposing): this pass is the length-n = N1 part of an axis of length
    // fs_total = N1*N2 whose elements are inner0 apart; it stores [n2][k1][inner0]
    bool transposing;
    long long fs_total, fs_n2, inner0;
    const void* fs_t1; const void* fs_t2;        // tables for `variant`
    const void* fs_t1_fb; const void* fs_t2_fb;  // ... and for `fallback`
};

}  // namespace

struct b2fft_plan {
    long long x, y, z;
    int axes_mask;
    int prec, layout, normalize, fast_math, device;
    double scale, norm_size;
    int apply_scale;
    std::vector<Pass> passes;
    void* workspace = nullptr;
    size_t workspace_bytes = 0;
    // destination-blocked output of the last pass (slab exchange), see b2fft_plan_set_output_blocks
    int nblocks = 0;
    void* blk0[B2_MAX_BLOCKS] = {};
    void* blk1[B2_MAX_BLOCKS] = {};
    long long blk_out_inner = 0, blk_out_outer_stride = 0;
    // source-blocked input of the first pass (b2fft_plan_set_input_blocks)
    int in_nblocks = 0;
    const void* in_blk[B2_MAX_BLOCKS] = {};
    // two-level outer index of the last pass (b2fft_plan_set_outer_split)
    int exchange_max_ctas = 0;   // grid cap of the last pass (b2fft_plan_set_exchange_ctas)
    // per-chunk progress counters of a single-pass plan (b2fft_plan_set_progress) and the grid cap that goes with them
    unsigned* progress = nullptr;
    long long progress_outer = 0;
    int progress_max_ctas = 0;
    int max_ctas = 0;            // grid cap of every pass whose kernel strides over its tiles (b2fft_plan_set_max_ctas)
    long long split_div = 0, split_in_lo = 0, split_in_hi = 0, split_out_lo = 0, split_out_hi = 0;
    std::atomic<long long> launches{0};
    long long fused_slots_for(long long slot_elems) const {
        return slot_elems > 0 ? (long long)(fused_scratch_bytes / ((size_t)slot_elems * (prec ? 16 : 8))) : 0;
    }
    int n_transposing = 0;   // four-step "A" passes (cannot run in place)
    // scratch of the fused two-step strided kernels: one super-tile slot per resident CTA, plan-owned, L2-resident in use
    void* fused_scratch = nullptr;
    size_t fused_scratch_bytes = 0;
    bool dry = false;        // b2fft_plan_preview: plan the passes only, no device work
};

namespace {

int stage_tables_r(b2fft_plan* pl, const int* radix, int S, long long n, const void** tw) {
    if (pl->dry) return 0;
    int NS = (int)n;
    for (int s = 0; s + 1 < S; ++s) {
        int rc = twiddles().get(pl->device, pl->prec, NS, radix[s], &tw[s]);
        if (rc) return rc;
        NS /= radix[s];
    }
    return 0;
}
int stage_tables(b2fft_plan* pl, const KernelVariant& kv, long long n, const void** tw) {
    return stage_tables_r(pl, kv.radix, kv.S, kv.kind == 3 ? (1LL << kv.log2n1) : n, tw);
}
// fused two-step variant: step B's stage tables and the inter-step twiddle tables w_N^(k1*n2) (same factorisation
// into two small tables as the four-step "A" passes, see FsTwiddleCache)
int fused_tables(b2fft_plan* pl, const KernelVariant& kv, Pass& p) {
    if (pl->dry) return 0;
    const long long n = 1LL << kv.log2n, n1 = 1LL << kv.log2n1, n2 = n >> kv.log2n1;
    int rc = stage_tables_r(pl, kv.radix_b, kv.S_b, n2, p.tw_b);
    if (rc) return rc;
    for (int which = 1; which <= 2; ++which) {
        rc = fs_twiddles().get(pl->device, pl->prec, which, n, n2, (int)(n1 / kv.E), kv.E, which == 1 ? &p.fs_t1 : &p.fs_t2);
        if (rc) return rc;
    }
    return 0;
}

// One kernel pass over [..][n][inner]; fs_total > 0 makes it the transposing first part of a
// four-step decomposition of an axis of length fs_total = n * N2 (inner = N2 * inner0).
int add_pass(b2fft_plan* pl, int axis, long long n, long long inner, bool contiguous, long long fs_total, long long inner0) {
    Registry& reg = registry();
    const int lg = ilog2ll(n);
    const bool fs = fs_total > 0;
    int vi = fs ? reg.pick_fs(pl->prec, lg, inner) : reg.pick(pl->prec, lg, contiguous, inner, pl->layout == B2FFT_SPLIT);
    if (vi < 0)
        return fail(B2FFT_E_UNSUPPORTED, "no kernel for axis %c of length %lld (%s) in this build", "xyz"[axis], n,
                    pl->prec ? "f64" : "f32");
    const KernelVariant& kv = reg.v[vi];
    Pass p{};
    p.variant = vi;
    p.fallback = kv.kind == 0 ? vi : (kv.kind == 1 || kv.kind == 4) ? reg.pick_direct_w1(pl->prec, lg) : reg.pick_direct_col(pl->prec, lg, inner, fs);
    if (kv.kind == 3 && fs) return fail(B2FFT_E_UNSUPPORTED, "fused variant picked for a transposing pass");
    if (p.fallback < 0) return fail(B2FFT_E_UNSUPPORTED, "no fallback kernel for axis %c", "xyz"[axis]);
    p.axis = axis;
    p.log2n = lg;
    p.n = n;
    p.inner = inner;
    int rc = stage_tables(pl, kv, n, p.tw);
    if (rc) return rc;
    cudaError_t e = pl->dry ? cudaSuccess : reg.prepare(vi);
    if (e != cudaSuccess) return fail(B2FFT_E_CUDA, "kernel attribute setup failed: %s", cudaGetErrorString(e));
    if (kv.kind == 3) {
        rc = fused_tables(pl, kv, p);
        if (rc) return rc;
        if (!pl->dry) {
            const long long slots = kv.grid_slots();
            if (slots <= 0) return fail(B2FFT_E_CUDA, "occupancy query failed for %s", kv.name);
            const size_t bytes = (size_t)slots * (size_t)kv.slot_elems * (pl->prec ? 16 : 8);
            if (bytes > pl->fused_scratch_bytes) pl->fused_scratch_bytes = bytes;
        }
    }
    const KernelVariant& fb = reg.v[p.fallback];
    if (p.fallback != vi) {
        rc = stage_tables(pl, fb, n, p.tw_fb);
        if (rc) return rc;
        e = pl->dry ? cudaSuccess : reg.prepare(p.fallback);
        if (e != cudaSuccess) return fail(B2FFT_E_CUDA, "kernel attribute setup failed: %s", cudaGetErrorString(e));
    } else {
        for (int s = 0; s < 3; ++s) p.tw_fb[s] = p.tw[s];
    }
    if (fs) {
        p.transposing = true;
        p.fs_total = fs_total;
        p.fs_n2 = fs_total / n;
        p.inner0 = inner0;
        for (int which = 1; which <= 2 && !pl->dry; ++which) {
            rc = fs_twiddles().get(pl->device, pl->prec, which, fs_total, p.fs_n2, (int)(n / kv.E), kv.E,
                                   which == 1 ? &p.fs_t1 : &p.fs_t2);
            if (rc) return rc;
            rc = fs_twiddles().get(pl->device, pl->prec, which, fs_total, p.fs_n2, (int)(n / fb.E), fb.E,
                                   which == 1 ? &p.fs_t1_fb : &p.fs_t2_fb);
            if (rc) return rc;
        }
    }
    pl->passes.push_back(p);
    return 0;
}

// All passes of one axis of length n whose elements are `inner` apart.  Lengths beyond what one
// CTA can hold are split four-step style, n = N1*N2: a transposing pass over n1 (stride N2*inner)
// with the inter-pass twiddle fused into its stores, then (recursively) the axis of length N2 with
// element stride N1*inner.  This replaces the reference's global-kernel chains
// (pyfft/plan.py:141-143, pyfft/kernel.py:259-283) with 2 (n <= 2^22) or 3 DRAM round trips.
int add_axis(b2fft_plan* pl, int axis, long long n, long long inner, bool contiguous) {
    Registry& reg = registry();
    const int lg = ilog2ll(n);
    if (lg <= reg.max_log2(pl->prec, contiguous)) return add_pass(pl, axis, n, inner, contiguous, 0, 0);
    const int col = reg.max_log2(pl->prec, false);
    if (col < 1) return fail(B2FFT_E_UNSUPPORTED, "no strided-axis kernels in this build");
    if (!is_pow2(inner)) return fail(B2FFT_E_UNSUPPORTED, "four-step split of an axis behind a non-power-of-two untransformed axis");
    const int npass = (lg + col - 1) / col;
    const int l1 = lg / npass;
    const long long n1 = 1LL << l1, n2 = n >> l1;
    int rc = add_pass(pl, axis, n1, n2 * inner, false, n, inner);
    if (rc) return rc;
    return add_axis(pl, axis, n2, n1 * inner, false);
}

int build_passes(b2fft_plan* pl) {
    const long long dims[3] = {pl->x, pl->y, pl->z};
    const long long inner[3] = {1, pl->x, pl->x * pl->y};
    for (int a = 0; a < 3; ++a) {
        if (!(pl->axes_mask & (1 << a)) || dims[a] <= 1) continue;
        int rc = add_axis(pl, a, dims[a], inner[a], a == 0);
        if (rc) return rc;
    }
    pl->n_transposing = 0;
    for (const Pass& p : pl->passes) pl->n_transposing += p.transposing ? 1 : 0;
    return 0;
}

template <typename T>
int launch_pass(b2fft_plan* pl, const Pass& ps, const void* in0, const void* in1, void* out0, void* out1, int inverse,
                long long outer_count, bool last, cudaStream_t stream) {
    const bool split = pl->layout == B2FFT_SPLIT;
    int vi = ps.variant;
    if (registry().v[vi].kind != 0) {   // TMA needs 16-byte aligned sources (and 16-byte multiples as row pitch)
        const size_t pitch = (size_t)ps.inner * (split ? sizeof(T) : 2 * sizeof(T));
        if (((uintptr_t)in0 % 16) != 0 || (split && ((uintptr_t)in1 % 16) != 0) ||
            (registry().v[vi].kind == 2 && (pitch % 16 != 0 || outer_count > 0x7fffffffLL)) ||
            (registry().v[vi].kind == 4 && (((uintptr_t)out0 % 16) != 0 || (split && ((uintptr_t)out1 % 16) != 0))))   // 16-byte stores too
            vi = ps.fallback;
    }
    const bool in_blocked = pl->in_nblocks > 0 && &ps == &pl->passes.front();
    if (in_blocked) vi = ps.fallback;                                    // source-blocked loads live in the plain kernels
    const bool outer_split = last && pl->split_div > 0;
    if (outer_split && registry().v[vi].kind != 0) vi = ps.fallback;   // TMA staging assumes a dense outer index
    if (last && pl->nblocks > 0 && !registry().v[vi].blk) vi = ps.fallback;
    if (registry().v[vi].kind == 3 && (!pl->fused_scratch || (last && pl->exchange_max_ctas > 0))) vi = ps.fallback;
    const KernelVariant& kv = registry().v[vi];
    b2::PassParams<T> p{};
    if (split && inverse) {   // IDFT(z) = swap(DFT(swap(z))): swap the planes instead of the registers
        p.in0 = (const T*)in1; p.in1 = (const T*)in0; p.out0 = (T*)out1; p.out1 = (T*)out0;
    } else {
        p.in0 = (const T*)in0; p.in1 = (const T*)in1; p.out0 = (T*)out0; p.out1 = (T*)out1;
    }
    for (int s = 0; s < 3; ++s) p.tw[s] = (const T*)(vi == ps.variant ? ps.tw[s] : ps.tw_fb[s]);
    p.inner = ps.inner;
    p.inner_blocks = ps.inner / kv.W;
    p.outer_stride = ps.n * ps.inner;
    p.n_tiles = outer_count * p.inner_blocks;
    p.out_inner = p.inner;
    p.out_outer_stride = p.outer_stride;
    p.out_blk_log2 = -1;
    p.in_blk_log2 = -1;
    if (in_blocked) {
        if (kv.W != 1 || kv.kind != 0 || split) return fail(B2FFT_E_UNSUPPORTED, "source-blocked loads need a plain contiguous-axis kernel");
        p.in_blk_log2 = ilog2ll(ps.n / pl->in_nblocks);
        for (int h = 0; h < pl->in_nblocks; ++h) p.in_blk[h] = (const T*)pl->in_blk[h];
    }
    if (kv.kind == 3) {
        for (int s = 0; s < 3; ++s) p.tw_b[s] = (const T*)ps.tw_b[s];
        p.fs_t1 = (const T*)ps.fs_t1;
        p.fs_t2 = (const T*)ps.fs_t2;
        p.scratch = (T*)pl->fused_scratch;
        p.scratch_slots = pl->fused_slots_for(kv.slot_elems);
    }
    if (ps.transposing) {
        p.out_inner = ps.inner0;
        p.fs_log2_inner = ilog2ll(ps.inner0);
        p.fs_n2 = ps.fs_n2;
        p.fs_col_stride = ps.n * ps.inner0;
        p.fs_t1 = (const T*)(vi == ps.variant ? ps.fs_t1 : ps.fs_t1_fb);
        p.fs_t2 = (const T*)(vi == ps.variant ? ps.fs_t2 : ps.fs_t2_fb);
        // contiguous axis: whole transposed output rows leave as bulk copies (16-byte alignment rules)
        static const bool fsb_ok = [] { const char* e = getenv("B2FFT_FS_BULK"); return !e || atoi(e) != 0; }();
        // measured (profiles/r01_tma_store_fs_bulk.md): +33..42 % for N = 2^16..2^18, +18 % at 2^24, a few % at
        // 2^20; the 2048-row x W=4 TMA tiles of 2^22 lose 7 % and keep their direct stores
        p.fs_bulk = fsb_ok && !split && ps.inner0 == 1 && kv.S > 1 && ((uintptr_t)out0 % 16) == 0 &&
                    (ps.n * (long long)(2 * sizeof(T))) % 16 == 0 && !(kv.kind == 2 && kv.log2n >= 11);
    }
    if (last && pl->nblocks > 0) {
        if (!kv.blk) return fail(B2FFT_E_UNSUPPORTED, "kernel %s has no destination-blocked store", kv.name);
        if (ps.n % pl->nblocks) return fail(B2FFT_E_INVALID, "last axis length %lld not divisible into %d blocks", ps.n, pl->nblocks);
        p.out_blk_log2 = ilog2ll(ps.n / pl->nblocks);
        p.out_inner = pl->blk_out_inner;
        p.out_outer_stride = pl->blk_out_outer_stride;
        const bool sw = split && inverse;
        bool aligned = true;
        for (int h = 0; h < pl->nblocks; ++h) {
            p.out_blk0[h] = (T*)(sw ? pl->blk1[h] : pl->blk0[h]);
            p.out_blk1[h] = (T*)(sw ? pl->blk0[h] : pl->blk1[h]);
            aligned = aligned && ((uintptr_t)p.out_blk0[h] % 16) == 0;
        }
        // contiguous-axis exchange pass: every (line, block) piece is contiguous -> TMA bulk stores
        static const bool bulk_ok = [] { const char* e = getenv("B2FFT_BLK_BULK"); return !e || atoi(e) != 0; }();
        const long long piece = (ps.n / pl->nblocks) * (long long)(2 * sizeof(T));
        const long long lo = pl->split_div > 0 ? pl->split_out_lo : pl->blk_out_outer_stride;
        const long long hi = pl->split_div > 0 ? pl->split_out_hi : 0;
        p.blk_bulk = (bulk_ok && !split && kv.W == 1 && kv.S > 1 && ps.inner == 1 && pl->blk_out_inner == 1 && aligned &&
                     piece % 16 == 0 && (lo * (long long)(2 * sizeof(T))) % 16 == 0 && (hi * (long long)(2 * sizeof(T))) % 16 == 0)
                        ? g_blk_bulk_mode.load() : 0;
    }
    if (pl->max_ctas > 0 && (kv.kind == 0 || kv.kind == 3)) p.max_ctas = pl->max_ctas;
    if (pl->progress) {
        if (!kv.progress) return fail(B2FFT_E_UNSUPPORTED, "kernel %s publishes no progress counters", kv.name);
        p.progress = pl->progress;
        p.progress_tiles = pl->progress_outer * p.inner_blocks;
        p.max_ctas = pl->progress_max_ctas;
    }
    if (last && pl->exchange_max_ctas > 0) {
        p.max_ctas = pl->exchange_max_ctas;
        if (registry().v[vi].kind != 0) return fail(B2FFT_E_UNSUPPORTED, "grid cap needs a plain kernel variant");
    }
    if (outer_split) {
        p.outer_div = pl->split_div;
        p.outer_stride = pl->split_in_lo;
        p.in_stride_hi = pl->split_in_hi;
        p.out_outer_stride = pl->split_out_lo;
        p.out_stride_hi = pl->split_out_hi;
    }
    p.scale = (T)1;
    p.scale_mode = 0;
    if (last && pl->apply_scale) {
        // pyfft/kernel.py:23-37: forward divides by 1/scale, inverse by scale * (size if normalize)
        double coeff = inverse ? (pl->normalize ? pl->norm_size : 1.0) * pl->scale : 1.0 / pl->scale;
        if (coeff != 1.0) {
            if (pl->fast_math) { p.scale = (T)(1.0 / coeff); p.scale_mode = 1; }
            else { p.scale = (T)coeff; p.scale_mode = 2; }
        }
    }
    cudaError_t e = kv.launch(split ? 1 : 0, inverse ? 1 : 0, &p, stream);
    if (e != cudaSuccess) return fail(B2FFT_E_CUDA, "launch of %s failed: %s", kv.name, cudaGetErrorString(e));
    pl->launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

}  // namespace

extern "C" {

int b2fft_version(void) { return 100; }

const char* b2fft_last_error(void) { return g_err.c_str(); }

int b2fft_plan_create_ex(b2fft_plan** out, const int64_t dims_xyz[3], int axes_mask, int precision, int layout,
                         int normalize, double scale, int fast_math, int device, double norm_size, int apply_scale) {
    if (!out || !dims_xyz) return fail(B2FFT_E_INVALID, "null argument");
    *out = nullptr;
    if (axes_mask < 0 || axes_mask > 7) return fail(B2FFT_E_INVALID, "bad axes mask %d", axes_mask);
    // transformed axes: powers of two (pyfft/plan.py:23-24); an axis the mask leaves out only counts lines and may have any
    // length (slab pipeline: the X pass over a z-chunk x three of eight y-chunks)
    for (int a = 0; a < 3; ++a)
        if (((axes_mask >> a) & 1) ? !is_pow2(dims_xyz[a]) : dims_xyz[a] < 1)
            return fail(B2FFT_E_INVALID, "Array dimensions must be powers of two");
    if (precision != B2FFT_F32 && precision != B2FFT_F64) return fail(B2FFT_E_INVALID, "bad precision %d", precision);
    if (layout != B2FFT_INTERLEAVED && layout != B2FFT_SPLIT) return fail(B2FFT_E_INVALID, "bad layout %d", layout);
    if (!(scale == scale) || scale == 0.0) return fail(B2FFT_E_INVALID, "scale must be a non-zero number");
    if (axes_mask < 0 || axes_mask > 7) return fail(B2FFT_E_INVALID, "bad axes mask %d", axes_mask);
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(B2FFT_E_INVALID, "bad device %d (have %d)", device, ndev);
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return fail(B2FFT_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(guard.err));

    b2fft_plan* pl = new b2fft_plan();
    pl->x = dims_xyz[0]; pl->y = dims_xyz[1]; pl->z = dims_xyz[2];
    pl->axes_mask = axes_mask;
    pl->prec = precision; pl->layout = layout; pl->normalize = normalize ? 1 : 0; pl->fast_math = fast_math ? 1 : 0;
    pl->device = device;
    pl->scale = scale;
    pl->norm_size = norm_size > 0 ? norm_size : (double)pl->x * (double)pl->y * (double)pl->z;
    pl->apply_scale = apply_scale ? 1 : 0;
    int rc = build_passes(pl);
    if (rc) { delete pl; return rc; }
    if (pl->fused_scratch_bytes) {
        cudaError_t e = cudaMalloc(&pl->fused_scratch, pl->fused_scratch_bytes);
        if (e != cudaSuccess) {
            delete pl;
            return fail(B2FFT_E_CUDA, "cudaMalloc of %zu scratch bytes: %s", pl->fused_scratch_bytes, cudaGetErrorString(e));
        }
    }
    // the table uploads above are cudaMemcpy calls from pageable memory, ordered on the legacy default stream only:
    // drain it so that an execute on a non-blocking stream right after plan creation cannot overtake them
    if (cudaError_t e = cudaStreamSynchronize(cudaStreamLegacy); e != cudaSuccess) {
        b2fft_plan_destroy(pl);
        return fail(B2FFT_E_CUDA, "cudaStreamSynchronize: %s", cudaGetErrorString(e));
    }
    *out = pl;
    return B2FFT_OK;
}

int b2fft_plan_create(b2fft_plan** out, int rank, const int64_t dims_xyz[3], int precision, int layout, int normalize,
                      double scale, int fast_math, int device) {
    if (rank < 1 || rank > 3) return fail(B2FFT_E_INVALID, "Wrong shape");
    return b2fft_plan_create_ex(out, dims_xyz, B2FFT_AXIS_X | B2FFT_AXIS_Y | B2FFT_AXIS_Z, precision, layout, normalize,
                                scale, fast_math, device, 0.0, 1);
}

// Single-pass-per-axis plans work tile-in-place and need no scratch memory.  Plans with four-step
// (transposing) passes need one buffer of the size of the data when the transposing pass cannot
// write straight into `out`: always for in-place executes, and for out-of-place executes when a
// transposing pass is not the first pass.
static size_t workspace_need(const b2fft_plan* plan, int64_t batch, bool in_place) {
    if (plan->n_transposing == 0) return 0;
    int later = 0;
    for (size_t i = 1; i < plan->passes.size(); ++i) later += plan->passes[i].transposing ? 1 : 0;
    if (!in_place && later == 0) return 0;
    const size_t csize = plan->prec ? 16 : 8;
    return (size_t)(plan->x * plan->y * plan->z) * (size_t)batch * csize;
}

int b2fft_plan_workspace_bytes(const b2fft_plan* plan, int64_t batch, size_t* out) {
    if (!plan || !out || batch < 0) return fail(B2FFT_E_INVALID, "bad argument");
    *out = workspace_need(plan, batch, true);
    return B2FFT_OK;
}

int b2fft_plan_workspace_bytes_ex(const b2fft_plan* plan, int64_t batch, int in_place, size_t* out) {
    if (!plan || !out || batch < 0) return fail(B2FFT_E_INVALID, "bad argument");
    *out = workspace_need(plan, batch, in_place != 0);
    return B2FFT_OK;
}

int b2fft_plan_set_workspace(b2fft_plan* plan, void* dptr, size_t bytes) {
    if (!plan) return fail(B2FFT_E_INVALID, "null plan");
    plan->workspace = dptr;
    plan->workspace_bytes = bytes;
    return B2FFT_OK;
}

int b2fft_execute(b2fft_plan* plan, const void* in0, const void* in1, void* out0, void* out1, int inverse,
                  int64_t batch, void* cuda_stream) {
    if (!plan) return fail(B2FFT_E_INVALID, "null plan");
    if (batch < 0) return fail(B2FFT_E_INVALID, "negative batch");
    const bool split = plan->layout == B2FFT_SPLIT;
    if (!in0 || (split && !in1)) return fail(B2FFT_E_INVALID, "null data pointer");
    if (!out0 || (split && !out1)) return fail(B2FFT_E_INVALID, "null data pointer");
    const size_t align = split ? (plan->prec ? 8 : 4) : (plan->prec ? 16 : 8);
    const void* ptrs[4] = {in0, in1, out0, out1};
    for (int i = 0; i < 4; ++i)
        if (ptrs[i] && ((uintptr_t)ptrs[i] % align) != 0)
            return fail(B2FFT_E_INVALID, "device pointer %d is not %zu-byte aligned", i, align);
    if (batch == 0) return B2FFT_OK;
    DeviceGuard guard(plan->device);
    if (guard.err != cudaSuccess) return fail(B2FFT_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(guard.err));
    cudaStream_t stream = (cudaStream_t)cuda_stream;

    if (plan->passes.empty()) {   // all axes of length 1 (or masked out): the identity times the scale factor
        const size_t reals = (size_t)(plan->x * plan->y * plan->z) * (size_t)batch * (split ? 1 : 2);
        double coeff = 1.0;   // pyfft/kernel.py:23-37, as in launch_pass()
        if (plan->apply_scale) coeff = inverse ? (plan->normalize ? plan->norm_size : 1.0) * plan->scale : 1.0 / plan->scale;
        const int mode = coeff == 1.0 ? 0 : plan->fast_math ? 1 : 2;
        const double f = mode == 1 ? 1.0 / coeff : coeff;
        for (int pl = 0; pl < (split ? 2 : 1); ++pl) {
            const void* src = pl ? in1 : in0;
            void* dst = pl ? out1 : out0;
            if (mode == 0) {
                if (src != dst) CUDA_TRY(cudaMemcpyAsync(dst, src, reals * (plan->prec ? 8 : 4), cudaMemcpyDeviceToDevice, stream));
                continue;
            }
            const unsigned blocks = (unsigned)((reals + 255) / 256 > 65535 * 16 ? 65535 * 16 : (reals + 255) / 256);
            if (plan->prec) scale_copy_kernel<double><<<blocks, 256, 0, stream>>>((const double*)src, (double*)dst, reals, f, mode);
            else scale_copy_kernel<float><<<blocks, 256, 0, stream>>>((const float*)src, (float*)dst, reals, (float)f, mode);
            CUDA_TRY(cudaGetLastError());
            plan->launches.fetch_add(1, std::memory_order_relaxed);
        }
        return B2FFT_OK;
    }
    const size_t np_all = plan->passes.size();
    if (plan->n_transposing > 0) {
        // ---- four-step plans: a transposing pass must write to a different buffer than it reads.
        // Walk backwards from "result in out": a transposing pass reads the other buffer of
        // {out, workspace}, any other pass runs in place.  The first pass of an out-of-place execute
        // reads `in` whatever it writes; an in-place execute that would have to start in the
        // workspace lets its last (never transposing) pass move the data instead.
        const bool in_place = (in0 == out0) || (split && in1 == out1);
        const size_t need = workspace_need(plan, batch, in_place);
        if (need > plan->workspace_bytes || (need && !plan->workspace))
            return fail(B2FFT_E_INVALID, "this plan needs a workspace of %zu bytes for batch %lld (have %zu); see "
                        "b2fft_plan_workspace_bytes_ex / b2fft_plan_set_workspace", need, (long long)batch, plan->workspace_bytes);
        if (need && ((uintptr_t)plan->workspace % 16) != 0) return fail(B2FFT_E_INVALID, "workspace must be 16-byte aligned");
        enum { OUT = 1, WS = 2 };
        int src[16], dst[16];
        if (np_all > 16) return fail(B2FFT_E_UNSUPPORTED, "too many passes");
        int loc = OUT;
        for (size_t k = np_all; k-- > 0;) {
            dst[k] = loc;
            const bool move = plan->passes[k].transposing ||
                              (in_place && k + 1 == np_all && (plan->n_transposing & 1));
            src[k] = move ? (loc == OUT ? WS : OUT) : loc;
            loc = src[k];
        }
        const size_t plane = (size_t)(plan->x * plan->y * plan->z) * (size_t)batch * (plan->prec ? 8 : 4);
        void* ws0 = plan->workspace;
        void* ws1 = split ? (void*)((char*)plan->workspace + plane) : nullptr;
        const long long vol = plan->x * plan->y * plan->z;
        for (size_t k = 0; k < np_all; ++k) {
            const Pass& ps = plan->passes[k];
            const void* ci0 = (k == 0) ? in0 : (src[k] == OUT ? out0 : ws0);
            const void* ci1 = (k == 0) ? in1 : (src[k] == OUT ? out1 : ws1);
            void* co0 = dst[k] == OUT ? out0 : ws0;
            void* co1 = dst[k] == OUT ? out1 : ws1;
            const long long blocks = batch * (vol / (ps.n * ps.inner));
            const bool last = k + 1 == np_all;
            int rc = plan->prec == B2FFT_F32
                         ? launch_pass<float>(plan, ps, ci0, ci1, co0, co1, inverse, blocks, last, stream)
                         : launch_pass<double>(plan, ps, ci0, ci1, co0, co1, inverse, blocks, last, stream);
            if (rc) return rc;
        }
        return B2FFT_OK;
    }
    // ---- schedule: which consecutive passes share an L2-resident chunk
    // unit = the smallest block of data that the passes of a group never leave: a whole transform
    // ("volume") when it fits the L2 budget, else an XY plane for the X+Y passes (Z then runs alone).
    const size_t csize = (size_t)(plan->prec ? 16 : 8);
    const long long plane_elems = plan->x * plan->y, vol_elems = plane_elems * plan->z;
    const long long budget = plan->nblocks > 0 ? 0 : g_l2_chunk_bytes.load();
    const size_t np = plan->passes.size();
    struct Group { size_t first, last; long long unit_elems, units, chunk_units; };
    Group groups[3];
    int ng = 0;
    const bool out_of_place = (in0 != out0);
    // out-of-place streams the input through L2 as well: halve the resident chunk
    const long long eff_budget = out_of_place ? budget / 2 : budget;
    if (np >= 2 && eff_budget > 0 && (long long)(vol_elems * csize) <= eff_budget) {
        const long long cu = eff_budget / (long long)(vol_elems * csize);
        groups[ng++] = {0, np - 1, vol_elems, batch, cu < 1 ? 1 : cu};
    } else if (np >= 2 && eff_budget > 0 && plan->passes[0].axis == 0 && plan->passes[1].axis == 1 &&
               (long long)(plane_elems * csize) <= eff_budget) {
        const long long cu = eff_budget / (long long)(plane_elems * csize);
        groups[ng++] = {0, 1, plane_elems, batch * plan->z, cu < 1 ? 1 : cu};
        if (np == 3) groups[ng++] = {2, 2, vol_elems, batch, batch};
    } else {
        for (size_t i = 0; i < np; ++i) groups[ng++] = {i, i, vol_elems, batch, batch};
    }
    // a chunk that covers (nearly) everything gains nothing: run whole passes
    for (int g = 0; g < ng; ++g)
        if (groups[g].chunk_units * 2 > groups[g].units) groups[g].chunk_units = groups[g].units;

    const size_t in_esz = split ? csize / 2 : csize;   // bytes per element of each plane pointer
    auto off = [&](const void* ptr, long long elems) -> const void* {
        return ptr ? (const void*)((const char*)ptr + (size_t)elems * in_esz) : nullptr;
    };
    bool first_group = true;
    for (int g = 0; g < ng; ++g) {
        const Group& gr = groups[g];
        for (long long u0 = 0; u0 < gr.units; u0 += gr.chunk_units) {
            const long long nu = (gr.units - u0 < gr.chunk_units) ? gr.units - u0 : gr.chunk_units;
            const long long eoff = u0 * gr.unit_elems;
            for (size_t i = gr.first; i <= gr.last; ++i) {
                const Pass& ps = plan->passes[i];
                const bool from_input = first_group && i == gr.first;
                const void* ci0 = off(from_input ? in0 : out0, eoff);
                const void* ci1 = off(from_input ? in1 : out1, eoff);
                void* co0 = (void*)off(out0, eoff);
                void* co1 = (void*)off(out1, eoff);
                // number of [n][inner] blocks of this pass inside nu units
                const long long blocks = nu * (gr.unit_elems / (ps.n * ps.inner));
                const bool last = i + 1 == np;
                int rc = plan->prec == B2FFT_F32
                             ? launch_pass<float>(plan, ps, ci0, ci1, co0, co1, inverse, blocks, last, stream)
                             : launch_pass<double>(plan, ps, ci0, ci1, co0, co1, inverse, blocks, last, stream);
                if (rc) return rc;
            }
        }
        first_group = false;
    }
    return B2FFT_OK;
}

int b2fft_plan_set_output_blocks(b2fft_plan* plan, int nblocks, void* const* blk0, void* const* blk1,
                                 int64_t out_inner, int64_t out_outer_stride) {
    if (!plan) return fail(B2FFT_E_INVALID, "null plan");
    if (nblocks == 0) { plan->nblocks = 0; return B2FFT_OK; }
    if (nblocks < 0 || nblocks > B2_MAX_BLOCKS || (nblocks & (nblocks - 1)))
        return fail(B2FFT_E_INVALID, "nblocks must be a power of two <= %d", B2_MAX_BLOCKS);
    if (plan->passes.empty()) return fail(B2FFT_E_INVALID, "plan has no pass to re-layout");
    const bool split = plan->layout == B2FFT_SPLIT;
    if (!blk0 || (split && !blk1)) return fail(B2FFT_E_INVALID, "null block pointer table");
    const Pass& lastp = plan->passes.back();
    if (lastp.n % nblocks) return fail(B2FFT_E_INVALID, "axis length %lld not divisible by %d", lastp.n, nblocks);
    if (!registry().v[lastp.variant].blk && !registry().v[lastp.fallback].blk)
        return fail(B2FFT_E_UNSUPPORTED, "kernel %s has no destination-blocked store", registry().v[lastp.variant].name);
    for (int h = 0; h < nblocks; ++h) {
        if (!blk0[h] || (split && !blk1[h])) return fail(B2FFT_E_INVALID, "null block pointer %d", h);
        plan->blk0[h] = blk0[h];
        plan->blk1[h] = split ? blk1[h] : nullptr;
    }
    plan->nblocks = nblocks;
    plan->blk_out_inner = out_inner;
    plan->blk_out_outer_stride = out_outer_stride;
    return B2FFT_OK;
}

int b2fft_plan_set_input_blocks(b2fft_plan* plan, int nblocks, const void* const* blk0) {
    if (!plan) return fail(B2FFT_E_INVALID, "null plan");
    if (nblocks == 0) { plan->in_nblocks = 0; return B2FFT_OK; }
    if (nblocks < 0 || nblocks > B2_MAX_BLOCKS || (nblocks & (nblocks - 1)))
        return fail(B2FFT_E_INVALID, "nblocks must be a power of two <= %d", B2_MAX_BLOCKS);
    if (plan->passes.size() != 1 || plan->passes[0].axis != 0 || plan->passes[0].transposing)
        return fail(B2FFT_E_UNSUPPORTED, "source-blocked loads need a plan with exactly one pass, over the contiguous axis");
    if (plan->layout == B2FFT_SPLIT) return fail(B2FFT_E_UNSUPPORTED, "source-blocked loads handle the interleaved layout");
    if (!blk0) return fail(B2FFT_E_INVALID, "null block pointer table");
    if (plan->passes[0].n % nblocks) return fail(B2FFT_E_INVALID, "axis length %lld not divisible by %d", plan->passes[0].n, nblocks);
    const size_t align = plan->prec ? 16 : 8;
    for (int h = 0; h < nblocks; ++h) {
        if (!blk0[h] || ((uintptr_t)blk0[h] % align)) return fail(B2FFT_E_INVALID, "bad block pointer %d", h);
        plan->in_blk[h] = blk0[h];
    }
    plan->in_nblocks = nblocks;
    return B2FFT_OK;
}

int b2fft_plan_set_outer_split(b2fft_plan* plan, int64_t outer_div, int64_t in_stride_lo, int64_t in_stride_hi,
                               int64_t out_stride_lo, int64_t out_stride_hi) {
    if (!plan) return fail(B2FFT_E_INVALID, "null plan");
    if (outer_div < 0) return fail(B2FFT_E_INVALID, "outer_div must be >= 0");
    if (outer_div > 0 && plan->passes.empty()) return fail(B2FFT_E_INVALID, "plan has no pass");
    if (outer_div > 0 && plan->n_transposing > 0) return fail(B2FFT_E_UNSUPPORTED, "outer split on a four-step plan");
    plan->split_div = outer_div;
    plan->split_in_lo = in_stride_lo; plan->split_in_hi = in_stride_hi;
    plan->split_out_lo = out_stride_lo; plan->split_out_hi = out_stride_hi;
    return B2FFT_OK;
}

int b2fft_plan_set_progress(b2fft_plan* plan, void* counters, int64_t outer_per_chunk, int max_ctas, int64_t* target) {
    if (!plan) return fail(B2FFT_E_INVALID, "null plan");
    if (target) *target = 0;
    if (!counters) { plan->progress = nullptr; plan->progress_outer = 0; plan->progress_max_ctas = 0; return B2FFT_OK; }
    if (outer_per_chunk <= 0 || max_ctas < 0) return fail(B2FFT_E_INVALID, "outer_per_chunk must be > 0 and max_ctas >= 0");
    if (plan->passes.size() != 1) return fail(B2FFT_E_UNSUPPORTED, "progress counters need a single-pass plan");
    const b2::KernelVariant& kv = registry().v[plan->passes[0].variant];
    if (!kv.progress || !plan->fused_scratch) return fail(B2FFT_E_UNSUPPORTED, "kernel %s publishes no progress counters", kv.name);
    plan->progress = (unsigned*)counters;
    plan->progress_outer = outer_per_chunk;
    plan->progress_max_ctas = max_ctas;
    if (target) *target = outer_per_chunk * (plan->passes[0].inner / kv.W);
    return B2FFT_OK;
}

int b2fft_plan_set_max_ctas(b2fft_plan* plan, int max_ctas) {
    if (!plan) return fail(B2FFT_E_INVALID, "null plan");
    if (max_ctas < 0) return fail(B2FFT_E_INVALID, "max_ctas must be >= 0");
    plan->max_ctas = max_ctas;
    return B2FFT_OK;
}

int b2fft_plan_set_exchange_ctas(b2fft_plan* plan, int ctas_per_sm) {
    if (!plan) return fail(B2FFT_E_INVALID, "null plan");
    if (ctas_per_sm < 0) return fail(B2FFT_E_INVALID, "ctas_per_sm must be >= 0");
    int sms = 0;
    DeviceGuard guard(plan->device);
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, plan->device));
    plan->exchange_max_ctas = ctas_per_sm * sms;
    return B2FFT_OK;
}

int b2fft_mem_alloc(size_t bytes, int device, void** out) {
    if (!out) return fail(B2FFT_E_INVALID, "null argument");
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return fail(B2FFT_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(guard.err));
    CUDA_TRY(cudaMalloc(out, bytes));
    return B2FFT_OK;
}

int b2fft_mem_free(void* dptr) {
    CUDA_TRY(cudaFree(dptr));
    return B2FFT_OK;
}

int b2fft_ipc_export(void* dptr, unsigned char handle[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, dptr));
    memcpy(handle, &h, 64);
    return B2FFT_OK;
}

int b2fft_ipc_import(const unsigned char handle[64], int device, void** out) {
    if (!out) return fail(B2FFT_E_INVALID, "null argument");
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return fail(B2FFT_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(guard.err));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CUDA_TRY(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
    return B2FFT_OK;
}

int b2fft_ipc_release(void* dptr) {
    CUDA_TRY(cudaIpcCloseMemHandle(dptr));
    return B2FFT_OK;
}

int b2fft_plan_destroy(b2fft_plan* plan) {
    if (plan && plan->fused_scratch) {
        DeviceGuard guard(plan->device);
        cudaFree(plan->fused_scratch);
        plan->fused_scratch = nullptr;
    }
    delete plan;
    return B2FFT_OK;
}

int b2fft_stream_synchronize(void* cuda_stream) {
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)cuda_stream));
    return B2FFT_OK;
}

int b2fft_plan_num_passes(const b2fft_plan* plan) { return plan ? (int)plan->passes.size() : B2FFT_E_INVALID; }

int b2fft_plan_describe(const b2fft_plan* plan, char* buf, size_t buflen) {
    if (!plan || !buf || !buflen) return fail(B2FFT_E_INVALID, "bad argument");
    std::string s;
    for (const Pass& p : plan->passes) {
        char line[256];
        if (p.transposing)
            snprintf(line, sizeof line, "axis=%c n=%lld inner=%lld variant=%s fs=%lldx%lld\n", "XYZ"[p.axis], p.n, p.inner,
                     registry().v[p.variant].name, p.n, p.fs_n2);
        else
            snprintf(line, sizeof line, "axis=%c n=%lld inner=%lld variant=%s\n", "XYZ"[p.axis], p.n, p.inner,
                     registry().v[p.variant].name);
        s += line;
    }
    snprintf(buf, buflen, "%s", s.c_str());
    return B2FFT_OK;
}

int b2fft_plan_preview(const int64_t dims_xyz[3], int axes_mask, int precision, int layout, char* buf, size_t buflen) {
    if (!dims_xyz || !buf || !buflen) return fail(B2FFT_E_INVALID, "bad argument");
    for (int a = 0; a < 3; ++a)
        if (!is_pow2(dims_xyz[a])) return fail(B2FFT_E_INVALID, "Array dimensions must be powers of two");
    if (precision != B2FFT_F32 && precision != B2FFT_F64) return fail(B2FFT_E_INVALID, "bad precision %d", precision);
    if (layout != B2FFT_INTERLEAVED && layout != B2FFT_SPLIT) return fail(B2FFT_E_INVALID, "bad layout %d", layout);
    if (axes_mask < 0 || axes_mask > 7) return fail(B2FFT_E_INVALID, "bad axes mask %d", axes_mask);
    b2fft_plan pl;
    pl.dry = true;
    pl.x = dims_xyz[0]; pl.y = dims_xyz[1]; pl.z = dims_xyz[2];
    pl.axes_mask = axes_mask; pl.prec = precision; pl.layout = layout; pl.device = -1;
    pl.normalize = 1; pl.fast_math = 1; pl.scale = 1.0; pl.norm_size = 1.0; pl.apply_scale = 1;
    int rc = build_passes(&pl);
    if (rc) return rc;
    return b2fft_plan_describe(&pl, buf, buflen);
}

int64_t b2fft_plan_launch_count(const b2fft_plan* plan) { return plan ? (int64_t)plan->launches.load() : -1; }

int b2fft_num_variants(void) { return (int)registry().v.size(); }

int b2fft_variant_info(int index, char* buf, size_t buflen) {
    Registry& reg = registry();
    if (index < 0 || index >= (int)reg.v.size() || !buf || !buflen) return fail(B2FFT_E_INVALID, "bad argument");
    const KernelVariant& k = reg.v[index];
    int occ = -1;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0 && reg.prepare(index) == cudaSuccess) occ = k.occupancy();
    else cudaGetLastError();
    snprintf(buf, buflen, "%s %d %d %d %d %d %d %d %lld %d %d", k.name, k.prec, k.log2n, k.W, k.G, k.E, k.S, k.threads,
             k.smem_bytes, k.minb, occ);
    return B2FFT_OK;
}

int b2fft_run_variant(int index, const void* in0, const void* in1, void* out0, void* out1, int split, int inverse,
                      int64_t n_tiles, int64_t inner, int device, void* cuda_stream) {
    Registry& reg = registry();
    if (index < 0 || index >= (int)reg.v.size()) return fail(B2FFT_E_INVALID, "bad variant index");
    const KernelVariant& k = reg.v[index];
    if (inner % k.W != 0) return fail(B2FFT_E_INVALID, "inner %lld not a multiple of W=%d", (long long)inner, k.W);
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return fail(B2FFT_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(guard.err));
    b2fft_plan tmp;   // only used for bookkeeping of launch_pass
    tmp.prec = k.prec; tmp.layout = split ? B2FFT_SPLIT : B2FFT_INTERLEAVED; tmp.apply_scale = 0; tmp.device = device;
    Pass p{};
    p.variant = index; p.fallback = index; p.log2n = k.log2n; p.n = 1LL << k.log2n; p.inner = inner;
    int rc0 = stage_tables(&tmp, k, p.n, p.tw);
    if (rc0) return rc0;
    cudaError_t e = reg.prepare(index);
    if (e != cudaSuccess) return fail(B2FFT_E_CUDA, "kernel attribute setup failed: %s", cudaGetErrorString(e));
    if (k.kind == 3) {   // tuning runs share one scratch buffer per device (never freed; single stream at a time)
        rc0 = fused_tables(&tmp, k, p);
        if (rc0) return rc0;
        static std::mutex mu;
        static std::map<int, std::pair<void*, size_t>> pool;
        std::lock_guard<std::mutex> lk(mu);
        const size_t need = (size_t)k.grid_slots() * (size_t)k.slot_elems * (k.prec ? 16 : 8);
        auto& ent = pool[device];
        if (ent.second < need) {
            if (ent.first) cudaFree(ent.first);
            CUDA_TRY(cudaMalloc(&ent.first, need));
            ent.second = need;
        }
        tmp.fused_scratch = ent.first;
        tmp.fused_scratch_bytes = ent.second;
    }
    const long long inner_blocks = inner / k.W;
    if (n_tiles % inner_blocks != 0) return fail(B2FFT_E_INVALID, "n_tiles must be a multiple of inner/W");
    const long long outer = n_tiles / inner_blocks;
    return k.prec == B2FFT_F32
               ? launch_pass<float>(&tmp, p, in0, in1, out0, out1, inverse, outer, false, (cudaStream_t)cuda_stream)
               : launch_pass<double>(&tmp, p, in0, in1, out0, out1, inverse, outer, false, (cudaStream_t)cuda_stream);
}

int b2fft_set_option(const char* key, double value) {
    if (!key) return fail(B2FFT_E_INVALID, "null key");
    if (!strcmp(key, "blk_bulk")) {
        if (value != 0 && value != 1 && value != 2) return fail(B2FFT_E_INVALID, "blk_bulk must be 0, 1 or 2");
        g_blk_bulk_mode.store((int)value);
        return B2FFT_OK;
    }
    if (!strcmp(key, "l2_chunk_bytes")) {
        if (value < 0) return fail(B2FFT_E_INVALID, "l2_chunk_bytes must be >= 0");
        g_l2_chunk_bytes.store((long long)value);
        return B2FFT_OK;
    }
    return fail(B2FFT_E_INVALID, "unknown option %s", key);
}

int b2fft_set_preferred_variants(const char* names) {
    registry().set_preferred(names);
    return B2FFT_OK;
}

}  // extern "C"
