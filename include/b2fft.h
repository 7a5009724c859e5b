/* b2fft -- C ABI of the B200-native batched C2C FFT library (libb2fft.so).
 *
 * This is the drop-in boundary for the reference's hot path.  The reference
 * (fjarri-attic/pyfft 0.3.9) has no FFI of its own: its boundary is the Python
 * operator API  pyfft.cuda.Plan(...) / plan.execute(...)  (pyfft/cuda.py:116-138,
 * pyfft/plan.py:70-109,173-284).  Each entry point below replaces one piece of that
 * path; pyfft_b200/plan.py binds them with ctypes and re-creates the Python API on top.
 *
 * Conventions: plain C, no torch / CUDA types in the signatures (streams and device
 * pointers travel as void*), 0 = success, negative = error (see B2FFT_E_*), message in
 * b2fft_last_error() (thread local).  b2fft_execute never allocates and never
 * synchronises: all kernels are enqueued on the given stream.
 */
#ifndef B2FFT_H_
#define B2FFT_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define B2FFT_API __attribute__((visibility("default")))
#else
#define B2FFT_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define B2FFT_OK 0
#define B2FFT_E_INVALID (-1)     /* bad argument: Python raises ValueError (pyfft/plan.py:24,48,87-89) */
#define B2FFT_E_CUDA (-2)        /* CUDA runtime error: RuntimeError */
#define B2FFT_E_UNSUPPORTED (-3) /* valid request this build has no kernel for: NotImplementedError */

#define B2FFT_F32 0
#define B2FFT_F64 1
#define B2FFT_INTERLEAVED 0 /* complex64 / complex128 arrays   (pyfft/plan.py:26-38, split = False) */
#define B2FFT_SPLIT 1       /* separate re / im real arrays    (pyfft/plan.py:26-38, split = True)  */

#define B2FFT_AXIS_X 1 /* contiguous (last numpy) axis */
#define B2FFT_AXIS_Y 2
#define B2FFT_AXIS_Z 4

typedef struct b2fft_plan b2fft_plan;

/* Library version, major*10000 + minor*100 + patch. */
B2FFT_API int b2fft_version(void);

/* Thread-local message of the last failing call on this thread ("" if none). */
B2FFT_API const char* b2fft_last_error(void);

/* Replaces FFTPlan.__init__ + _FFTParams + _generateKernelCode (pyfft/plan.py:16-133) and
 * the Mako render / nvcc compile behind them (pyfft/kernel.py:46-83): validates the shape
 * (every dimension a power of two, plan.py:23-24), picks the pass list (one DRAM round trip
 * per axis instead of plan.py:135-171's local/global chains) and uploads twiddle tables.
 *   dims_xyz   {x, y, z} with x the contiguous axis; unused axes = 1 (plan.py:73-89)
 *   rank       1, 2 or 3 (informational; axes of length 1 are no-ops)
 *   precision  B2FFT_F32 | B2FFT_F64,   layout  B2FFT_INTERLEAVED | B2FFT_SPLIT
 *   normalize, scale   semantics of _FFTKernel.getScaleCoeffFunc (pyfft/kernel.py:23-37):
 *                      forward *= scale (never normalised); inverse /= scale * (x*y*z if normalize)
 *   fast_math  1: scaling multiplies by a precomputed reciprocal; 0: true division as the
 *              reference's stores do (pyfft/kernel.mako:65,271-278).  Twiddles always come from
 *              tables of correctly rounded values in either mode.
 *   device     CUDA device ordinal the plan lives on. */
B2FFT_API int b2fft_plan_create(b2fft_plan** out, int rank, const int64_t dims_xyz[3], int precision, int layout,
                      int normalize, double scale, int fast_math, int device);

/* Same, but transforms only the axes in axes_mask (B2FFT_AXIS_* bits) and takes the
 * normalisation size explicitly (norm_size <= 0 means x*y*z); apply_scale = 0 skips the
 * scale/normalise step.  Used by the slab-decomposed multi-GPU 3D transform, where the
 * local X/Y passes and the Z pass after the all-to-all are separate plans (no reference
 * counterpart: the reference is single-device). */
B2FFT_API int b2fft_plan_create_ex(b2fft_plan** out, const int64_t dims_xyz[3], int axes_mask, int precision, int layout,
                         int normalize, double scale, int fast_math, int device, double norm_size,
                         int apply_scale);

/* Replaces the temp-buffer sizing of FFTPlan._execute (pyfft/plan.py:184-192).  The caller
 * (Python: torch caching allocator or the user's mempool) owns the workspace.  Plans whose axes all
 * fit one CTA (x <= 2^14 single / 2^13 double, y and z <= 2^11) need none.  Longer axes are split
 * four-step style into a transposing pass + the rest (the counterpart of the reference's
 * global-kernel chains, pyfft/plan.py:141-143); such plans need one data-sized buffer for in-place
 * executes (b2fft_plan_workspace_bytes) and, only when the long axis is not the first pass, for
 * out-of-place executes too (b2fft_plan_workspace_bytes_ex with in_place = 0).  The workspace must
 * be 16-byte aligned and stay valid until the execute has finished on its stream. */
B2FFT_API int b2fft_plan_workspace_bytes(const b2fft_plan* plan, int64_t batch, size_t* out);
B2FFT_API int b2fft_plan_workspace_bytes_ex(const b2fft_plan* plan, int64_t batch, int in_place, size_t* out);
B2FFT_API int b2fft_plan_set_workspace(b2fft_plan* plan, void* dptr, size_t bytes);

/* Replaces FFTPlan._execute + Function.__call__ (pyfft/plan.py:173-259, pyfft/cuda.py:35-46).
 *   interleaved: in0/out0 = complex arrays, in1/out1 = NULL
 *   split:       in0/in1 = re/im planes, out0/out1 = re/im planes
 *   out == in means in-place (plan.py:264-266,276-279); otherwise the input is left intact.
 *   batch transforms are stored back to back (doc/source/index.rst:254-255).
 *   cuda_stream is a cudaStream_t (NULL = legacy default stream).  Asynchronous. */
B2FFT_API int b2fft_execute(b2fft_plan* plan, const void* in0, const void* in1, void* out0, void* out1, int inverse,
                  int64_t batch, void* cuda_stream);

/* Slab-decomposed multi-GPU transforms (no reference counterpart).  Makes the
 * plan's LAST pass write its output destination-blocked: the transformed axis (length n) is cut
 * into nblocks equal blocks and output index k goes to block h = k / (n/nblocks), element
 *     blk0[h][ outer*out_outer_stride + inner_index + (k % (n/nblocks)) * out_inner ]
 * (blk1 = imaginary planes for the split layout).  Block pointers may be local send-buffer
 * chunks (followed by an NCCL all-to-all) or peer GPUs' receive buffers mapped over NVLink, in
 * which case the pass's stores ARE the exchange.  out0/out1 of b2fft_execute are then ignored
 * for that pass.  nblocks = 0 restores plain output. */
B2FFT_API int b2fft_plan_set_output_blocks(b2fft_plan* plan, int nblocks, void* const* blk0, void* const* blk1,
                                 int64_t out_inner, int64_t out_outer_stride);

/* Mirror image of b2fft_plan_set_output_blocks for the inverse slab transform: the plan (which must consist of one
 * contiguous-axis pass, interleaved layout) LOADS input index n of every line from block h = n / (n_axis/nblocks),
 *     blk0[h][ line_offset + n % (n_axis/nblocks) ]
 * with line_offset given by the two-level outer index of b2fft_plan_set_outer_split (input strides).  Block
 * pointers may be peer GPUs' buffers mapped over NVLink: the pass then pulls its lines together from the ranks
 * that hold the pieces and no separate exchange step exists.  in0 of b2fft_execute is ignored.  nblocks = 0 restores
 * plain input. */
B2FFT_API int b2fft_plan_set_input_blocks(b2fft_plan* plan, int nblocks, const void* const* blk0);

/* Slab exchange passes only.  Gives the plan's LAST pass a two-level outer index: its outer index o
 * (row number for a contiguous-axis pass, [n][inner] block number for a strided pass) is split as
 * (hi, lo) = (o / outer_div, o % outer_div) and the line/tile starts at
 *     input:  in  + hi*in_stride_hi  + lo*in_stride_lo
 *     output: out + hi*out_stride_hi + lo*out_stride_lo      (out = the block pointer when blocked)
 * in complex elements, instead of o*n*inner.  With it the X pass of the x-slab transform walks the
 * rows {all local z} x {one chunk of y} of a z-slab and scatters them as [y][z][x-block] into the
 * ranks' x-slabs.  outer_div = 0 restores the dense index. */
B2FFT_API int b2fft_plan_set_outer_split(b2fft_plan* plan, int64_t outer_div, int64_t in_stride_lo, int64_t in_stride_hi,
                               int64_t out_stride_lo, int64_t out_stride_hi);

/* Slab exchange passes only.  Caps the grid of the plan's LAST pass at ctas_per_sm CTAs per SM
 * (its CTAs then stride over the tiles).  A pass whose stores cross NVLink is link-bound; keeping it
 * to a sliver of each SM lets the HBM-bound pass of the previous chunk run beside it.  0 = no cap. */
B2FFT_API int b2fft_plan_set_exchange_ctas(b2fft_plan* plan, int ctas_per_sm);

/* Slab pipeline.  A plan that consists of ONE strided-axis pass run by a kernel that supports it (the streamed fused
 * two-step kernel, complex64 N = 2048 / 1024) publishes its progress: `counters` (device memory, 32-bit words, zeroed by
 * the caller before every execute) gets one word per chunk of `outer_per_chunk` consecutive outer blocks ([n][inner]
 * planes of the array); every finished tile adds 1 to its chunk's word with release semantics, so word k has reached
 * *target (= outer_per_chunk * inner / W, W the kernel's tile width) when chunk k is complete.  A consumer on another stream can then work on chunk k while
 * the launch continues behind it.  max_ctas > 0 caps the (persistent) grid, which leaves the remaining SMs to that
 * consumer.  B2FFT_E_UNSUPPORTED if the plan's kernel cannot do it; counters = NULL switches it off. */
B2FFT_API int b2fft_plan_set_progress(b2fft_plan* plan, void* counters, int64_t outer_per_chunk, int max_ctas, int64_t* target);

/* Caps the grid of every pass of the plan whose kernel walks its tiles with a grid stride (the plain and the fused
 * two-step kernels; the TMA-staged persistent kernels size their own grid and ignore it) at max_ctas CTAs, so that a
 * kernel on another stream finds free SMs beside it (slab pipeline: the Z passes leave room for the NVLink-bound X pass).
 * 0 = no cap. */
B2FFT_API int b2fft_plan_set_max_ctas(b2fft_plan* plan, int max_ctas);

/* Peer-visible device memory for the slab exchange: plain cudaMalloc'ed buffers whose CUDA IPC
 * handles (64 bytes) can be exchanged between the per-GPU processes (e.g. through
 * torch.distributed.all_gather_object) and opened on the other ranks. */
B2FFT_API int b2fft_mem_alloc(size_t bytes, int device, void** out);
B2FFT_API int b2fft_mem_free(void* dptr);
B2FFT_API int b2fft_ipc_export(void* dptr, unsigned char handle[64]);
B2FFT_API int b2fft_ipc_import(const unsigned char handle[64], int device, void** out);
B2FFT_API int b2fft_ipc_release(void* dptr);

B2FFT_API int b2fft_plan_destroy(b2fft_plan* plan);

/* ---- slab-decomposed 3-D transform over several GPUs (no reference counterpart: the
 * reference runs every kernel of a plan on one stream of one device, pyfft/plan.py:204-245).
 *
 * One b2fft_slab_plan per rank (= per GPU; one process per GPU, or one process driving several devices).  Rank g
 * owns z in [g*Z/G, (g+1)*Z/G) of the (Z, Y, X) array as the z-slab [Z/G][Y][X]; b2fft_slab_forward leaves rank h
 * with x in [h*X/G, (h+1)*X/G) as the x-slab [Y][Z][X/G] (one exchange, result stays distributed);
 * b2fft_slab_inverse goes back.  The exchange is fused into an FFT pass over peer-mapped memory (forward: the X
 * pass stores its rows as G contiguous pieces into the ranks' x-slabs; inverse: the X pass pulls them), ranks
 * synchronise through epoch words in peer-mapped memory, and the passes of neighbouring chunks overlap the NVLink
 * traffic on plan-owned streams -- see pyfft_b200/csrc/slab.cu.  Scaling follows pyfft/kernel.py:23-37 with
 * size = X*Y*Z.  Interleaved layout only.
 *
 *   y_chunks / z_chunks      pipeline depth of the exchange (0 = defaults: 8 y-chunks; 8 z-chunks when the Y pass can
 *                            be hidden under the exchange -- see b2fft_slab_plan_set_overlap -- else 1)
 *   exchange_ctas_per_sm     grid cap of the NVLink-bound X pass (0 = none), see b2fft_plan_set_exchange_ctas
 * The caller provides the buffers (b2fft_slab_plan_sizes; e.g. b2fft_mem_alloc + b2fft_ipc_export/import between
 * processes, or plain cudaMalloc + cudaDeviceEnablePeerAccess inside one process) and passes, for EVERY rank r, the
 * address under which rank r's x-slab and flag buffer are reachable from this plan's device.  Flag buffers must be
 * attached before any rank calls forward/inverse.  All ranks must issue the same sequence of forward/inverse calls.
 * Both calls are asynchronous on `cuda_stream`. */
typedef struct b2fft_slab_plan b2fft_slab_plan;
B2FFT_API int b2fft_slab_plan_create(b2fft_slab_plan** out, const int64_t dims_xyz[3], int precision, int normalize,
                                     double scale, int fast_math, int device, int rank, int nranks, int y_chunks,
                                     int z_chunks, int exchange_ctas_per_sm);
/* The order of the X-pass launches b2fft_slab_forward would issue for this pipeline shape, without touching a device:
 * "k:c;" = rows {z in chunk k} x {y in chunk c}, "k:0-n;" = {z in chunk k} x {y-chunks 0..n}, "all:c;" = {all local z} x
 * {y in chunk c}.  hidden_y = the Y pass runs as one launch with progress counters (b2fft_slab_plan_set_overlap);
 * overlap_columns = 0 picks the default for `nranks`. */
B2FFT_API int b2fft_slab_schedule_preview(int nranks, int y_chunks, int z_chunks, int hidden_y, int overlap_columns, char* buf, size_t buflen);
B2FFT_API int b2fft_slab_plan_sizes(const b2fft_slab_plan* plan, size_t* slab_bytes, size_t* xslab_bytes, size_t* flag_bytes);
/* {Z/G, X/G, y_chunks, z_chunks, Y/y_chunks, (Z/G)/z_chunks, G, rank} */
B2FFT_API int b2fft_slab_plan_geometry(const b2fft_slab_plan* plan, int64_t out[8]);
B2FFT_API int b2fft_slab_plan_attach(b2fft_slab_plan* plan, void* slab, void* const* xslab_of_rank, void* const* flags_of_rank);
/* Hides the local Y pass under the exchange (forward transform, z_chunks > 1): the Y pass becomes ONE persistent launch
 * over the whole slab that leaves `reserved_sms` SMs free and publishes a progress counter per z-chunk
 * (b2fft_plan_set_progress); the X passes of z-chunk k -- whose stores are the exchange -- start on the free SMs as soon
 * as chunk k is through its Y pass, instead of sharing every SM with the Y kernels.  reserved_sms = 0 switches it off
 * (one Y launch per z-chunk, ordered by events), -1 picks the default (on for >= 4 ranks when the Y kernel supports
 * progress counters, B2FFT_SLAB_OVERLAP_SMS overrides).  Returns B2FFT_E_UNSUPPORTED (and stays off) when the Y pass of
 * these dimensions cannot publish progress.  Call before the first forward. */
B2FFT_API int b2fft_slab_plan_set_overlap(b2fft_slab_plan* plan, int reserved_sms);
/* Tuning knobs of the pipeline.  "overlap_columns": how many y-chunks are sent z-chunk by z-chunk while the hidden Y launch
 * is still running (the rest go out one y-chunk at a time over all z); 0 = default (the fraction 0.364*G/(G-1) of the y-chunks,
 * i.e. as long as the Y launch takes: 3 of 8 on 8 GPUs, 4 of 8 on 4). */
B2FFT_API int b2fft_slab_plan_set_option(b2fft_slab_plan* plan, const char* key, double value);
B2FFT_API int b2fft_slab_forward(b2fft_slab_plan* plan, void* cuda_stream);
B2FFT_API int b2fft_slab_inverse(b2fft_slab_plan* plan, void* cuda_stream);
/* 0 = healthy; non-zero = a cross-rank wait timed out (a peer never signalled); synchronises the device */
B2FFT_API int b2fft_slab_plan_status(b2fft_slab_plan* plan, int* out);
B2FFT_API int64_t b2fft_slab_plan_launch_count(const b2fft_slab_plan* plan);
/* Timeline of the exchange pipeline (measurement aid): with tracing on, b2fft_slab_forward records a timing event at
 * every phase boundary; b2fft_slab_plan_trace synchronises the device and writes "name:ms;..." (completion time of the
 * Y pass of z-chunk k, the X pass of (z-chunk, y-chunk), the arrival of y-chunk c from all peers, the Z pass of c),
 * relative to the start of the last forward call. */
B2FFT_API int b2fft_slab_plan_set_trace(b2fft_slab_plan* plan, int on);
B2FFT_API int b2fft_slab_plan_trace(b2fft_slab_plan* plan, char* buf, size_t buflen);
B2FFT_API int b2fft_slab_plan_describe(const b2fft_slab_plan* plan, char* buf, size_t buflen);
B2FFT_API int b2fft_slab_plan_destroy(b2fft_slab_plan* plan);
B2FFT_API const char* b2fft_slab_last_error(void);

/* Replaces Context.wait (pyfft/cuda.py:98-101): blocks until the stream has drained.  Only
 * needed by callers that hold a raw cudaStream_t; torch callers synchronise their own stream. */
B2FFT_API int b2fft_stream_synchronize(void* cuda_stream);

/* ---- introspection (used by bench.py / tests) ---- */
B2FFT_API int b2fft_plan_num_passes(const b2fft_plan* plan);
/* One line per pass: "axis=X n=4096 inner=1 variant=float_n12_w1_..." */
B2FFT_API int b2fft_plan_describe(const b2fft_plan* plan, char* buf, size_t buflen);
/* The pass list b2fft_plan_create would build for these dimensions (same text as
 * b2fft_plan_describe), without touching a device: which axes run as one pass, which are split
 * four-step style ("fs=N1xN2" marks the transposing pass) and which kernel variant each pass uses. */
B2FFT_API int b2fft_plan_preview(const int64_t dims_xyz[3], int axes_mask, int precision, int layout, char* buf, size_t buflen);
/* Number of kernels this plan has launched since creation (bench.py's gpu_launches). */
B2FFT_API int64_t b2fft_plan_launch_count(const b2fft_plan* plan);

/* ---- kernel-variant tuning interface ---- */
B2FFT_API int b2fft_num_variants(void);
/* "name prec log2n W G E S threads smem_bytes minb occupancy" */
B2FFT_API int b2fft_variant_info(int index, char* buf, size_t buflen);
/* Runs ONE pass of variant `index` over n_tiles tiles of an [outer][N][inner] array. */
B2FFT_API int b2fft_run_variant(int index, const void* in0, const void* in1, void* out0, void* out1, int split,
                      int inverse, int64_t n_tiles, int64_t inner, int device, void* cuda_stream);
/* Comma-separated variant names that take precedence over the default preference order when
 * plans are created afterwards ("" clears).  Also read from $B2FFT_PREFER at load time. */
B2FFT_API int b2fft_set_preferred_variants(const char* names);

/* Global tuning options.  "blk_bulk": how the destination-blocked stores of a contiguous-axis pass (slab exchange) leave
 * the SM: 0 = warp stores, 1 = TMA bulk copies staged in the exchange buffer, 2 = bulk copies from a staging buffer of
 * their own, so that they drain while the CTA transforms its next lines (also $B2FFT_BLK_BULK).  "l2_chunk_bytes": multi-pass transforms run their passes chunk by
 * chunk so the intermediate stays in the 126 MB L2 between passes (bytes; default 0 = whole-array
 * passes, also settable through $B2FFT_L2_CHUNK_MB). */
B2FFT_API int b2fft_set_option(const char* key, double value);

#ifdef __cplusplus
}
#endif
#endif /* B2FFT_H_ */
