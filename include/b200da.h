/*
 * b200da.h -- C ABI of libb200da.so, the B200-native execution backend for the
 * data-parallel hot path of dask-array.
 *
 * The reference (mrocklin/dask-array) has NO FFI for this path: every hot loop is a
 * per-block Python call of a NumPy function.  Each entry point below therefore cites
 * the reference *Python* interface it replaces (file:line under /root/reference/).
 * The binding a maintainer would add on the reference side (ctypes stub registered
 * through dask_array/_chunk_types.py:31 and dask_array/_dispatch.py:145-151) is shown
 * in INTEGRATION.md.
 *
 * Conventions
 *  - Every function returns 0 on success or a negative b2_status; the message for the
 *    calling thread is returned by b2_last_error().  No C++ exception crosses the ABI.
 *  - The caller owns every buffer (device memory allocated by the host runtime, e.g.
 *    torch.empty) and passes raw device pointers.  The library never allocates output
 *    memory; scratch comes from a caller-provided workspace.
 *  - All launches are asynchronous on the caller's stream (a cudaStream_t passed as
 *    void*); the library never synchronises the device.
 *  - The current CUDA device is the caller's (cudaSetDevice before the call).
 *  - Re-entrant; the only shared state is the JIT module cache (mutex protected).
 *  - There is NO CPU fallback: without a CUDA device every compute entry point fails
 *    with B2_ERR_CUDA.
 */
#ifndef B200DA_H
#define B200DA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2_ABI_VERSION 1
#define B2_MAX_IN 6      /* external inputs of one fused kernel                         */
#define B2_MAX_ND 4      /* dims kept for arg-reduction index bookkeeping               */

typedef enum b2_status {
    B2_OK = 0,
    B2_ERR_INVALID = -1,      /* bad argument                                            */
    B2_ERR_CUDA = -2,         /* CUDA runtime / driver error (includes "no device")      */
    B2_ERR_NVRTC = -3,        /* JIT compilation failed; log in b2_last_error()          */
    B2_ERR_UNSUPPORTED = -4,  /* layout / dtype combination not implemented              */
    B2_ERR_WORKSPACE = -5     /* caller workspace too small                              */
} b2_status;

/* element types (NumPy kinds the reference's chunk kernels see) */
typedef enum b2_dtype {
    B2_BOOL = 0, B2_I8 = 1, B2_U8 = 2, B2_I16 = 3, B2_U16 = 4, B2_I32 = 5, B2_U32 = 6,
    B2_I64 = 7, B2_U64 = 8, B2_F32 = 9, B2_F64 = 10, B2_BF16 = 11, B2_F16 = 12
} b2_dtype;

/* reduction operators: the chunk / combine / aggregate kernels of
 * dask_array/reductions/_common.py:57-167 (sum/min/max), :270-320 (mean),
 * :368-505 (moment -> var/std), :675-750 (arg reductions). */
typedef enum b2_redop {
    B2_RED_NONE = 0, B2_RED_SUM = 1, B2_RED_MIN = 2, B2_RED_MAX = 3,
    B2_RED_ARGMIN = 4, B2_RED_ARGMAX = 5, B2_RED_MOMENT = 6,
    B2_RED_PROD = 7, B2_RED_ANY = 8, B2_RED_ALL = 9,
    B2_RED_NANMIN = 10, B2_RED_NANMAX = 11   /* np.nanmin / np.nanmax: NaN only if every element is NaN */
} b2_redop;

/* which canonical axes of a (B, R, C) block are reduced */
typedef enum b2_redmode {
    B2_MODE_EW = 0,   /* element-wise, full (B,R,C) output                              */
    B2_MODE_R = 1,    /* reduce rows   -> (B, C)   e.g. x.sum(axis=0) of a 2-D block      */
    B2_MODE_C = 2,    /* reduce cols   -> (B, R)   e.g. x.argmax(axis=1)                  */
    B2_MODE_RC = 3,   /* reduce both   -> (B,)     e.g. x.std() chunk step               */
    /* cumulative scans (reductions/_cumulative.py:100-265 CumReduction): full (B,R,C) output of the
     * accumulator type; `out1` = carry of the preceding blocks (NULL for the first block)          */
    B2_MODE_SR = 4,   /* scan along rows, columns kept: carry[b*C + c]   e.g. x.cumsum(axis=0)   */
    B2_MODE_SC = 5    /* scan along columns, rows kept:  carry[b*R + r]   e.g. x.cumsum(axis=1)   */
} b2_redmode;

/* what the last level of a tree reduction applies after combining
 * (mean_agg _common.py:308-320, moment_agg :456-505, std's sqrt :625-653) */
typedef enum b2_post {
    B2_POST_NONE = 0, B2_POST_MEAN = 1, B2_POST_VAR = 2, B2_POST_STD = 3
} b2_post;

/*
 * One output block of a fused launch, canonicalised by the host to a 3-D view
 * (B, R, C) with C innermost.  Strides are in ELEMENTS of the operand's dtype;
 * 0 = broadcast (creation/_utils.py:65-72 zero-stride constant blocks, NumPy
 * broadcasting inside Elemwise, _blockwise.py:1030-1074); swapped strides express
 * np.transpose views (manipulation/_transpose.py:14-75).
 * Replaces: the per-block argument tuple of FusedBlockwise._task
 * (_blockwise.py:1697-1728).
 */
typedef struct b2_block {
    const void* in[B2_MAX_IN];
    int64_t in_sb[B2_MAX_IN];
    int64_t in_sr[B2_MAX_IN];
    int64_t in_sc[B2_MAX_IN];
    void* out0;            /* EW: (B,R,C) contiguous.  Reductions: values / totals       */
    void* out1;            /* arg reductions: int64 indices; moment: unused              */
    int64_t B, R, C;
    int64_t tile_begin;    /* exclusive prefix sum of this block's tiles in the launch   */
    int64_t tiles_r, tiles_c; /* tile grid of this block (filled by b2_fused_plan)      */
    void* work;            /* two-stage partials for this block (from the workspace)     */
    unsigned int* counter; /* arrival counters for this block (from the workspace)       */
    /* arg reductions (arg_chunk, _common.py:704-732): offset along the reduced axis,
     * or -- ravel mode -- the block's N-d shape/offset inside the whole array.         */
    int64_t arg_offset;
    int32_t arg_ndim;      /* 0 = plain axis mode; >0 = ravel mode with the fields below */
    int32_t mirror;        /* mirror-pair kernels (f(x, x.T), manipulation/_transpose.py:14-75 feeding an
                            * Elemwise): table index of the transposed partner block, else 0         */
    int64_t arg_shape[B2_MAX_ND];
    int64_t arg_start[B2_MAX_ND];
    int64_t arg_total[B2_MAX_ND];
} b2_block;

/* scalar operands of the chain (Python scalars of elemwise(), _blockwise.py:1030) */
typedef struct b2_scalars {
    double f[8];
    int64_t i[8];
} b2_scalars;

/* ------------------------------------------------------------------ housekeeping */
int b2_abi_version(void);
const char* b2_last_error(void);
/* number of kernels this library has launched in this process (bench `gpu_launches`) */
int64_t b2_launch_count(void);
/* SM count of the current device (grid sizing) */
int b2_device_sm_count(int* out);

/* ------------------------------------------------------------------ fused kernels
 * FusedBlockwise (_blockwise.py:1574-1738) + the reduction chunk step
 * (reductions/_reduction.py:154-226): one JIT-compiled kernel per fused expression,
 * one launch per (expression, device) covering every resident block.               */
typedef struct b2_kernel b2_kernel;

/* Compile CUDA C++ `source` (produced by the host-side generator; it includes the
 * library's own device header "b2_device.cuh", supplied in-memory) for sm_100a and
 * return the cubin.  Works without a GPU.  *cubin is malloc'ed; free with b2_free. */
int b2_jit_compile(const char* source, const char* name, void** cubin, size_t* cubin_size);
void b2_free(void* p);
/* text of the device header (for diagnostics / offline nvcc builds) */
const char* b2_device_header(void);

/* compile-time geometry the generator baked into a kernel (must match its #defines) */
typedef struct b2_geom {
    int32_t mode;          /* b2_redmode                                                  */
    int32_t redop;         /* b2_redop                                                    */
    int32_t vec;           /* elements per thread per load (B2_V)                         */
    int32_t tx, ty;        /* thread layout: tx over columns, ty over rows (B2_TX, B2_TY) */
    int32_t rpt;           /* rows per tile (B2_RPT)                                      */
    int32_t packed_bytes;  /* sizeof of one two-stage partial (0 for EW / mode C)         */
    int32_t _pad;
} b2_geom;

/* Load a cubin on the current device and look up `entry` (needs a GPU). */
int b2_kernel_load(const void* cubin, size_t cubin_size, const char* entry,
                   const b2_geom* geom, b2_kernel** out);
int b2_kernel_free(b2_kernel* k);

/* Fill tile_begin / tiles_r / tiles_c / work / counter of every block and return the
 * launch geometry.  `blocks` is HOST memory; workspace pointers are device pointers
 * carved from [workspace, workspace + workspace_bytes).  Returns B2_ERR_WORKSPACE and
 * the required size in *needed when the workspace is too small (call with NULL to
 * query). */
int b2_fused_plan(const b2_kernel* k, b2_block* blocks, int nblocks,
                  void* workspace, size_t workspace_bytes, size_t* needed,
                  int64_t* total_tiles);

/* Launch over `nblocks` blocks whose descriptors (already planned) live in DEVICE
 * memory at d_blocks.  The counters region of the workspace must be zero on entry;
 * the kernel leaves it zero on exit (self-resetting), so it is reusable.            */
int b2_fused_launch(const b2_kernel* k, const b2_block* d_blocks, int nblocks,
                    int64_t total_tiles, const b2_scalars* scalars, void* stream);

/* ------------------------------------------------------------------ tree levels
 * PartialReduce (reductions/_reduction.py:900-983): one level of the
 * chunk -> combine -> aggregate tree.  For every output element e in [0, nelem) the
 * `fanin` partials are folded IN THE GIVEN ORDER (the lol_tuples nesting order of
 * :968-983).  parts[g] / parts1[g] are device pointers listed in a DEVICE table.
 *  - SUM/MIN/MAX/PROD/ANY/ALL: parts[g] -> dtype values.
 *  - ARGMIN/ARGMAX: parts[g] values, parts1[g] int64 indices (_arg_combine :675-701).
 *  - MOMENT: parts[g] -> packed (n, mean, M2) fp64 triples (Chan merge of
 *    moment_combine :415-453 carried in fp64).
 * post != NONE finishes mean/var/std (count = elements per output, ddof as given) and
 * writes `out_dtype` values.                                                        */
int b2_combine(int redop, int dtype, const void* const* d_parts, const void* const* d_parts1,
               int fanin, int64_t nelem, void* out0, void* out1,
               int post, int out_dtype, double count, double ddof, void* stream);

/* All output blocks of ONE PartialReduce level in a single launch: `ngroups` groups, group g
 * folding `fanin` partials of `nelem` elements each (same semantics as b2_combine).  The group
 * table lives in DEVICE memory; parts / parts1 point into a device pointer table. */
typedef struct b2_group {
    const void* const* parts;
    const void* const* parts1;
    void* out0;
    void* out1;
    int64_t nelem;
    int64_t elem_begin;   /* exclusive prefix sum of nelem over the groups */
    int32_t fanin;
    int32_t post;         /* b2_post */
    double count;
    double ddof;
} b2_group;
int b2_combine_groups(int redop, int dtype, int out_dtype, const b2_group* d_groups, int ngroups,
                      int64_t total_elems, void* stream);

/* ------------------------------------------------------------------ data movement
 * Rechunk / slicing / concatenation (_rechunk.py:1252-1323 split+merge tasks,
 * _chunk.py:285-317 getitem, _core_utils.py:1182-1248 concatenate3) as ONE tiled
 * gather: each descriptor copies a (rows x row_bytes) rectangle.                     */
typedef struct b2_copy {
    const void* src;
    void* dst;
    int64_t rows;
    int64_t row_bytes;
    int64_t src_pitch;   /* bytes between consecutive rows */
    int64_t dst_pitch;
    /* filled by b2_gather_plan: */
    int64_t tile_begin;  /* exclusive prefix sum of tiles                                 */
    int32_t tile_rows;   /* rows per tile                                                 */
    int32_t tiles_c;     /* column tiles per row band (each B2_GATHER_COL_BYTES wide)     */
    int32_t vec_bytes;   /* widest aligned access: 16, 8, 4, 2 or 1                       */
    int32_t _pad;
} b2_copy;
#define B2_GATHER_COL_BYTES 4096

int b2_gather_plan(b2_copy* copies, int n, int64_t* total_tiles);
int b2_gather_launch(const b2_copy* d_copies, int n, int64_t total_tiles, void* stream);
/* same plan, moved with TMA bulk copies (cp.async.bulk global->shared->global, no register staging).
 * Requires vec_bytes == 16 for EVERY rectangle (all addresses, pitches and row lengths 16-byte
 * multiples) and tile_rows * min(row_bytes, B2_GATHER_COL_BYTES) <= 64 KiB (what b2_gather_plan makes). */
int b2_gather_launch_bulk(const b2_copy* d_copies, int n, int64_t total_tiles, void* stream);

/* Top-k selection along the contiguous axis (chunk.topk / chunk.argtopk and their aggregates,
 * _chunk.py:200-290; routines/_topk.py:14-80).  Every row of `rows` x `n` elements is cut into segments
 * of `seg` elements (seg <= B2_TOPK_SEG_BYTES / (itemsize <= 4 ? 8 : 12), a power of two); one CTA sorts
 * one segment in shared memory (bitonic network on (value, index) pairs; NaN sorts as the largest value,
 * like np.sort) and writes its first min(k, segment length) entries, best first, to
 *     out_vals[row * out_pitch + s * kk + j],  out_idx[...] = in_idx ? in_idx[row * n + e] : idx_offset + e
 * with kk = min(k, seg).  Calling it again on the candidates (n' = nseg * kk, in_idx = the previous out_idx)
 * until one segment remains gives the k best of every row, sorted.  largest != 0: the k largest, descending;
 * else the k smallest, ascending.  Segments shorter than kk are padded with out_idx = -1 entries, which
 * sort last and must be ignored by the caller (`valid` counts come from the lengths).                   */
#define B2_TOPK_SEG_BYTES 32768
int b2_topk_rows(int dtype, const void* src, int64_t rows, int64_t n, int64_t src_pitch, int seg, int k,
                 int largest, void* out_vals, int64_t* out_idx, int64_t out_pitch, const int64_t* in_idx,
                 int64_t idx_offset, void* stream);

/* ---- peer memory (one process per GPU) ------------------------------------------
 * The reference moves blocks between workers by pickling them through the scheduler
 * (rechunk: _rechunk.py:1171-1323 getitem + concatenate3 tasks; transposed reads:
 * manipulation/_transpose.py:66-75).  Here a rank exports the allocation behind a block once
 * (CUDA IPC), the other ranks map it, and the SAME gather / fused kernels then store to or load
 * from the peer's HBM over NVLink: the all-to-all of a rechunk is the rechunk kernel itself.
 * b2_ipc_export: handle of the allocation containing `ptr` + the offset of `ptr` in it.
 * b2_ipc_open:   map it in this process (cached per allocation) and return the peer's `ptr`.   */
typedef struct b2_ipc_handle {
    unsigned char reserved[64];   /* cudaIpcMemHandle_t */
    int64_t offset;               /* ptr - allocation base */
    int64_t size;                 /* allocation size */
} b2_ipc_handle;
int b2_ipc_export(const void* ptr, b2_ipc_handle* out);
int b2_ipc_open(const b2_ipc_handle* h, void** ptr);
/* Stream-ordered barrier over all ranks through peer memory: rank `me` stores `epoch` into slot `me`
 * of every rank's signal array (d_sig_table[p] = rank p's array of `world` uint64, mapped here) and
 * waits until its own slots reach `epoch`.  Everything enqueued before it on any rank has completed,
 * and is visible, before anything enqueued after it runs.  Epochs must increase by one per call, in
 * the same order on every rank.  A peer that does not arrive within 120 s traps the kernel (a loud
 * failure instead of a hung GPU). */
int b2_peer_barrier(void* const* d_sig_table, int me, int world, uint64_t epoch, void* stream);
/* The same barrier with the epoch kept in a DEVICE counter (*d_epoch, this rank's own memory, zero at
 * start) that the kernel advances by one: the launch arguments never change, so the launch can be
 * captured in a CUDA graph and replayed.  All barriers of one signal table must use the same variant. */
int b2_peer_barrier_dev(void* const* d_sig_table, uint64_t* d_epoch, int me, int world, void* stream);
/* All-gather of the per-block partials of a tree reduction across ranks (the exchange step between
 * the chunk Blockwise and the first PartialReduce level, reductions/_reduction.py:751-806, when the
 * members of a group live on different GPUs) in ONE launch through peer memory: barrier (the receive
 * windows are free) -> this rank's `nbytes` stored into slot `me` (byte offset me * slot_bytes) of
 * every rank's window d_windows[p] over NVLink -> barrier (all records visible).  Advances *d_epoch
 * by two.  `send` and slot_bytes are 16-byte aligned; nbytes <= slot_bytes. */
int b2_peer_allgather(void* const* d_sig_table, uint64_t* d_epoch, void* const* d_windows, const void* send,
                      int64_t nbytes, int64_t slot_bytes, int me, int world, void* stream);

/* Strided host<->device block transfer: from_array's per-block getitem of a host array
 * (io/_from_array.py:60-160) and finalize's concatenate3 into the host result
 * (_core_utils.py:1426-1448), as one cudaMemcpy2DAsync per block. kind: 0 = H2D, 1 = D2H. */
int b2_memcpy2d(void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch,
                int64_t row_bytes, int64_t rows, int kind, void* stream);

/* fill a contiguous buffer with one element (Ones/Zeros/Full materialisation,
 * creation/_ones_zeros.py:17-137) */
int b2_fill(void* dst, int64_t nelem, int itemsize, const void* value, void* stream);

/* Sliding-window reductions: reduction(sliding_window_view(x, window), axis=window_axis) -- SlidingWindowReduction /
 * _sliding_window_banded_reduce (reductions/_sliding_window.py:96-160, 405-560) -- for an associative redop
 * (B2_RED_SUM / PROD / MIN / MAX; any / all are MAX / MIN on bool, mean = SUM with `mean` != 0: the result is
 * divided by `window` in the element type, sliding_window_finalize _reduction.py:499-502), over every block of
 * one launch.  Job i: src is (B, R, C) contiguous and already carries the window - 1 trailing halo elements
 * along the sliding axis: along_cols == 0 slides along R -> dst (B, R - window + 1, C); along_cols != 0 slides
 * along C -> dst (B, R, C - window + 1).  src and dst have the element type `dtype` (f32, f64, i32, i64, u8 /
 * bool).  `jobs` is a HOST array (its tile fields are filled here); `d_jobs` is a caller-owned DEVICE buffer of
 * njobs * sizeof(b2_window_job) bytes the table is copied to.  O(1) operations per element for any window
 * (two-direction scans over segments of `window` elements), 2 N bytes of DRAM traffic. */
typedef struct b2_window_job {
    const void* src;
    void* dst;
    int64_t B, R, C;
    int64_t src_pitch;                          /* along_cols only: elements between source rows (0 = C, dense) */
    int64_t tile_begin, col_tiles, row_tiles;   /* filled by the call */
} b2_window_job;
int b2_window_reduce_batched(int redop, int dtype, b2_window_job* jobs, int njobs, void* d_jobs,
                             int64_t window, int along_cols, int mean, void* stream);

/* Integer-array gather used by the arg-reduction combine step on the chunk type (_arg_combine,
 * reductions/_common.py:687-697: `vals.ravel()[local_args]` and the np.ogrid take-along-axis
 * `vals[ogrid.., local_args, ..]`).  inner == 0: out[j] = src[idx[j]], j < count, idx in [-n, n).
 * inner > 0: src is (outer, n, inner) contiguous, idx / out are (outer, inner): out[o,i] = src[o, idx[o,i], i];
 * count = outer * inner. */
int b2_take(int itemsize, const void* src, const int64_t* idx, void* out, int64_t count, int64_t n,
            int64_t inner, void* stream);

/* ------------------------------------------------------------------ contraction
 * matmul / tensordot block GEMM (linalg/_tensordot.py:194-249): C (+)= A @ B^T with
 * A (M,K) and B (N,K) row-major ("TN"), accumulating over the k block index instead
 * of materialising the (M,1,N) partials.  dtype B2_BF16: tcgen05 bf16 x bf16 -> fp32;
 * B2_F32: 3xBF16 split on tcgen05 (documented tolerance).                            */
int b2_gemm_tn(int dtype, const void* A, int64_t lda, const void* B, int64_t ldb,
               float* C, int64_t ldc, int64_t M, int64_t N, int64_t K,
               int accumulate, void* stream);
/* C (+)= sum_p A[p] @ B[p]^T in ONE accumulation (TMEM): the pair list runs over the contracted
 * block index k of `_sum_wo_cat` (linalg/_tensordot.py:216-249) -- the (M,1,N) partials of
 * `_matmul` (:194-213) are never materialised -- and over the bf16 x 3 split products when the
 * operands were fp32.  A[p], B[p]: HOST arrays of device pointers to bf16 planes. */
int b2_gemm_tn_pairs(int dtype, const void* const* A, const void* const* B, int npairs,
                     int64_t lda, int64_t ldb, float* C, int64_t ldc, int64_t M, int64_t N, int64_t K,
                     int accumulate, void* stream);
/* Every output block of a blocked matmul in ONE launch (no wave quantisation between blocks).
 * `problems` is a HOST array; the tensor maps and the problem table are written into the
 * caller's DEVICE workspace (64-byte aligned; call with workspace == NULL to get *needed). */
typedef struct b2_gemm_problem {
    const void* const* A;   /* host array of npairs device pointers, each (M, K) bf16, leading dim lda */
    const void* const* B;   /* host array of npairs device pointers, each (N, K) bf16, leading dim ldb */
    int32_t npairs;
    int32_t accumulate;
    int64_t lda, ldb;
    float* C;
    int64_t ldc;
    int64_t M, N, K;
    /* optional (may be NULL): ragged contraction blocks.  Pair p then has contraction length
     * Kpair[p] <= K and dense operands (leading dimensions Kpair[p]); the tail is zero-filled. */
    const int64_t* Kpair;
} b2_gemm_problem;
int b2_gemm_tn_batched(int dtype, const b2_gemm_problem* problems, int nproblems,
                       void* workspace, size_t workspace_bytes, size_t* needed, void* stream);
/* The same contraction for the NumPy number types the tensor cores do not serve -- float64 (the dtype
 * of the reference's own matmul / tensordot tests, tests/test_routines.py:321-399), float32 in IEEE
 * arithmetic, int32 / uint32 / int64 / uint64 -- on the CUDA cores, every product and sum in the
 * element type (np.matmul semantics per block, linalg/_tensordot.py:207).  C has the operands' dtype. */
int b2_gemm_tn_simt(int dtype, const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc,
                    int64_t M, int64_t N, int64_t K, int accumulate, void* stream);
/* fp32 -> bf16 hi/mid/lo planes with hi + mid + lo == x to ~2^-24 (operand preparation of the
 * fp32 matmul; tensor cores have no IEEE fp32 mode) */
int b2_split3_bf16(const float* src, void* hi, void* mid, void* lo, int64_t nelem, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200DA_H */
