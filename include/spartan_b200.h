/* spartan_b200 -- C ABI of the B200 (sm_100a) tile evaluator.
 *
 * This is the drop-in boundary for Spartan's per-tile hot path.  The reference
 * has no C ABI: its seam is the Python kernel contract
 *   mapper_fn(ex, **kw) -> LocalKernelResult        (spartan/core.pyx:159-169)
 * invoked per tile by Worker._run_kernel             (spartan/worker.py:232-304)
 * plus the tile store BlobCtx.create/get/update      (spartan/blob_ctx.py:18-284).
 * Every entry point below names the reference code it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; device pointers are raw CUDA addresses
 *     (the Python host obtains them from torch tensors, any allocator works);
 *   - every function returns 0 (SP_OK) or a negative sp_status; the message of
 *     the last failure on the calling thread is sp_last_error();
 *   - launches are asynchronous on `stream` (a cudaStream_t passed as void*;
 *     NULL = legacy default stream); nothing here synchronises the device;
 *   - there is NO CPU fallback: without a CUDA device the compute entry points
 *     fail with SP_ERR_CUDA.  The extent algebra (sp_extent_*) is host-only.
 */
#ifndef SPARTAN_B200_H_
#define SPARTAN_B200_H_

#ifdef __CUDACC_RTC__ /* compiled at run time by NVRTC (csrc/jit.cu): no system headers there */
typedef signed char int8_t;
typedef unsigned char uint8_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
#else
#include <stddef.h>
#include <stdint.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  SP_OK = 0,
  SP_ERR_INVALID = -1,      /* bad argument */
  SP_ERR_CUDA = -2,         /* CUDA runtime / driver error */
  SP_ERR_UNSUPPORTED = -3,  /* expression not mappable to the device evaluator */
  SP_ERR_NOMEM = -4
} sp_status;

typedef enum {
  SP_F32 = 0,
  SP_F64 = 1,
  SP_I32 = 2,
  SP_I64 = 3,
  SP_U8 = 4,
  SP_BOOL = 5 /* numpy bool_: one byte, 0 or 1 */
} sp_dtype;

#define SP_MAX_DIM 32 /* spartan/array/extent.pyx:20-21 MAX_DIM */

const char* sp_last_error(void);
int sp_version(void);
/* Housekeeping for hosts without their own allocator / stream library (the Python host uses torch for these):
 * sp_init selects the device and creates its context, sp_tile_alloc / sp_tile_free are stream-ordered HBM buffers
 * (BlobCtx.create / destroy, blob_ctx.py:221-254,181-196), sp_sync waits for a stream, sp_event_* time a region. */
int sp_init(int device);
int sp_shutdown(void);
int sp_tile_alloc(int64_t bytes, void** out, void* stream);
int sp_tile_free(void* p, void* stream);
int sp_sync(void* stream);
int sp_event_create(void** out);
int sp_event_record(void* event, void* stream);
int sp_event_elapsed(void* start, void* stop, float* ms);
int sp_event_destroy(void* event);
/* Device properties of the current CUDA device (fails without a GPU). */
int sp_device_info(int* n_sms, int64_t* hbm_bytes, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------
 * Extent algebra (host only, int64, bit-exact with spartan/array/extent.pyx).
 * An extent is (ul[ndim], lr[ndim]) inside an array of shape array_shape.
 * Functions that can produce an empty result return 1 when the result is
 * valid and 0 when it is "None" in the reference (any ul >= lr), <0 on error.
 * ------------------------------------------------------------------------ */
/* extent.pyx:367-387 intersection */
int sp_extent_intersection(int ndim, const int64_t* a_ul, const int64_t* a_lr, const int64_t* b_ul,
                           const int64_t* b_lr, int64_t* out_ul, int64_t* out_lr);
/* extent.pyx:207-219 ravelled_pos (C order) */
int64_t sp_extent_ravelled_pos(int ndim, const int64_t* idx, const int64_t* array_shape);
/* extent.pyx:196-205 unravelled_pos */
int sp_extent_unravelled_pos(int64_t idx, int ndim, const int64_t* array_shape, int64_t* out_idx);
/* extent.pyx:411-432 drop_axis; axis < 0 wraps; axis == SP_AXIS_NONE gives the 0-d extent (out_ndim = 0). */
#define SP_AXIS_NONE (-1000)
int sp_extent_drop_axis(int ndim, const int64_t* ul, const int64_t* lr, const int64_t* array_shape, int axis,
                        int64_t* out_ul, int64_t* out_lr, int64_t* out_shape, int* out_ndim);
/* extent.pyx:121-127 TileExtent.to_global */
int64_t sp_extent_to_global(int ndim, const int64_t* ul, const int64_t* lr, const int64_t* array_shape,
                            int64_t idx, int axis);
/* extent.pyx:501-570 change_partition_axis, one-dimensional target (the grid
 * branch :545-552 is a reference defect, SURVEY.md section 9 Q1: it is NOT
 * reproduced -- grid-tiled input returns SP_ERR_UNSUPPORTED).  1 = valid, 0 = None. */
int sp_extent_change_partition_axis(int ndim, const int64_t* ul, const int64_t* lr, const int64_t* array_shape,
                                    int axis, int64_t* out_ul, int64_t* out_lr);
/* distarray.py:26-48 good_tile_shape (Python-2 floor division semantics) */
int sp_good_tile_shape(int ndim, const int64_t* shape, int64_t num_shards, int64_t* out_tile_shape);
/* distarray.py:51-110 compute_splits + compute_extents: enumerates tiles in
 * itertools.product (row-major) order.  Call with out_ul == NULL to get the
 * tile count; otherwise fills out_ul/out_lr [n_tiles * ndim] and
 * out_worker[n_tiles] = index % num_shards (index if num_shards == -1). */
int64_t sp_compute_extents(int ndim, const int64_t* shape, const int64_t* tile_hint /* may be NULL */,
                           int64_t num_shards, int64_t* out_ul, int64_t* out_lr, int64_t* out_worker);

/* ------------------------------------------------------------------------
 * Tile fill (creation.py:67-106 _make_zeros/_make_ones, :135-141 _arange_mapper,
 * srandom.py:40-47 _make_rand/_make_randn) -- written straight into HBM.
 * ------------------------------------------------------------------------ */
typedef enum {
  SP_FILL_CONST = 0, /* dst[i] = a */
  SP_FILL_IOTA = 1,  /* dst[i] = a + b * (offset + i)   (arange: a=start, b=step) */
  SP_FILL_RAND = 2,  /* uniform [0,1): Philox4x32-10 keyed by seed, counter = offset + i */
  SP_FILL_RANDN = 3  /* standard normal (Box-Muller over the same Philox stream) */
} sp_fill_kind;
int sp_fill(void* dst, int dtype, int64_t n, int kind, double a, double b, uint64_t seed, int64_t offset,
            void* stream);
/* Strided 2-D form: element (r, c) of a tile takes stream index offset + r * index_row_pitch + c, so a tile that is a
 * sub-rectangle of its array regenerates exactly its share of the array-wide sequence (dst_row_stride in elements). */
int sp_fill2d(void* dst, int dtype, int64_t rows, int64_t cols, int64_t dst_row_stride, int kind, double a, double b,
              uint64_t seed, int64_t offset, int64_t index_row_pitch, void* stream);

/* ------------------------------------------------------------------------
 * Fused element-wise map (local.py:115-127 FnCallExpr.evaluate over the tree
 * MapMapFusion builds, optimize.py:133-187) and fused map+reduce
 * (reduce.py:21-70 _reduce_mapper with ReduceMapFusion, optimize.py:190-227).
 *
 * A program is the postfix bytecode of the LocalExpr tree.  Operands are
 * addressed through a common 3-D iteration space (d0, d1, d2), d2 innermost;
 * each operand has element strides (s0, s1, s2) with 0 = broadcast
 * (broadcast.py:28-158).  Scalars are SP_OP_CONST immediates.
 * ------------------------------------------------------------------------ */
typedef enum {
  /* leaves */
  SP_OP_IN = 0,    /* push operand #arg */
  SP_OP_CONST = 1, /* push immediate #arg (sp_program.consts[arg]) */
  SP_OP_INDEX = 2, /* push index_base + i0*index_stride[0] + i1*index_stride[1] + i2*index_stride[2]: the position of the
                      element (map_with_location.py:22-60; argmin/argmax, sorting.py:67-124) */
  /* binary: pop b, pop a, push f(a, b)      (base.py:331-388, mathematics.py, logic.py) */
  SP_OP_ADD = 8,
  SP_OP_SUB = 9,
  SP_OP_MUL = 10,
  SP_OP_DIV = 11,   /* np.divide on floats (true division); floor division on ints */
  SP_OP_MOD = 12,   /* np.mod (sign of divisor) */
  SP_OP_POW = 13,
  SP_OP_MAX = 14,   /* np.maximum (NaN propagating) */
  SP_OP_MIN = 15,
  SP_OP_EQ = 16,
  SP_OP_NE = 17,
  SP_OP_LT = 18,
  SP_OP_LE = 19,
  SP_OP_GT = 20,
  SP_OP_GE = 21,
  SP_OP_AND = 22,   /* logical_and */
  SP_OP_OR = 23,
  SP_OP_XOR = 24,
  SP_OP_FMOD = 25,  /* np.fmod (sign of dividend) */
  SP_OP_FLOORDIV = 26,
  /* unary: pop a, push f(a) */
  SP_OP_NEG = 40,
  SP_OP_ABS = 41,
  SP_OP_SQRT = 42,
  SP_OP_EXP = 43,
  SP_OP_LOG = 44,
  SP_OP_SQUARE = 45,
  SP_OP_RECIP = 46,
  SP_OP_NOT = 47,
  SP_OP_NONZERO = 48, /* x != 0 -> 1/0  (count_nonzero, np.all/any inputs) */
  SP_OP_ISZERO = 49,
  /* casts to a narrower type inside a wider compute type (arrays.py:26-35 astype and
   * NumPy per-ufunc result dtypes): value := (compute_t)(target_t)value */
  SP_OP_CAST_F32 = 56,
  SP_OP_CAST_I64 = 57,
  SP_OP_CAST_I32 = 58,
  SP_OP_CAST_BOOL = 59,
  SP_OP_CAST_U8 = 60
} sp_opcode;

#define SP_MAX_PROGRAM 64
#define SP_MAX_OPERANDS 8
#define SP_MAX_CONSTS 16
#define SP_MAX_STACK 4 /* evaluation-stack depth of one fused kernel (register resident) */

typedef struct {
  int32_t n_ops;
  int32_t compute_dtype; /* SP_F32, SP_F64 or SP_I64: the register type of the evaluation stack */
  uint8_t op[SP_MAX_PROGRAM];
  uint8_t arg[SP_MAX_PROGRAM];
  double consts[SP_MAX_CONSTS];   /* used when compute_dtype is SP_F32 / SP_F64 */
  int64_t iconsts[SP_MAX_CONSTS]; /* used when compute_dtype is SP_I64 */
  int64_t index_stride[3];        /* SP_OP_INDEX: coefficients of the iteration coordinates (d0, d1, d2) */
  int64_t index_base;
} sp_program;

typedef struct {
  const void* ptr; /* device pointer to element (0,0,0) of this operand's view */
  int32_t dtype;
  int32_t pad;
  int64_t stride[3]; /* in elements; 0 = broadcast along that dim */
} sp_operand;

/* out[d0,d1,d2] = program(in...), out written with strides out->stride, cast to out->dtype. */
int sp_map(const sp_program* prog, int n_in, const sp_operand* in, const sp_operand* out, const int64_t dims[3],
           void* stream);

/* Run-time specialisation (csrc/jit.cu): a fused chain outside the library's static catalogue is compiled ONCE, by
 * NVRTC for sm_100a, into a straight-line instance of the same streaming kernel -- the counterpart of the reference's
 * per-expression code generation (local.py:58-152 `codegen`).  Large launches only; the interpreter runs everything
 * else and is the fallback when libnvrtc is missing.  sp_jit_compile_check compiles (never loads) the specialisation of
 * `prog` -- usable without a GPU; returns the cubin size. */
int sp_jit_enable(int on);
int sp_jit_set_nvrtc_path(const char* path);
int sp_jit_stats(int64_t* compiled, int64_t* launches, int64_t* failures);
const char* sp_jit_last_log(void);
int64_t sp_jit_compile_check(const sp_program* prog, int n_in, int mode);

typedef enum {
  SP_RED_SUM = 0,  /* mathematics.py:126 _sum_local;  combiner np.add */
  SP_RED_MIN = 1,  /* statistics.py:59;               combiner np.minimum */
  SP_RED_MAX = 2,  /* statistics.py:40;               combiner np.maximum */
  SP_RED_PROD = 3, /* mathematics.py:146 _prod_local; combiner np.multiply */
  SP_RED_ALL = 4,  /* logic.py:25 (value != 0 folded with AND) */
  SP_RED_ANY = 5   /* logic.py:37 */
} sp_reduce_op;

/* Reduce program(in...) over d1 of the iteration space (d0 = outer, d1 = reduced
 * axis, d2 = inner):  out[d0, d2] = op_{d1} program(...)[d0, d1, d2].
 * axis=None is expressed as dims = (1, n, 1).  `out` strides are (s_outer, ignored, s_inner).
 * If accumulate != 0 the result is combined into the existing out values with the same
 * op (Tile.merge "reducer(old, new)", tile.pyx:263-283); otherwise it replaces them
 * (first write).  scratch: device buffer of at least sp_map_reduce_scratch_bytes(). */
int64_t sp_map_reduce_scratch_bytes(const int64_t dims[3], int compute_dtype);
int sp_map_reduce(const sp_program* prog, int n_in, const sp_operand* in, const sp_operand* out,
                  const int64_t dims[3], int reduce_op, int accumulate, void* scratch, int64_t scratch_bytes,
                  void* stream);

/* Cross-tile combiner on one device: dst = op(dst, src) element-wise over n contiguous
 * elements (Tile.merge dense path, tile.pyx:250-283).  Across GPUs the host uses
 * ncclAllReduce with the matching op instead. */
int sp_combine(void* dst, const void* src, int dtype, int64_t n, int reduce_op, void* stream);

/* A transposed view (transpose.py:27-67) made dense: dst (R x C, ld ldd) = transpose of src (C x R, ld lds), through a
 * shared-memory tile so both sides are coalesced.  elem_size 1, 4 or 8 bytes; C <= 65535 * 32 per call. */
int sp_transpose_2d(void* dst, int64_t ldd, const void* src, int64_t lds, int64_t R, int64_t C, int elem_size, void* stream);
/* Tile.merge on a partially written tile (tile.pyx:270-283): per element, dst = mask ? reducer(dst, src) : src, then
 * mask = 1.  reduce_op < 0: no reducer.  3-D strided views (strides in elements); the reduction is evaluated in the
 * NumPy result type of the two dtypes.  Pairs: equal dtypes, (f32,f64), (f64,f32), (i64,i32), (i32,i64). */
int sp_merge_masked(void* dst, const int64_t dst_stride[3], int dst_dtype, const void* src, const int64_t src_stride[3],
                    int src_dtype, uint8_t* mask, const int64_t mask_stride[3], const int64_t dims[3], int reduce_op,
                    void* stream);
/* Strided rectangle copy between tiles (DistArrayImpl.fetch stitching, distarray.py:294-367,
 * and update splitting, :372-422): copies dims[0..2] elements; strides in elements. */
int sp_copy_rect(void* dst, const int64_t dst_stride[3], const void* src, const int64_t src_stride[3],
                 const int64_t dims[3], int dtype, void* stream);

/* Host <-> device rectangle transfers (pitched DMA; asynchronous on `stream` when the host buffer is pinned).
 * Replaces the pickled-ndarray payloads of BlobCtx.get / update (blob_ctx.py:127-179) at the host boundary. */
int sp_upload_2d(void* dst_device, int64_t dst_pitch_bytes, const void* src_host, int64_t src_pitch_bytes,
                 int64_t width_bytes, int64_t rows, void* stream);
int sp_download_2d(void* dst_host, int64_t dst_pitch_bytes, const void* src_device, int64_t src_pitch_bytes,
                   int64_t width_bytes, int64_t rows, void* stream);

/* ------------------------------------------------------------------------
 * Dense contraction (dot.py:195-238 dot_map2_mapper / dot_outer_mapper:
 * `tiles[0].dot(tiles[1])`, partials merged with np.add).
 * ------------------------------------------------------------------------ */
typedef enum {
  SP_GEMM_TF32X1 = 0, /* one tcgen05 kind::tf32 pass over RN-rounded operands */
  SP_GEMM_TF32X3 = 1, /* hi/lo split, 3 passes: ~fp32-faithful (error ~2^-21 per product) */
  SP_GEMM_SIMT = 2,   /* CUDA-core reference path, exact per-dtype arithmetic (also f64 / i64 / i32) */
  SP_GEMM_BF16X3 = 3  /* bf16 hi/lo split, 3 kind::f16 passes at twice the tf32 rate (error ~2^-16 per product) */
} sp_gemm_precision;

#define SP_GEMM_MAX_TERMS 24
#define SP_GEMM_MAX_SEGMENTS 8 /* (A strip, B strip) pairs per launch */

typedef struct {
  const float* A; /* [M, K] row-major, leading dimension lda */
  int64_t lda;
  const float* B; /* [K, N] row-major, leading dimension ldb */
  int64_t ldb;
  int64_t K;
} sp_gemm_segment;

/* C[M,N] (+)= sum_s A_s[M,K_s] * B_s[K_s,N]: one launch for a whole strip-joined dot
 * (join_mapper, map.py:243-286).  accumulate != 0 adds into C (np.add combiner).
 * The tensor-core accumulator truncates, so partial sums are promoted to fp32 registers
 * (round-to-nearest) every few k-blocks; sp_gemm_set_chunk_kblocks overrides that interval
 * (0 = per-mode default) -- a tuning / test hook. */
int sp_gemm_set_chunk_kblocks(int k_blocks);
/* Test / tuning hook: kernel variant of the tensor-core paths. 0 = choose by shape (default), 1 = one CTA per
 * 128 x 256 tile, 2 = CTA pairs (tcgen05 cta_group::2) on 256 x 256 tiles. Results are identical bit for bit. */
int sp_gemm_set_variant(int variant);
/* Test / tuning hook: 1 = the CTA pairs of a launch check in at every tile round (bounded wait) so their K loops stay
 * within a few k-blocks of each other and shared operand tiles are served from L2; 0 = free running. */
int sp_gemm_set_round_sync(int on);
int sp_gemm_set_tuning(int sync_k_blocks, int group_m); /* extra check-ins inside a tile; tile rows per rasterisation group */
int64_t sp_gemm_f32_workspace_bytes(int64_t M, int64_t N, int n_seg, const int64_t* seg_k, int precision);
int sp_gemm_f32_segments(int n_seg, const sp_gemm_segment* segs, float* C, int64_t ldc, int64_t M, int64_t N,
                         int accumulate, int precision, void* workspace, int64_t workspace_bytes, void* stream);
/* Split form for multi-GPU dots: a rank prepares (rounds / splits / transposes) only the strips it owns, the
 * prepared strips are plain byte buffers that can be exchanged (ncclAllGather), and the contraction runs over
 * prepared operands.  Layout of a prepared operand: [copies][rows][Kp] elements (rows = M for A, N for B;
 * Kp = sp_gemm_kpad(K): depth padded to the k-block; 128-byte aligned). */
typedef struct {
  const void* A; /* prepared A strip */
  const void* B; /* prepared B strip (transposed) */
  int64_t Kp;
} sp_gemm_prepared_segment;
int64_t sp_gemm_kpad(int64_t K, int precision);
int64_t sp_gemm_prepared_bytes(int64_t rows, int64_t Kp, int precision);
int sp_gemm_prepare_a(const float* A, int64_t lda, int64_t M, int64_t K, int precision, void* out, int64_t Kp,
                      int64_t k_offset, int64_t out_bytes, void* stream);
int sp_gemm_prepare_b(const float* B, int64_t ldb, int64_t K, int64_t N, int precision, void* out, int64_t Kp,
                      int64_t k_offset, int64_t out_bytes, void* stream);
int sp_gemm_prepared(int n_seg, const sp_gemm_prepared_segment* segs, float* C, int64_t ldc, int64_t M, int64_t N,
                     int accumulate, int precision, void* stream);
/* Row-range forms for a pipelined dot (host strips arrive one by one over PCIe while earlier strips are being
 * contracted): a strip is prepared straight into its rows of a full-size prepared operand
 * [copies][rows_total][Kp], and a contraction addresses any contiguous row range of such an operand.
 * `out` / A / B point at the first row of the range inside copy 0; copy 1 (the lo halves of the split modes)
 * starts *_copy_stride bytes later (= rows_total * Kp * element bytes). */
typedef struct {
  const void* A;
  int64_t a_copy_stride;
  const void* B;
  int64_t b_copy_stride;
  int64_t Kp;
} sp_gemm_prepared_view;
int sp_gemm_prepare_a_rows(const float* A, int64_t lda, int64_t M, int64_t K, int precision, void* out,
                           int64_t copy_stride, int64_t Kp, int64_t k_offset, void* stream);
int sp_gemm_prepare_b_rows(const float* B, int64_t ldb, int64_t K, int64_t N, int precision, void* out,
                           int64_t copy_stride, int64_t Kp, int64_t k_offset, void* stream);
int sp_gemm_prepared_views(int n_seg, const sp_gemm_prepared_view* segs, float* C, int64_t ldc, int64_t M, int64_t N,
                           int accumulate, int precision, void* stream);
/* Gated form for the multi-GPU dot (dot.py:195-238; join_mapper's remote strip fetches, map.py:243-286): segment s is
 * read only after *ready_flag[s] has reached ready_value[s] (NULL flag = not gated).  Peers push operand strips into this
 * GPU's memory and write the flag behind them (sp_peer_push); the contraction runs on what is present and picks up the
 * rest as it lands, in one launch.  A gate closed for ~4 s is given up and counted in *gate_status (device word or NULL). */
int sp_gemm_prepared_views_gated(int n_seg, const sp_gemm_prepared_view* segs, const uint32_t* const* ready_flag,
                                 const uint32_t* ready_value, uint32_t* gate_status, float* C, int64_t ldc, int64_t M,
                                 int64_t N, int accumulate, int precision, void* stream);
/* Fused row-argmin epilogue over prepared operands (no C is stored): for every row and every 128-column half tile,
 * part_val[row][p] = min_j (col_bias[j] - 2 * (A.B)[row, j]) and part_idx[row][p] = arg min (ties: smallest j);
 * p < sp_gemm_argmin_parts(N).  k-means assignment: col_bias = |c_j|^2 (k_means_.py:61-66). */
int64_t sp_gemm_argmin_parts(int64_t N);
int sp_gemm_prepared_argmin(int n_seg, const sp_gemm_prepared_segment* segs, int64_t M, int64_t N, const float* col_bias,
                            float* part_val, int32_t* part_idx, int precision, void* stream);
/* The whole k-means assignment in the GEMM (k_means_.py:61-97): labels[i] = argmin_j (col_bias[j] - 2 (A.B)[i, j]),
 * counts[labels[i]] += 1, sums[labels[i], :] += pts[i, :d].  Column tiles of a row tile are walked back to back, the running
 * arg min stays in registers and the epilogue warps accumulate while the tensor cores work on the next row tile.
 * d, ldp multiples of 4; pts, sums 16-byte aligned. */
int sp_gemm_prepared_kmeans(const sp_gemm_prepared_segment* seg, int64_t M, int64_t N, const float* col_bias,
                            const float* pts, int64_t ldp, int64_t d, int32_t* labels, float* sums, int64_t* counts,
                            int precision, void* stream);
int sp_gemm_f32(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int64_t M,
                int64_t N, int64_t K, int accumulate, int precision, void* workspace, int64_t workspace_bytes,
                void* stream);
/* Operands that are transposed VIEWS (transpose.py:27-67; tests/test_transpose.py:27-37 dot(A, transpose(B))):
 * a_trans != 0: A is passed as At [K, M] row-major; b_trans != 0: B is passed as Bt [N, K] row-major.  No transpose is
 * materialised -- the operand preparation reads the view's layout directly. */
int sp_gemm_f32_ex(const float* A, int64_t lda, int a_trans, const float* B, int64_t ldb, int b_trans, float* C,
                   int64_t ldc, int64_t M, int64_t N, int64_t K, int accumulate, int precision, void* workspace,
                   int64_t workspace_bytes, void* stream);
/* ------------------------------------------------------------------------
 * Application kernels on the hot path (BASELINE configs 4, 5).
 * ------------------------------------------------------------------------ */
/* One k-means pass over a block of points (examples/sklearn/cluster/k_means_.py:61-97):
 *   labels[i] = argmin_j |X_i - centers_j|^2   (kmeans_map2_dist_mapper; ties -> smallest j, like np.argmin)
 *   counts[labels[i]] += 1                      (kmeans_count_mapper)
 *   sums[labels[i], :] += X_i                   (kmeans_center_mapper)
 * sums [k, d] and counts [k] are ACCUMULATED into (zero them before the first block); the caller divides and
 * all-reduces across GPUs.  Distances go through the tensor cores (bf16x3 split GEMM of X . centers^T). */
int64_t sp_kmeans_workspace_bytes(int64_t n, int64_t d, int64_t k);
int sp_kmeans_assign(const float* X, int64_t ldx, int64_t n, int64_t d, const float* centers, int64_t k,
                     int32_t* labels, float* sums, int64_t* counts, void* workspace, int64_t workspace_bytes,
                     void* stream);
/* The same pass split in two, because the points of a run never change (k_means_.py:130-160 iterates over the same X):
 * sp_kmeans_prepare_points lays X out once as the tensor cores consume it ([2][n][Kp] bf16 hi/lo,
 * sp_kmeans_prepared_bytes bytes); sp_kmeans_assign_prepared is one iteration over it. */
int64_t sp_kmeans_prepared_bytes(int64_t n, int64_t d);
int sp_kmeans_prepare_points(const float* X, int64_t ldx, int64_t n, int64_t d, void* out, int64_t out_bytes, void* stream);
int64_t sp_kmeans_assign_workspace_bytes(int64_t n, int64_t d, int64_t k);
int sp_kmeans_set_fused(int on); /* test hook: 1 = everything in the GEMM epilogue (default), 0 = candidates + second kernel */
int sp_kmeans_assign_prepared(const void* Xprep, const float* X, int64_t ldx, int64_t n, int64_t d, const float* centers,
                              int64_t k, int32_t* labels, float* sums, int64_t* counts, void* workspace,
                              int64_t workspace_bytes, void* stream);
/* y (+)= A x with A in CSR (dot.py:213-217 `tocsr().dot(dense)`; sparse.pyx:103-158 dot_coo_dense_unordered_map).
 * rowptr[n_rows + 1] int32 (rowptr_is_i64 == 0; use it whenever nnz < 2^31) or int64, colidx int32, values fp32; x, y
 * dense fp32.  avg_nnz_per_row (0 = unknown) picks the number of threads that share a row. */
int sp_spmv_csr(const void* rowptr, int rowptr_is_i64, const int32_t* colidx, const float* values, int64_t n_rows,
                const float* x, float* y, int accumulate, int avg_nnz_per_row, void* stream);

/* CUDA-core GEMM for dtypes the tensor path does not carry exactly (reference tests use
 * float64 / int64 operands: tests/test_dot.py:8-103, tests/test_matmul.py:12-22). */
int sp_gemm_simt(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, int64_t M, int64_t N,
                 int64_t K, int dtype, int accumulate, void* stream);

/* ------------------------------------------------------------------------
 * Peer memory: the cross-rank tile movement of the path (one process per GPU).
 * Replaces the strip fetches / partial-tile updates the reference sends over its RPC layer
 * (join_mapper map.py:243-286 -> DistArrayImpl.fetch distarray.py:294-367 -> Worker.get worker.py:188-217).
 * A rank allocates symmetric buffers (sp_peer_alloc: cudaMalloc'ed, zero-filled), exports them (sp_peer_export writes
 * sp_peer_handle_bytes() bytes the host exchanges by any means) and opens its peers' (sp_peer_import / sp_peer_close).
 * sp_peer_push moves `bytes` from a local buffer to n destinations with the copy engines over NVLink and writes, behind
 * each copy in stream order, the 4-byte word at flag_src into flag_dst[i] (may be NULL) -- the word a consumer polls
 * (sp_gemm_prepared_views_gated, sp_wait_u32).  sp_write_u32 / sp_wait_u32 are the stream-side set / wait of such a
 * word (wait gives up after timeout_ms and increments *status, a device uint32 or NULL).
 * ------------------------------------------------------------------------ */
int sp_peer_alloc(int64_t bytes, void** out);
int sp_peer_free(void* p);
int sp_peer_handle_bytes(void);
int sp_peer_export(void* p, void* handle_out);
int sp_peer_import(const void* handle, void** out);
int sp_peer_close(void* p);
int sp_peer_push(int n, void* const* dst, const void* src, int64_t bytes, void* const* flag_dst, const void* flag_src,
                 void* stream);
int sp_peer_push_2d(int n, void* const* dst, int64_t dst_pitch, const void* src, int64_t src_pitch, int64_t width_bytes,
                    int64_t rows, void* const* flag_dst, const void* flag_src, void* stream);
int sp_write_u32(void* p, uint32_t value, void* stream);
int sp_wait_u32(const void* flag, uint32_t value, int64_t timeout_ms, void* status, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPARTAN_B200_H_ */
