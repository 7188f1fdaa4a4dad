// k-means assignment/accumulation and CSR SpMV: the two application kernels on the hot path
// (BASELINE configs 4 and 5).
//
// k-means (reference: spartan/examples/sklearn/cluster/k_means_.py:61-97)
//   kmeans_map2_dist_mapper   labels = argmin(cdist(points, centers), axis=1)      O(n k d) scalar C in SciPy
//   kmeans_count_mapper       counts = bincount(labels)
//   kmeans_center_mapper      sums[c] = points[labels == c].sum(0)                 Python loop over k
// Here: distances through the tensor cores -- argmin_j |x - c_j|^2 = argmin_j (|c_j|^2 - 2 x.c_j): one tcgen05
// GEMM per chunk of rows (bf16x3, fp32-faithful) whose EPILOGUE reduces every 128-column half tile to its
// (min, arg min), so the n x k distance matrix never exists in HBM -- then one warp per point picks the label
// among the candidates (ties: smallest index, like np.argmin) and adds the point into its centroid with
// vectorised fire-and-forget float atomics.  Roofline: distance GEMM = tensor pipe (2 n d k flop);
// label + accumulate = HBM (n*d point bytes read twice + atomics).
//
// SpMV (reference: spartan/expr/dot.py:213-217 `tocsr().dot(dense)`, spartan/array/sparse.pyx:103-158)
//   y (+)= A x, A in CSR: one 8-thread group per row, coalesced val/col loads, gather of x through
//   the read-only path, shuffle reduction.  Roofline: HBM, 8 B per non-zero + 12 B per row.
#include "sp_common.h"
#include <algorithm>
#include <float.h>
#include <stdlib.h>

namespace sp {

__global__ void row_sqnorm_kernel(const float* __restrict__ C, int64_t ldc, int k, int d, float* __restrict__ out) {
  const int row = blockIdx.x * (blockDim.x / 32) + (threadIdx.x / 32);
  const int lane = threadIdx.x & 31;
  if (row >= k) return;
  float s = 0.f;
  for (int j = lane; j < d; j += 32) {
    const float v = C[static_cast<int64_t>(row) * ldc + j];
    s += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (lane == 0) out[row] = s;
}

// Candidates of one point -> its label.  parts == 8 (k = 1024): every lane reads the 8 (value, index) pairs itself
// (uniform 16-byte loads, one transaction per warp) and picks the minimum in registers -- no shuffles; otherwise the lanes
// split the candidates and reduce by shuffle.  Ties resolve to the smaller index, like np.argmin.
__device__ __forceinline__ int kmeans_pick(const float* __restrict__ part_val, const int* __restrict__ part_idx, int parts,
                                           int64_t i, int lane) {
  float best = FLT_MAX;
  int best_j = 0x7fffffff;
  if (parts == 8) {
    const float4 v0 = __ldcs(reinterpret_cast<const float4*>(part_val + i * 8));
    const float4 v1 = __ldcs(reinterpret_cast<const float4*>(part_val + i * 8) + 1);
    const int4 j0 = __ldcs(reinterpret_cast<const int4*>(part_idx + i * 8));
    const int4 j1 = __ldcs(reinterpret_cast<const int4*>(part_idx + i * 8) + 1);
    const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    const int j[8] = {j0.x, j0.y, j0.z, j0.w, j1.x, j1.y, j1.z, j1.w};
#pragma unroll
    for (int p = 0; p < 8; ++p)
      if (v[p] < best || (v[p] == best && j[p] < best_j)) { best = v[p]; best_j = j[p]; }
    return best_j;
  }
  for (int p = lane; p < parts; p += 32) {
    const float v = __ldcs(part_val + i * parts + p);
    const int c = __ldcs(part_idx + i * parts + p);
    if (v < best || (v == best && c < best_j)) { best = v; best_j = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oj = __shfl_xor_sync(0xffffffffu, best_j, o);
    if (ob < best || (ob == best && oj < best_j)) { best = ob; best_j = oj; }
  }
  return best_j;
}

// one warp per point, two points in flight per warp: the rows of both points are loaded before either label is known
// (their addresses do not depend on it), then both labels are picked, then both rows go into their centroids with
// vectorised fire-and-forget float atomics (red.global.add.v4.f32).
template <bool VEC>
__global__ void __launch_bounds__(256)
kmeans_label_accumulate_kernel(const float* __restrict__ part_val, const int* __restrict__ part_idx, int parts,
                               const float* __restrict__ X, int64_t ldx, int64_t n, int d,
                               int32_t* __restrict__ labels, float* __restrict__ sums, unsigned long long* __restrict__ counts) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = static_cast<int64_t>(gridDim.x) * (blockDim.x / 32);
  const int64_t w = blockIdx.x * static_cast<int64_t>(blockDim.x / 32) + (threadIdx.x / 32);
  for (int64_t i = w; i < n; i += 2 * warps) {
    const int64_t i2 = i + warps;
    const bool two = i2 < n;
    float4 xa[2], xb[2];
    if (VEC) {                          // d == 256: two float4 per lane cover a row
      const float4* pa = reinterpret_cast<const float4*>(X + i * ldx);
      xa[0] = __ldcs(pa + lane); xa[1] = __ldcs(pa + 32 + lane);
      if (two) {
        const float4* pb = reinterpret_cast<const float4*>(X + i2 * ldx);
        xb[0] = __ldcs(pb + lane); xb[1] = __ldcs(pb + 32 + lane);
      }
    }
    const int la = kmeans_pick(part_val, part_idx, parts, i, lane);
    const int lb = two ? kmeans_pick(part_val, part_idx, parts, i2, lane) : 0;
    if (lane == 0) {
      labels[i] = la;
      atomicAdd(counts + la, 1ull);
      if (two) {
        labels[i2] = lb;
        atomicAdd(counts + lb, 1ull);
      }
    }
    if (VEC) {
      float4* da = reinterpret_cast<float4*>(sums + static_cast<int64_t>(la) * d);
      atomicAdd(da + lane, xa[0]); atomicAdd(da + 32 + lane, xa[1]);
      if (two) {
        float4* db = reinterpret_cast<float4*>(sums + static_cast<int64_t>(lb) * d);
        atomicAdd(db + lane, xb[0]); atomicAdd(db + 32 + lane, xb[1]);
      }
    } else {
      for (int pt = 0; pt < (two ? 2 : 1); ++pt) {
        const float* x = X + (pt ? i2 : i) * ldx;
        float* dst = sums + static_cast<int64_t>(pt ? lb : la) * d;
        if ((d & 3) == 0 && ((reinterpret_cast<uint64_t>(x) | reinterpret_cast<uint64_t>(dst)) & 15) == 0) {
          for (int j = lane * 4; j < d; j += 128) atomicAdd(reinterpret_cast<float4*>(dst + j), *reinterpret_cast<const float4*>(x + j));
        } else {
          for (int j = lane; j < d; j += 32) atomicAdd(dst + j, x[j]);
        }
      }
    }
  }
}

// CSR SpMV: GROUP threads per row; PTR = int32_t row pointers whenever nnz < 2^31 (4 B per row instead of 8)
template <int GROUP, typename PTR>
__global__ void __launch_bounds__(256)
spmv_csr_kernel(const PTR* __restrict__ rowptr, const int32_t* __restrict__ col, const float* __restrict__ val,
                int64_t n_rows, const float* __restrict__ x, float* __restrict__ y, int accumulate) {
  const int g = threadIdx.x % GROUP;
  const int64_t groups = static_cast<int64_t>(gridDim.x) * (blockDim.x / GROUP);
  for (int64_t row = blockIdx.x * static_cast<int64_t>(blockDim.x / GROUP) + threadIdx.x / GROUP; row < n_rows;
       row += groups) {
    const int64_t lo = rowptr[row], hi = rowptr[row + 1];
    float s = 0.f;
    for (int64_t p = lo + g; p < hi; p += GROUP) s += __ldcs(val + p) * __ldg(x + __ldcs(col + p));
#pragma unroll
    for (int o = GROUP / 2; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o, GROUP);
    if (g == 0) y[row] = accumulate ? y[row] + s : s;
  }
}

// CSR SpMV, non-zero-parallel: a warp owns 32 consecutive rows and walks THEIR non-zeros 32 at a time, so val / col are
// read with perfectly coalesced 128-byte accesses whatever the row lengths (the row-parallel kernel above touches every
// 32-byte sector of val / col about twice: once per 8-lane row group and again for rows longer than the group).  Each lane
// finds the row of its non-zero by a shuffle binary search over the warp's 33 row pointers, the products are combined by
// a segmented shuffle scan, and the last lane of every row segment adds its sum to the row's slot in shared memory.  What
// remains is the gather of x[col] -- one sector per non-zero -- which bounds SpMV on random columns (L1 tag stage).
template <typename PTR>
__global__ void __launch_bounds__(256)
spmv_csr_stream_kernel(const PTR* __restrict__ rowptr, const int32_t* __restrict__ col, const float* __restrict__ val,
                       int64_t n_rows, const float* __restrict__ x, float* __restrict__ y, int accumulate) {
  __shared__ float s_acc[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t n_blocks = (n_rows + 31) / 32;
  for (int64_t rb = blockIdx.x * 8ll + w; rb < n_blocks; rb += gridDim.x * 8ll) {
    const int64_t row0 = rb * 32;
    const int64_t my_row = row0 + lane;
    const long long rp = static_cast<long long>(rowptr[min(my_row, n_rows)]);       // lanes past the end: total nnz
    const long long lo = __shfl_sync(0xffffffffu, rp, 0);
    const long long hi = static_cast<long long>(rowptr[min(row0 + 32, n_rows)]);
    s_acc[w][lane] = 0.f;
    __syncwarp();
    for (long long p = lo; p < hi; p += 32) {
      const long long e = p + lane;
      const bool live = e < hi;
      float prod = 0.f;
      if (live) prod = __ldcs(val + e) * __ldg(x + __ldcs(col + e));
      int r = 0;                                   // largest r with rowptr[row0 + r] <= e
#pragma unroll
      for (int step = 16; step > 0; step >>= 1) {
        const int cand = r + step;
        const long long v = __shfl_sync(0xffffffffu, rp, cand & 31);
        if (cand < 32 && v <= e) r = cand;
      }
      if (!live) r = 32 + lane;                    // dead lanes: unique keys, never merged
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {           // segmented inclusive scan over equal rows
        const float up = __shfl_up_sync(0xffffffffu, prod, d);
        const int rup = __shfl_up_sync(0xffffffffu, r, d);
        if (lane >= d && rup == r) prod += up;
      }
      const int rnext = __shfl_down_sync(0xffffffffu, r, 1);
      if (live && (lane == 31 || rnext != r)) s_acc[w][r] += prod;       // segment tails: distinct rows within a batch
      __syncwarp();
    }
    if (my_row < n_rows) y[my_row] = accumulate ? y[my_row] + s_acc[w][lane] : s_acc[w][lane];
    __syncwarp();
  }
}

template <typename PTR>
static void launch_spmv(int group, int blocks, const PTR* rowptr, const int32_t* colidx, const float* values,
                        int64_t n_rows, const float* x, float* y, int accumulate, cudaStream_t stream) {
  switch (group) {
    case 2: spmv_csr_kernel<2, PTR><<<blocks, 256, 0, stream>>>(rowptr, colidx, values, n_rows, x, y, accumulate); break;
    case 4: spmv_csr_kernel<4, PTR><<<blocks, 256, 0, stream>>>(rowptr, colidx, values, n_rows, x, y, accumulate); break;
    case 8: spmv_csr_kernel<8, PTR><<<blocks, 256, 0, stream>>>(rowptr, colidx, values, n_rows, x, y, accumulate); break;
    default: spmv_csr_kernel<32, PTR><<<blocks, 256, 0, stream>>>(rowptr, colidx, values, n_rows, x, y, accumulate); break;
  }
}

static int g_spmv_stream = -1;      // -1: from the environment (SPARTAN_SPMV_KERNEL=rows|stream, default stream)
static bool spmv_stream() {
  if (g_spmv_stream < 0) {
    const char* e = getenv("SPARTAN_SPMV_KERNEL");
    g_spmv_stream = (e && e[0] == 'r') ? 0 : 1;
  }
  return g_spmv_stream == 1;
}
static int g_kmeans_fused = -1;     // -1: from the environment (SPARTAN_KMEANS_FUSED, default 1)
static bool kmeans_fused() {
  if (g_kmeans_fused < 0) {
    const char* e = getenv("SPARTAN_KMEANS_FUSED");
    g_kmeans_fused = (e && e[0] == '0') ? 0 : 1;
  }
  return g_kmeans_fused == 1;
}

}  // namespace sp

using namespace sp;

extern "C" int64_t sp_kmeans_workspace_bytes(int64_t n, int64_t d, int64_t k) {
  const int64_t Kp = sp_gemm_kpad(d, SP_GEMM_BF16X3);
  return sp_gemm_prepared_bytes(n, Kp, SP_GEMM_BF16X3) + sp_kmeans_assign_workspace_bytes(n, d, k) + 2048;
}

// Workspace of sp_kmeans_assign_prepared: prepared centres + per-row argmin candidates + centre norms.
extern "C" int64_t sp_kmeans_assign_workspace_bytes(int64_t n, int64_t d, int64_t k) {
  const int64_t Kp = sp_gemm_kpad(d, SP_GEMM_BF16X3);
  return sp_gemm_prepared_bytes(k, Kp, SP_GEMM_BF16X3) + n * sp_gemm_argmin_parts(k) * 8 + k * 4 + 8192;
}

// The points as the tensor cores consume them: [2][n][Kp] bf16 (hi, lo), Kp = d rounded up to the k-block.  The points
// of a k-means run never change, so this is done once per fit, not once per iteration.
extern "C" int64_t sp_kmeans_prepared_bytes(int64_t n, int64_t d) {
  return sp_gemm_prepared_bytes(n, sp_gemm_kpad(d, SP_GEMM_BF16X3), SP_GEMM_BF16X3);
}

extern "C" int sp_kmeans_prepare_points(const float* X, int64_t ldx, int64_t n, int64_t d, void* out, int64_t out_bytes,
                                        void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SP_REQUIRE(n >= 0 && d > 0 && d < (1ll << 31), SP_ERR_INVALID, "sp_kmeans_prepare_points: bad shape");
  if (n == 0) return SP_OK;
  const int64_t Kp = sp_gemm_kpad(d, SP_GEMM_BF16X3);
  if (Kp != (d + 3) / 4 * 4) SP_CUDA_CHECK(cudaMemsetAsync(out, 0, sp_kmeans_prepared_bytes(n, d), stream));
  return sp_gemm_prepare_a(X, ldx, n, d, SP_GEMM_BF16X3, out, Kp, 0, out_bytes, stream);
}

// One assignment pass over points already prepared by sp_kmeans_prepare_points (Xprep) -- X itself is still read by the
// accumulation (fp32 sums of the original values).
extern "C" int sp_kmeans_assign_prepared(const void* Xprep, const float* X, int64_t ldx, int64_t n, int64_t d,
                                         const float* centers, int64_t k, int32_t* labels, float* sums, int64_t* counts,
                                         void* workspace, int64_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SP_REQUIRE(n >= 0 && d > 0 && k > 0 && k < (1ll << 31) && d < (1ll << 31) && n < (1ll << 31), SP_ERR_INVALID,
             "sp_kmeans_assign_prepared: bad shape");
  if (n == 0) return SP_OK;
  SP_REQUIRE(Xprep && X && centers && labels && sums && counts, SP_ERR_INVALID, "sp_kmeans_assign_prepared: null pointer");
  SP_REQUIRE(workspace != nullptr && workspace_bytes >= sp_kmeans_assign_workspace_bytes(n, d, k), SP_ERR_INVALID,
             "sp_kmeans_assign_prepared: workspace too small");
  const int prec = SP_GEMM_BF16X3;
  const int64_t Kp = sp_gemm_kpad(d, prec);
  uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uint64_t>(workspace) + 1023) & ~1023ull);
  const int64_t b_bytes = (sp_gemm_prepared_bytes(k, Kp, prec) + 1023) / 1024 * 1024;
  uint8_t* b_prep = ws;
  const int parts = static_cast<int>(sp_gemm_argmin_parts(k));
  float* part_val = reinterpret_cast<float*>(b_prep + b_bytes);
  int32_t* part_idx = reinterpret_cast<int32_t*>(part_val + n * parts);
  float* cnorm = reinterpret_cast<float*>((reinterpret_cast<uint64_t>(part_idx + n * parts) + 15) & ~15ull);   // 16-byte loads
  if (Kp != (d + 3) / 4 * 4) SP_CUDA_CHECK(cudaMemsetAsync(b_prep, 0, b_bytes, stream));
  // the centres are already "B transposed" ([k, d] = [N, K]): prepare them like an A operand
  int rc = sp_gemm_prepare_a(centers, d, k, d, prec, b_prep, Kp, 0, b_bytes, stream);
  if (rc) return rc;
  row_sqnorm_kernel<<<static_cast<unsigned>((k + 7) / 8), 256, 0, stream>>>(centers, d, static_cast<int>(k),
                                                                            static_cast<int>(d), cnorm);
  sp_gemm_prepared_segment seg;
  seg.A = Xprep; seg.B = b_prep; seg.Kp = Kp;
  const bool fused = kmeans_fused() && (d & 3) == 0 && (ldx & 3) == 0 &&
                     ((reinterpret_cast<uint64_t>(X) | reinterpret_cast<uint64_t>(sums)) & 15) == 0;
  if (fused)      // labels, counts and sums straight from the GEMM's epilogue: no candidates in HBM, no second pass
    return sp_gemm_prepared_kmeans(&seg, n, k, cnorm, X, ldx, d, labels, sums, counts, prec, stream);
  rc = sp_gemm_prepared_argmin(1, &seg, n, k, cnorm, part_val, part_idx, prec, stream);   // no n x k matrix in HBM
  if (rc) return rc;
  const int blocks = static_cast<int>(std::min<int64_t>((n + 7) / 8, static_cast<int64_t>(num_sms()) * 8));
  const bool vec = d == 256 && (ldx & 3) == 0 && ((reinterpret_cast<uint64_t>(X) | reinterpret_cast<uint64_t>(sums)) & 15) == 0;
  if (vec)
    kmeans_label_accumulate_kernel<true><<<blocks, 256, 0, stream>>>(part_val, part_idx, parts, X, ldx, n, static_cast<int>(d),
                                                                     labels, sums, reinterpret_cast<unsigned long long*>(counts));
  else
    kmeans_label_accumulate_kernel<false><<<blocks, 256, 0, stream>>>(part_val, part_idx, parts, X, ldx, n, static_cast<int>(d),
                                                                      labels, sums, reinterpret_cast<unsigned long long*>(counts));
  SP_CUDA_CHECK(cudaGetLastError());
  return SP_OK;
}

// Test / tuning hook: 1 (default) = labels / counts / sums from the GEMM's own epilogue, 0 = candidates + second kernel.
extern "C" int sp_kmeans_set_fused(int on) {
  g_kmeans_fused = on ? 1 : 0;
  return SP_OK;
}

// Convenience: prepare + assign in one call (the workspace also holds the prepared points).
extern "C" int sp_kmeans_assign(const float* X, int64_t ldx, int64_t n, int64_t d, const float* centers, int64_t k,
                                int32_t* labels, float* sums, int64_t* counts, void* workspace,
                                int64_t workspace_bytes, void* stream_) {
  SP_REQUIRE(n >= 0 && d > 0 && k > 0, SP_ERR_INVALID, "sp_kmeans_assign: bad shape");
  if (n == 0) return SP_OK;
  SP_REQUIRE(workspace != nullptr && workspace_bytes >= sp_kmeans_workspace_bytes(n, d, k), SP_ERR_INVALID,
             "sp_kmeans_assign: workspace too small");
  uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uint64_t>(workspace) + 1023) & ~1023ull);
  const int64_t a_bytes = (sp_kmeans_prepared_bytes(n, d) + 1023) / 1024 * 1024;
  int rc = sp_kmeans_prepare_points(X, ldx, n, d, ws, a_bytes, stream_);
  if (rc) return rc;
  return sp_kmeans_assign_prepared(ws, X, ldx, n, d, centers, k, labels, sums, counts, ws + a_bytes,
                                   workspace_bytes - a_bytes - 1024, stream_);
}

extern "C" int sp_spmv_csr(const void* rowptr, int rowptr_is_i64, const int32_t* colidx, const float* values,
                           int64_t n_rows, const float* x, float* y, int accumulate, int avg_nnz_per_row, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SP_REQUIRE(n_rows >= 0, SP_ERR_INVALID, "sp_spmv_csr: negative row count");
  if (n_rows == 0) return SP_OK;
  SP_REQUIRE(rowptr && x && y, SP_ERR_INVALID, "sp_spmv_csr: null pointer");
  // threads per row follow the average row length (passed by the caller; <= 0 means unknown -> 8)
  if (avg_nnz_per_row >= 4 && spmv_stream()) {     // enough non-zeros per 32 rows to fill the 32-wide batches
    const int64_t n_blocks = (n_rows + 31) / 32;
    const int blocks = static_cast<int>(std::min<int64_t>((n_blocks + 7) / 8, static_cast<int64_t>(num_sms()) * 32));
    if (rowptr_is_i64)
      spmv_csr_stream_kernel<int64_t><<<blocks, 256, 0, stream>>>(static_cast<const int64_t*>(rowptr), colidx, values, n_rows, x, y, accumulate);
    else
      spmv_csr_stream_kernel<int32_t><<<blocks, 256, 0, stream>>>(static_cast<const int32_t*>(rowptr), colidx, values, n_rows, x, y, accumulate);
    SP_CUDA_CHECK(cudaGetLastError());
    return SP_OK;
  }
  const int group = avg_nnz_per_row <= 0 ? 8 : avg_nnz_per_row <= 2 ? 2 : avg_nnz_per_row <= 4 ? 4 : avg_nnz_per_row <= 16 ? 8 : 32;
  const int64_t groups_per_block = 256 / group;
  const int blocks = static_cast<int>(std::min<int64_t>((n_rows + groups_per_block - 1) / groups_per_block,
                                                        static_cast<int64_t>(num_sms()) * 16));
  if (rowptr_is_i64)
    launch_spmv(group, blocks, static_cast<const int64_t*>(rowptr), colidx, values, n_rows, x, y, accumulate, stream);
  else
    launch_spmv(group, blocks, static_cast<const int32_t*>(rowptr), colidx, values, n_rows, x, y, accumulate, stream);
  SP_CUDA_CHECK(cudaGetLastError());
  return SP_OK;
}
