// CUDA-core GEMM: exact per-dtype arithmetic for operand types the tensor path
// does not carry (float64 / int64 / int32) and an independent fp32 cross-check
// of the tcgen05 kernel.  Reference call site: `tiles[0].dot(tiles[1])`
// (spartan/expr/dot.py:217,238) on float64/int operands, as pinned by
// tests/test_dot.py:8-103 and tests/test_matmul.py:12-22 (exact equality).
//
// 64x64 output tile per CTA, 16-deep K slices staged in shared memory, 4x4
// register micro-tile per thread; products accumulate in K order with separate
// multiply and add (no FMA contraction, see -fmad=false) so float results are
// the plain IEEE left-to-right sum over k within a tile.
#include "sp_common.h"

namespace sp {

template <typename T>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const T* __restrict__ A, int64_t lda, const T* __restrict__ B, int64_t ldb, T* __restrict__ C,
                 int64_t ldc, int M, int N, int K, int accumulate) {
  constexpr int TM = 64, TN = 64, TK = 16;
  __shared__ T As[TK][TM + 1];
  __shared__ T Bs[TK][TN + 1];
  const int tx = threadIdx.x % 16;  // column group
  const int ty = threadIdx.x / 16;  // row group
  const int m0 = blockIdx.y * TM;
  const int n0 = blockIdx.x * TN;
  T acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = T(0);

  for (int k0 = 0; k0 < K; k0 += TK) {
    // A tile: TM x TK, loaded with k fastest
    for (int i = threadIdx.x; i < TM * TK; i += 256) {
      const int r = i / TK, c = i % TK;
      const int gm = m0 + r, gk = k0 + c;
      As[c][r] = (gm < M && gk < K) ? A[static_cast<int64_t>(gm) * lda + gk] : T(0);
    }
    for (int i = threadIdx.x; i < TK * TN; i += 256) {
      const int r = i / TN, c = i % TN;
      const int gk = k0 + r, gn = n0 + c;
      Bs[r][c] = (gk < K && gn < N) ? B[static_cast<int64_t>(gk) * ldb + gn] : T(0);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      T a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = acc[i][j] + a[i] * b[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      T* dst = C + static_cast<int64_t>(gm) * ldc + gn;
      *dst = accumulate ? (*dst + acc[i][j]) : acc[i][j];
    }
  }
}

template <typename T>
static int launch_simt(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, int64_t M,
                       int64_t N, int64_t K, int accumulate, cudaStream_t stream) {
  dim3 grid(static_cast<unsigned>((N + 63) / 64), static_cast<unsigned>((M + 63) / 64));
  gemm_simt_kernel<T><<<grid, 256, 0, stream>>>(static_cast<const T*>(A), lda, static_cast<const T*>(B), ldb,
                                               static_cast<T*>(C), ldc, static_cast<int>(M), static_cast<int>(N),
                                               static_cast<int>(K), accumulate);
  SP_CUDA_CHECK(cudaGetLastError());
  return SP_OK;
}

}  // namespace sp

extern "C" int sp_gemm_simt(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, int64_t M,
                             int64_t N, int64_t K, int dtype, int accumulate, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SP_REQUIRE(M > 0 && N > 0 && K >= 0 && M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), SP_ERR_INVALID,
             "sp_gemm_simt: bad shape %lld %lld %lld", (long long)M, (long long)N, (long long)K);
  SP_REQUIRE((N + 63) / 64 <= 2147483647ll && (M + 63) / 64 <= 65535, SP_ERR_INVALID,
             "sp_gemm_simt: M=%lld too large for one launch", (long long)M);
  switch (dtype) {
    case SP_F32: return sp::launch_simt<float>(A, lda, B, ldb, C, ldc, M, N, K, accumulate, stream);
    case SP_F64: return sp::launch_simt<double>(A, lda, B, ldb, C, ldc, M, N, K, accumulate, stream);
    case SP_I64: return sp::launch_simt<long long>(A, lda, B, ldb, C, ldc, M, N, K, accumulate, stream);
    case SP_I32: return sp::launch_simt<int>(A, lda, B, ldb, C, ldc, M, N, K, accumulate, stream);
    default:
      sp::set_error("sp_gemm_simt: unsupported dtype %d", dtype);
      return SP_ERR_UNSUPPORTED;
  }
}
