// Tile-level utility kernels: fill / iota / counter-based RNG, the same-device combiner and
// strided rectangle copies.  All HBM-bound, one pass over the data.
//
// Reference call sites:
//   fill      creation.py:67-68 np.zeros, :92-93 np.ones, :135-141 np.arange per tile,
//             srandom.py:40-47 np.random.rand / randn per tile
//   combine   Tile.merge dense path  reducer(old, update)          tile.pyx:263-283
//   copy_rect DistArrayImpl.fetch stitching / update splitting     distarray.py:294-422
#include "sp_common.h"

namespace sp {

// ---------------------------------------------------------------------------- Philox4x32-10
struct Philox {
  static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  __host__ __device__ static inline void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint64_t p0 = static_cast<uint64_t>(M0) * c[0];
    const uint64_t p1 = static_cast<uint64_t>(M1) * c[2];
    const uint32_t hi0 = static_cast<uint32_t>(p0 >> 32), lo0 = static_cast<uint32_t>(p0);
    const uint32_t hi1 = static_cast<uint32_t>(p1 >> 32), lo1 = static_cast<uint32_t>(p1);
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
  // counter = 64-bit block index, key = 64-bit seed -> 4 x 32 random bits
  __host__ __device__ static inline void block(uint64_t counter, uint64_t seed, uint32_t (&out)[4]) {
    uint32_t c[4] = {static_cast<uint32_t>(counter), static_cast<uint32_t>(counter >> 32), 0u, 0u};
    uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
    for (int r = 0; r < 10; ++r) {      // constant trip count: unrolled by nvcc without a pragma (gcc warns on it)
      round(c, k0, k1);
      k0 += W0; k1 += W1;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
  }
};

// element i of the stream: uniform in [0,1) with 24 (f32) / 53 (f64) random bits
__device__ __forceinline__ float u01_f32(uint32_t x) { return (x >> 8) * (1.0f / 16777216.0f); }
__device__ __forceinline__ double u01_f64(uint32_t hi, uint32_t lo) {
  const uint64_t v = ((static_cast<uint64_t>(hi) << 32) | lo) >> 11;
  return v * (1.0 / 9007199254740992.0);
}

template <typename T>
__global__ void fill_const_kernel(T* __restrict__ dst, int64_t n, T value) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    dst[i] = value;
}

template <typename T>
__global__ void fill_iota_kernel(T* __restrict__ dst, int64_t n, double a, double b, int64_t offset, int integral) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    if (integral) {
      dst[i] = static_cast<T>(static_cast<long long>(a) + static_cast<long long>(b) * (offset + i));
    } else {
      dst[i] = static_cast<T>(a + b * static_cast<double>(offset + i));
    }
  }
}

// Each Philox block yields 4 f32 (or 2 f64) values; element e uses block e/4 (e/2), lane e%4 (e%2),
// so any sub-range of the stream can be regenerated independently (tiles of one array share a seed
// and use their global element offset as `offset`).
template <typename T>
__global__ void fill_rand_kernel(T* __restrict__ dst, int64_t n, uint64_t seed, int64_t offset, int normal) {
  constexpr int PER = (sizeof(T) == 4) ? 4 : 2;
  const int64_t first_blk = offset / PER;
  const int64_t last_blk = (offset + n - 1) / PER;
  const int64_t nblk = last_blk - first_blk + 1;
  for (int64_t bi = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; bi < nblk;
       bi += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t blk = first_blk + bi;
    uint32_t r[4];
    Philox::block(static_cast<uint64_t>(blk), seed, r);
    T vals[PER];
    if constexpr (sizeof(T) == 4) {
      if (normal) {   // Box-Muller on pairs
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          const float u1 = 1.0f - u01_f32(r[2 * p]);           // (0,1]
          const float u2 = u01_f32(r[2 * p + 1]);
          const float rad = sqrtf(-2.0f * logf(u1));
          float s, c;
          sincospif(2.0f * u2, &s, &c);
          vals[2 * p] = rad * c;
          vals[2 * p + 1] = rad * s;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) vals[j] = u01_f32(r[j]);
      }
    } else {
      const double a = u01_f64(r[0], r[1]), b = u01_f64(r[2], r[3]);
      if (normal) {
        const double rad = sqrt(-2.0 * log(1.0 - a));
        double s, c;
        sincospi(2.0 * b, &s, &c);
        vals[0] = rad * c; vals[1] = rad * s;
      } else {
        vals[0] = a; vals[1] = b;
      }
    }
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int64_t e = blk * PER + j - offset;
      if (e >= 0 && e < n) dst[e] = vals[j];
    }
  }
}

// 2-D (strided) fill: element (r, c) takes stream index offset + r * pitch + c, so a tile that is a
// sub-rectangle of its array reproduces exactly the values of the array-wide stream.
template <typename T>
__global__ void fill2d_kernel(T* __restrict__ dst, int64_t rows, int64_t cols, int64_t dst_stride, int kind, double a,
                              double b, uint64_t seed, int64_t offset, int64_t pitch, int integral) {
  constexpr int PER = (sizeof(T) == 4) ? 4 : 2;
  const int64_t total = rows * cols;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / cols, c = i - r * cols;
    const int64_t idx = offset + r * pitch + c;
    T v;
    if (kind == SP_FILL_CONST) {
      v = static_cast<T>(a);
    } else if (kind == SP_FILL_IOTA) {
      v = integral ? static_cast<T>(static_cast<long long>(a) + static_cast<long long>(b) * idx)
                   : static_cast<T>(a + b * static_cast<double>(idx));
    } else {
      uint32_t rr[4];
      Philox::block(static_cast<uint64_t>(idx / PER), seed, rr);
      const int lane = static_cast<int>(idx % PER);
      if constexpr (sizeof(T) == 4) {
        if (kind == SP_FILL_RANDN) {
          const int p = lane >> 1;
          const float u1 = 1.0f - u01_f32(rr[2 * p]);
          const float u2 = u01_f32(rr[2 * p + 1]);
          const float rad = sqrtf(-2.0f * logf(u1));
          float sn, cs;
          sincospif(2.0f * u2, &sn, &cs);
          v = static_cast<T>((lane & 1) ? rad * sn : rad * cs);
        } else {
          v = static_cast<T>(u01_f32(rr[lane]));
        }
      } else {
        const double x = u01_f64(rr[0], rr[1]), y = u01_f64(rr[2], rr[3]);
        if (kind == SP_FILL_RANDN) {
          const double rad = sqrt(-2.0 * log(1.0 - x));
          double sn, cs;
          sincospi(2.0 * y, &sn, &cs);
          v = static_cast<T>(lane ? rad * sn : rad * cs);
        } else {
          v = static_cast<T>(lane ? y : x);
        }
      }
    }
    dst[r * dst_stride + c] = v;
  }
}

template <typename T>
__device__ __forceinline__ T combine_op(int op, T a, T b) {
  switch (op) {
    case SP_RED_SUM: return a + b;
    case SP_RED_PROD: return a * b;
    case SP_RED_MIN: return (a < b || a != a) ? a : b;
    case SP_RED_MAX: return (a > b || a != a) ? a : b;
    case SP_RED_ALL: return static_cast<T>((a != T(0)) && (b != T(0)));
    default: return static_cast<T>((a != T(0)) || (b != T(0)));
  }
}

template <typename T>
__global__ void combine_kernel(T* __restrict__ dst, const T* __restrict__ src, int64_t n, int op) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    dst[i] = combine_op<T>(op, dst[i], src[i]);
}

template <typename T>
__global__ void copy_rect_kernel(T* __restrict__ dst, int64_t ds0, int64_t ds1, int64_t ds2, const T* __restrict__ src,
                                 int64_t ss0, int64_t ss1, int64_t ss2, int64_t d0, int64_t d1, int64_t d2) {
  const int64_t total = d0 * d1 * d2;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t i2 = i % d2;
    const int64_t t = i / d2;
    const int64_t i1 = t % d1;
    const int64_t i0 = t / d1;
    dst[i0 * ds0 + i1 * ds1 + i2 * ds2] = src[i0 * ss0 + i1 * ss1 + i2 * ss2];
  }
}

// Tile.merge on a partially written tile (tile.pyx:270-283): element by element, a first write replaces and a later
// write reduces -- decided by the tile's mask, which is updated in the same pass.  The reduction is evaluated in CT, the
// NumPy result type of (DT, ST), and stored back as DT, like `old_region[updated] = reducer(old_region[updated], update)`.
template <typename DT, typename ST, typename CT>
__global__ void merge_masked_kernel(DT* __restrict__ dst, int64_t ds0, int64_t ds1, int64_t ds2, const ST* __restrict__ src,
                                    int64_t ss0, int64_t ss1, int64_t ss2, uint8_t* __restrict__ mask, int64_t ms0,
                                    int64_t ms1, int64_t ms2, int64_t d0, int64_t d1, int64_t d2, int op) {
  const int64_t total = d0 * d1 * d2;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t i2 = i % d2;
    const int64_t t = i / d2;
    const int64_t i1 = t % d1;
    const int64_t i0 = t / d1;
    const int64_t od = i0 * ds0 + i1 * ds1 + i2 * ds2;
    const int64_t om = i0 * ms0 + i1 * ms1 + i2 * ms2;
    const CT nv = static_cast<CT>(src[i0 * ss0 + i1 * ss1 + i2 * ss2]);
    const bool written = mask[om] != 0;
    dst[od] = (written && op >= 0) ? static_cast<DT>(combine_op<CT>(op, static_cast<CT>(dst[od]), nv)) : static_cast<DT>(nv);
    mask[om] = 1;
  }
}

// dst[i][j] = src[j][i]: a transposed view (spartan/expr/operator/transpose.py:27-67) made dense through a 32 x 32
// shared-memory tile, so that both the reads of `src` and the writes of `dst` are coalesced.
template <typename T>
__global__ void transpose_2d_kernel(T* __restrict__ dst, int64_t ldd, const T* __restrict__ src, int64_t lds, int64_t R,
                                    int64_t C) {
  __shared__ T tile[32][33];
  const int64_t i0 = static_cast<int64_t>(blockIdx.x) * 32;     // dst rows = src columns
  const int64_t j0 = static_cast<int64_t>(blockIdx.y) * 32;     // dst columns = src rows
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int64_t j = j0 + r, i = i0 + threadIdx.x;
    if (j < C && i < R) tile[r][threadIdx.x] = src[j * lds + i];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int64_t i = i0 + r, j = j0 + threadIdx.x;
    if (i < R && j < C) dst[i * ldd + j] = tile[threadIdx.x][r];
  }
}

static inline int grid_for(int64_t n) {
  const int64_t want = (n + 255) / 256;
  return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(want, static_cast<int64_t>(num_sms()) * 16)));
}

template <typename T>
static int fill_t(void* dst, int64_t n, int kind, double a, double b, uint64_t seed, int64_t offset, int integral,
                  cudaStream_t stream) {
  T* p = static_cast<T*>(dst);
  switch (kind) {
    case SP_FILL_CONST: fill_const_kernel<T><<<grid_for(n), 256, 0, stream>>>(p, n, static_cast<T>(a)); break;
    case SP_FILL_IOTA: fill_iota_kernel<T><<<grid_for(n), 256, 0, stream>>>(p, n, a, b, offset, integral); break;
    default:
      set_error("sp_fill: kind %d not valid for this dtype", kind);
      return SP_ERR_UNSUPPORTED;
  }
  SP_CUDA_CHECK(cudaGetLastError());
  return SP_OK;
}

}  // namespace sp

using namespace sp;

extern "C" int sp_fill(void* dst, int dtype, int64_t n, int kind, double a, double b, uint64_t seed, int64_t offset,
                       void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SP_REQUIRE(n >= 0, SP_ERR_INVALID, "sp_fill: negative length");
  if (n == 0) return SP_OK;
  SP_REQUIRE(dst != nullptr, SP_ERR_INVALID, "sp_fill: null destination");
  if (kind == SP_FILL_RAND || kind == SP_FILL_RANDN) {
    SP_REQUIRE(offset >= 0, SP_ERR_INVALID, "sp_fill: negative stream offset");
    const int normal = (kind == SP_FILL_RANDN);
    if (dtype == SP_F32) {
      fill_rand_kernel<float><<<grid_for((n + 3) / 4 + 1), 256, 0, stream>>>(static_cast<float*>(dst), n, seed, offset, normal);
    } else if (dtype == SP_F64) {
      fill_rand_kernel<double><<<grid_for((n + 1) / 2 + 1), 256, 0, stream>>>(static_cast<double*>(dst), n, seed, offset, normal);
    } else {
      set_error("sp_fill: random fill needs a float dtype, got %d", dtype);
      return SP_ERR_UNSUPPORTED;
    }
    SP_CUDA_CHECK(cudaGetLastError());
    return SP_OK;
  }
  switch (dtype) {
    case SP_F32: return fill_t<float>(dst, n, kind, a, b, seed, offset, 0, stream);
    case SP_F64: return fill_t<double>(dst, n, kind, a, b, seed, offset, 0, stream);
    case SP_I32: return fill_t<int32_t>(dst, n, kind, a, b, seed, offset, 1, stream);
    case SP_I64: return fill_t<long long>(dst, n, kind, a, b, seed, offset, 1, stream);
    case SP_U8:
    case SP_BOOL: return fill_t<uint8_t>(dst, n, kind, a, b, seed, offset, 1, stream);
    default:
      set_error("sp_fill: bad dtype %d", dtype);
      return SP_ERR_INVALID;
  }
}

extern "C" int sp_fill2d(void* dst, int dtype, int64_t rows, int64_t cols, int64_t dst_row_stride, int kind, double a,
                         double b, uint64_t seed, int64_t offset, int64_t index_row_pitch, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SP_REQUIRE(rows >= 0 && cols >= 0 && offset >= 0, SP_ERR_INVALID, "sp_fill2d: bad shape / offset");
  if (rows * cols == 0) return SP_OK;
  SP_REQUIRE(dst != nullptr, SP_ERR_INVALID, "sp_fill2d: null destination");
  const bool rnd = (kind == SP_FILL_RAND || kind == SP_FILL_RANDN);
  SP_REQUIRE(kind >= SP_FILL_CONST && kind <= SP_FILL_RANDN, SP_ERR_INVALID, "sp_fill2d: bad kind %d", kind);
  const int g = grid_for(rows * cols);
#define SP_FILL2D(T, INTEGRAL)                                                                                      \
  fill2d_kernel<T><<<g, 256, 0, stream>>>(static_cast<T*>(dst), rows, cols, dst_row_stride, kind, a, b, seed, offset, \
                                          index_row_pitch, INTEGRAL)
  switch (dtype) {
    case SP_F32: SP_FILL2D(float, 0); break;
    case SP_F64: SP_FILL2D(double, 0); break;
    case SP_I32: SP_REQUIRE(!rnd, SP_ERR_UNSUPPORTED, "random fill needs a float dtype"); SP_FILL2D(int32_t, 1); break;
    case SP_I64: SP_REQUIRE(!rnd, SP_ERR_UNSUPPORTED, "random fill needs a float dtype"); SP_FILL2D(long long, 1); break;
    default:
      set_error("sp_fill2d: unsupported dtype %d", dtype);
      return SP_ERR_UNSUPPORTED;
  }
#undef SP_FILL2D
  SP_CUDA_CHECK(cudaGetLastError());
  return SP_OK;
}

extern "C" int sp_combine(void* dst, const void* src, int dtype, int64_t n, int reduce_op, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SP_REQUIRE(n >= 0 && reduce_op >= SP_RED_SUM && reduce_op <= SP_RED_ANY, SP_ERR_INVALID, "sp_combine: bad arguments");
  if (n == 0) return SP_OK;
  SP_REQUIRE(dst != nullptr && src != nullptr, SP_ERR_INVALID, "sp_combine: null pointer");
  const int g = grid_for(n);
  switch (dtype) {
    case SP_F32: combine_kernel<float><<<g, 256, 0, stream>>>(static_cast<float*>(dst), static_cast<const float*>(src), n, reduce_op); break;
    case SP_F64: combine_kernel<double><<<g, 256, 0, stream>>>(static_cast<double*>(dst), static_cast<const double*>(src), n, reduce_op); break;
    case SP_I32: combine_kernel<int32_t><<<g, 256, 0, stream>>>(static_cast<int32_t*>(dst), static_cast<const int32_t*>(src), n, reduce_op); break;
    case SP_I64: combine_kernel<long long><<<g, 256, 0, stream>>>(static_cast<long long*>(dst), static_cast<const long long*>(src), n, reduce_op); break;
    case SP_U8:
    case SP_BOOL: combine_kernel<uint8_t><<<g, 256, 0, stream>>>(static_cast<uint8_t*>(dst), static_cast<const uint8_t*>(src), n, reduce_op); break;
    default:
      set_error("sp_combine: bad dtype %d", dtype);
      return SP_ERR_INVALID;
  }
  SP_CUDA_CHECK(cudaGetLastError());
  return SP_OK;
}

extern "C" int sp_copy_rect(void* dst, const int64_t dst_stride[3], const void* src, const int64_t src_stride[3],
                            const int64_t dims[3], int dtype, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SP_REQUIRE(dims[0] >= 0 && dims[1] >= 0 && dims[2] >= 0, SP_ERR_INVALID, "sp_copy_rect: negative dims");
  const int64_t total = dims[0] * dims[1] * dims[2];
  if (total == 0) return SP_OK;
  SP_REQUIRE(dst != nullptr && src != nullptr, SP_ERR_INVALID, "sp_copy_rect: null pointer");
  const int g = grid_for(total);
#define SP_COPY(T)                                                                                                  \
  copy_rect_kernel<T><<<g, 256, 0, stream>>>(static_cast<T*>(dst), dst_stride[0], dst_stride[1], dst_stride[2],     \
                                             static_cast<const T*>(src), src_stride[0], src_stride[1], src_stride[2], \
                                             dims[0], dims[1], dims[2])
  switch (dtype_size(dtype)) {
    case 1: SP_COPY(uint8_t); break;
    case 4: SP_COPY(uint32_t); break;
    case 8: SP_COPY(uint64_t); break;
    default:
      set_error("sp_copy_rect: bad dtype %d", dtype);
      return SP_ERR_INVALID;
  }
#undef SP_COPY
  SP_CUDA_CHECK(cudaGetLastError());
  return SP_OK;
}

// Host <-> device rectangle transfers (BlobCtx.get / update across the host boundary, blob_ctx.py:127-179;
// from_numpy's upload, write_array.py:424-445).  Pitched DMA: a column block of a row-major host array moves
// in one call at full PCIe rate when the host buffer is pinned.
extern "C" int sp_upload_2d(void* dst_device, int64_t dst_pitch_bytes, const void* src_host, int64_t src_pitch_bytes,
                            int64_t width_bytes, int64_t rows, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SP_REQUIRE(width_bytes >= 0 && rows >= 0, SP_ERR_INVALID, "sp_upload_2d: bad extent");
  if (width_bytes == 0 || rows == 0) return SP_OK;
  SP_REQUIRE(dst_device && src_host, SP_ERR_INVALID, "sp_upload_2d: null pointer");
  SP_CUDA_CHECK(cudaMemcpy2DAsync(dst_device, static_cast<size_t>(dst_pitch_bytes), src_host,
                                  static_cast<size_t>(src_pitch_bytes), static_cast<size_t>(width_bytes),
                                  static_cast<size_t>(rows), cudaMemcpyHostToDevice, stream));
  return SP_OK;
}

extern "C" int sp_download_2d(void* dst_host, int64_t dst_pitch_bytes, const void* src_device, int64_t src_pitch_bytes,
                              int64_t width_bytes, int64_t rows, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SP_REQUIRE(width_bytes >= 0 && rows >= 0, SP_ERR_INVALID, "sp_download_2d: bad extent");
  if (width_bytes == 0 || rows == 0) return SP_OK;
  SP_REQUIRE(dst_host && src_device, SP_ERR_INVALID, "sp_download_2d: null pointer");
  SP_CUDA_CHECK(cudaMemcpy2DAsync(dst_host, static_cast<size_t>(dst_pitch_bytes), src_device,
                                  static_cast<size_t>(src_pitch_bytes), static_cast<size_t>(width_bytes),
                                  static_cast<size_t>(rows), cudaMemcpyDeviceToHost, stream));
  return SP_OK;
}


// Tile.merge, partial-region path (tile.pyx:270-283): dst[i] = mask[i] ? reducer(dst[i], src[i]) : src[i]; mask[i] = 1.
// reduce_op < 0: no reducer (every element is replaced).  dst / src / mask are 3-D strided views (element strides).
// Supported (dst, src) dtype pairs: equal dtypes, (f32, f64), (f64, f32), (i64, i32), (i32, i64); the host casts others.
extern "C" int sp_merge_masked(void* dst, const int64_t dst_stride[3], int dst_dtype, const void* src,
                               const int64_t src_stride[3], int src_dtype, uint8_t* mask, const int64_t mask_stride[3],
                               const int64_t dims[3], int reduce_op, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SP_REQUIRE(dims[0] >= 0 && dims[1] >= 0 && dims[2] >= 0 && reduce_op <= SP_RED_ANY, SP_ERR_INVALID,
             "sp_merge_masked: bad arguments");
  const int64_t total = dims[0] * dims[1] * dims[2];
  if (total == 0) return SP_OK;
  SP_REQUIRE(dst && src && mask, SP_ERR_INVALID, "sp_merge_masked: null pointer");
  const int g = grid_for(total);
#define SP_MERGE(DT, ST, CT)                                                                                          \
  merge_masked_kernel<DT, ST, CT><<<g, 256, 0, stream>>>(static_cast<DT*>(dst), dst_stride[0], dst_stride[1],          \
      dst_stride[2], static_cast<const ST*>(src), src_stride[0], src_stride[1], src_stride[2], mask, mask_stride[0],   \
      mask_stride[1], mask_stride[2], dims[0], dims[1], dims[2], reduce_op)
  const int key = dst_dtype * 16 + src_dtype;
  switch (key) {
    case SP_F32 * 16 + SP_F32: SP_MERGE(float, float, float); break;
    case SP_F32 * 16 + SP_F64: SP_MERGE(float, double, double); break;
    case SP_F64 * 16 + SP_F64: SP_MERGE(double, double, double); break;
    case SP_F64 * 16 + SP_F32: SP_MERGE(double, float, double); break;
    case SP_I64 * 16 + SP_I64: SP_MERGE(long long, long long, long long); break;
    case SP_I64 * 16 + SP_I32: SP_MERGE(long long, int32_t, long long); break;
    case SP_I32 * 16 + SP_I32: SP_MERGE(int32_t, int32_t, int32_t); break;
    case SP_I32 * 16 + SP_I64: SP_MERGE(int32_t, long long, long long); break;
    case SP_U8 * 16 + SP_U8:
    case SP_BOOL * 16 + SP_BOOL: SP_MERGE(uint8_t, uint8_t, uint8_t); break;
    default:
      set_error("sp_merge_masked: unsupported dtype pair (%d, %d)", dst_dtype, src_dtype);
      return SP_ERR_UNSUPPORTED;
  }
#undef SP_MERGE
  SP_CUDA_CHECK(cudaGetLastError());
  return SP_OK;
}


// dst (R x C, leading dimension ldd) = transpose of src (C x R, leading dimension lds); elem_size in {1, 4, 8} bytes.
extern "C" int sp_transpose_2d(void* dst, int64_t ldd, const void* src, int64_t lds, int64_t R, int64_t C, int elem_size,
                               void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SP_REQUIRE(R >= 0 && C >= 0 && ldd >= C && lds >= R, SP_ERR_INVALID, "sp_transpose_2d: bad shape");
  if (R == 0 || C == 0) return SP_OK;
  SP_REQUIRE(dst && src, SP_ERR_INVALID, "sp_transpose_2d: null pointer");
  const int64_t gx = (R + 31) / 32, gy = (C + 31) / 32;
  SP_REQUIRE(gy <= 65535 && gx < (1ll << 31), SP_ERR_INVALID, "sp_transpose_2d: too many columns for one launch");
  dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>(gy)), block(32, 8);
  switch (elem_size) {
    case 1: transpose_2d_kernel<uint8_t><<<grid, block, 0, stream>>>(static_cast<uint8_t*>(dst), ldd, static_cast<const uint8_t*>(src), lds, R, C); break;
    case 4: transpose_2d_kernel<uint32_t><<<grid, block, 0, stream>>>(static_cast<uint32_t*>(dst), ldd, static_cast<const uint32_t*>(src), lds, R, C); break;
    case 8: transpose_2d_kernel<uint64_t><<<grid, block, 0, stream>>>(static_cast<uint64_t*>(dst), ldd, static_cast<const uint64_t*>(src), lds, R, C); break;
    default:
      set_error("sp_transpose_2d: bad element size %d", elem_size);
      return SP_ERR_INVALID;
  }
  SP_CUDA_CHECK(cudaGetLastError());
  return SP_OK;
}
