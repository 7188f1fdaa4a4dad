// Peer memory for the cross-rank steps of the tile evaluator: the dot row x column shuffle
// (spartan/expr/dot.py:195-238 -- join_mapper fetches the remote strips over RPC, spartan/expr/operator/map.py:243-286)
// and any other exchange where a rank hands a buffer to its peers.
//
// One process per GPU.  Each rank allocates "symmetric" buffers with sp_peer_alloc (plain cudaMalloc, outside any
// caching allocator so that the IPC handle maps exactly this buffer), exports them with sp_peer_export, and opens the
// other ranks' handles with sp_peer_import.  Data then moves with sp_peer_push: copy-engine transfers over
// NVLink/NVSwitch (cudaMemcpyAsync between a local and an imported pointer) each followed, in stream order, by a
// 4-byte copy of an epoch word into a flag slot of the destination rank.  No SM is involved, so a persistent
// tensor-core kernel that occupies every SM keeps running while operands arrive; the consumer kernel polls the flag
// (gemm_tcgen05.cu, sp_gemm_prepared_views_gated) and starts on a K segment the moment it has landed.
#include "sp_common.h"
#include <string.h>

namespace sp {

__global__ void write_u32_kernel(unsigned int* p, unsigned int v) {
  *p = v;
  __threadfence_system();
}

// Spins until *flag reaches `value` (wrap-safe) or ~timeout_ns passed; a time-out is recorded in *status.
__global__ void wait_u32_kernel(const unsigned int* flag, unsigned int value, unsigned long long timeout_ns,
                                unsigned int* status) {
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    unsigned int seen;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
    if (static_cast<int>(seen - value) >= 0) return;
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > timeout_ns) {
      if (status) atomicAdd(status, 1u);
      return;
    }
    __nanosleep(200);
  }
}

}  // namespace sp

using namespace sp;

extern "C" int sp_peer_alloc(int64_t bytes, void** out) {
  SP_REQUIRE(bytes > 0 && out != nullptr, SP_ERR_INVALID, "sp_peer_alloc: bad arguments");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, static_cast<size_t>(bytes));
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("sp_peer_alloc: cudaMalloc(%lld) failed: %s", (long long)bytes, cudaGetErrorString(e));
    return e == cudaErrorMemoryAllocation ? SP_ERR_NOMEM : SP_ERR_CUDA;
  }
  SP_CUDA_CHECK(cudaMemset(p, 0, static_cast<size_t>(bytes)));
  *out = p;
  return SP_OK;
}

extern "C" int sp_peer_free(void* p) {
  if (p) SP_CUDA_CHECK(cudaFree(p));
  return SP_OK;
}

extern "C" int sp_peer_handle_bytes(void) { return static_cast<int>(sizeof(cudaIpcMemHandle_t)); }

extern "C" int sp_peer_export(void* p, void* handle_out) {
  SP_REQUIRE(p && handle_out, SP_ERR_INVALID, "sp_peer_export: null pointer");
  cudaIpcMemHandle_t h;
  SP_CUDA_CHECK(cudaIpcGetMemHandle(&h, p));
  memcpy(handle_out, &h, sizeof(h));
  return SP_OK;
}

extern "C" int sp_peer_import(const void* handle, void** out) {
  SP_REQUIRE(handle && out, SP_ERR_INVALID, "sp_peer_import: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  SP_CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *out = p;
  return SP_OK;
}

extern "C" int sp_peer_close(void* p) {
  if (p) SP_CUDA_CHECK(cudaIpcCloseMemHandle(p));
  return SP_OK;
}

// For every destination i: dst[i][0:bytes) <- src (copy engine), then, in stream order, flag_dst[i] <- *flag_src.
// `src` and `flag_src` are local device pointers, dst / flag_dst local or imported ones.  flag_dst may be NULL.
extern "C" int sp_peer_push(int n, void* const* dst, const void* src, int64_t bytes, void* const* flag_dst,
                            const void* flag_src, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SP_REQUIRE(n >= 0 && (n == 0 || (dst && src)) && bytes >= 0, SP_ERR_INVALID, "sp_peer_push: bad arguments");
  for (int i = 0; i < n; ++i) {
    if (bytes > 0) SP_CUDA_CHECK(cudaMemcpyAsync(dst[i], src, static_cast<size_t>(bytes), cudaMemcpyDefault, stream));
    if (flag_dst && flag_dst[i] && flag_src)
      SP_CUDA_CHECK(cudaMemcpyAsync(flag_dst[i], flag_src, 4, cudaMemcpyDefault, stream));
  }
  return SP_OK;
}

// 2-D form of the data copy (row ranges of a pitched buffer), same flag protocol.
extern "C" int sp_peer_push_2d(int n, void* const* dst, int64_t dst_pitch, const void* src, int64_t src_pitch,
                               int64_t width_bytes, int64_t rows, void* const* flag_dst, const void* flag_src,
                               void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SP_REQUIRE(n >= 0 && (n == 0 || (dst && src)) && width_bytes >= 0 && rows >= 0, SP_ERR_INVALID,
             "sp_peer_push_2d: bad arguments");
  for (int i = 0; i < n; ++i) {
    if (width_bytes > 0 && rows > 0)
      SP_CUDA_CHECK(cudaMemcpy2DAsync(dst[i], static_cast<size_t>(dst_pitch), src, static_cast<size_t>(src_pitch),
                                      static_cast<size_t>(width_bytes), static_cast<size_t>(rows), cudaMemcpyDefault,
                                      stream));
    if (flag_dst && flag_dst[i] && flag_src)
      SP_CUDA_CHECK(cudaMemcpyAsync(flag_dst[i], flag_src, 4, cudaMemcpyDefault, stream));
  }
  return SP_OK;
}

extern "C" int sp_write_u32(void* p, uint32_t value, void* stream_) {
  SP_REQUIRE(p != nullptr, SP_ERR_INVALID, "sp_write_u32: null pointer");
  write_u32_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream_)>>>(static_cast<unsigned int*>(p), value);
  SP_CUDA_CHECK(cudaGetLastError());
  return SP_OK;
}

// Stream-side wait: the work queued on `stream` after this call starts once *flag >= value (or after timeout_ms,
// which increments *status when status != NULL).
extern "C" int sp_wait_u32(const void* flag, uint32_t value, int64_t timeout_ms, void* status, void* stream_) {
  SP_REQUIRE(flag != nullptr && timeout_ms > 0, SP_ERR_INVALID, "sp_wait_u32: bad arguments");
  wait_u32_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream_)>>>(static_cast<const unsigned int*>(flag), value,
                                                                 static_cast<unsigned long long>(timeout_ms) * 1000000ull,
                                                                 static_cast<unsigned int*>(status));
  SP_CUDA_CHECK(cudaGetLastError());
  return SP_OK;
}
