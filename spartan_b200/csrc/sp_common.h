// Shared declarations for the spartan_b200 CUDA shim (sm_100a only).
//
// Error convention for every C-ABI entry point (include/spartan_b200.h):
// return 0 on success, a negative sp_status on failure; the message of the
// most recent failure on the calling thread is available from sp_last_error().
#pragma once
#ifdef __CUDACC_RTC__
// Run-time compilation of a fused kernel (jit.cu): device code only, headers come from memory.
#include "spartan_b200.h"
#else
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/spartan_b200.h"

namespace sp {

void set_error(const char* fmt, ...);

inline int fail_cuda(cudaError_t e, const char* what, const char* file, int line) {
  set_error("%s failed at %s:%d: %s", what, file, line, cudaGetErrorString(e));
  return SP_ERR_CUDA;
}

#define SP_CUDA_CHECK(expr)                                            \
  do {                                                                 \
    cudaError_t _e = (expr);                                           \
    if (_e != cudaSuccess) return sp::fail_cuda(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define SP_REQUIRE(cond, code, ...)        \
  do {                                     \
    if (!(cond)) {                         \
      sp::set_error(__VA_ARGS__);          \
      return (code);                       \
    }                                      \
  } while (0)

inline size_t dtype_size(int dt) {
  switch (dt) {
    case SP_F32: return 4;
    case SP_F64: return 8;
    case SP_I32: return 4;
    case SP_I64: return 8;
    case SP_U8:  return 1;
    case SP_BOOL: return 1;
    default: return 0;
  }
}

int num_sms();

}  // namespace sp
#endif  // !__CUDACC_RTC__
