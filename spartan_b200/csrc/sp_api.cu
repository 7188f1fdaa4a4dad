// Error state, version and device queries for the spartan_b200 C ABI.
#include "sp_common.h"
#include <string.h>

namespace sp {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int cached = 0;
  if (cached) return cached;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
  cached = n;
  return n;
}

}  // namespace sp

extern "C" const char* sp_last_error(void) { return sp::g_err; }

extern "C" int sp_version(void) { return 100; }

extern "C" int sp_device_info(int* n_sms, int64_t* hbm_bytes, int* cc_major, int* cc_minor) {
  int dev = 0;
  SP_CUDA_CHECK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  SP_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
  if (n_sms) *n_sms = prop.multiProcessorCount;
  if (hbm_bytes) *hbm_bytes = static_cast<int64_t>(prop.totalGlobalMem);
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return SP_OK;
}
