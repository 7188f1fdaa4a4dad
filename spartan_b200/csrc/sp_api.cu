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


// ---- device, memory and stream housekeeping for hosts that do not bring their own allocator / stream library
// (the Python host uses torch for these; SURVEY.md section 8b lists them as part of a self-sufficient boundary) ----
extern "C" int sp_init(int device) {
  SP_CUDA_CHECK(cudaSetDevice(device));
  SP_CUDA_CHECK(cudaFree(nullptr));                 // creates the context
  return SP_OK;
}

extern "C" int sp_shutdown(void) {
  SP_CUDA_CHECK(cudaDeviceSynchronize());
  return SP_OK;
}

// A tile buffer of `bytes` bytes in the HBM of the current device (stream-ordered when `stream` is given).
extern "C" int sp_tile_alloc(int64_t bytes, void** out, void* stream_) {
  SP_REQUIRE(bytes >= 0 && out != nullptr, SP_ERR_INVALID, "sp_tile_alloc: bad arguments");
  void* p = nullptr;
  if (bytes > 0) {
    cudaError_t e = cudaMallocAsync(&p, static_cast<size_t>(bytes), static_cast<cudaStream_t>(stream_));
    if (e != cudaSuccess) {
      cudaGetLastError();
      sp::set_error("sp_tile_alloc(%lld): %s", (long long)bytes, cudaGetErrorString(e));
      return e == cudaErrorMemoryAllocation ? SP_ERR_NOMEM : SP_ERR_CUDA;
    }
  }
  *out = p;
  return SP_OK;
}

extern "C" int sp_tile_free(void* p, void* stream_) {
  if (p) SP_CUDA_CHECK(cudaFreeAsync(p, static_cast<cudaStream_t>(stream_)));
  return SP_OK;
}

extern "C" int sp_sync(void* stream_) {
  SP_CUDA_CHECK(cudaStreamSynchronize(static_cast<cudaStream_t>(stream_)));
  return SP_OK;
}

extern "C" int sp_event_create(void** out) {
  SP_REQUIRE(out != nullptr, SP_ERR_INVALID, "sp_event_create: null pointer");
  cudaEvent_t ev;
  SP_CUDA_CHECK(cudaEventCreate(&ev));
  *out = ev;
  return SP_OK;
}

extern "C" int sp_event_record(void* event, void* stream_) {
  SP_CUDA_CHECK(cudaEventRecord(static_cast<cudaEvent_t>(event), static_cast<cudaStream_t>(stream_)));
  return SP_OK;
}

// Milliseconds between two recorded events (waits for `stop`).
extern "C" int sp_event_elapsed(void* start, void* stop, float* ms) {
  SP_REQUIRE(ms != nullptr, SP_ERR_INVALID, "sp_event_elapsed: null pointer");
  SP_CUDA_CHECK(cudaEventSynchronize(static_cast<cudaEvent_t>(stop)));
  SP_CUDA_CHECK(cudaEventElapsedTime(ms, static_cast<cudaEvent_t>(start), static_cast<cudaEvent_t>(stop)));
  return SP_OK;
}

extern "C" int sp_event_destroy(void* event) {
  if (event) SP_CUDA_CHECK(cudaEventDestroy(static_cast<cudaEvent_t>(event)));
  return SP_OK;
}
