// Run-time specialisation of the streaming fused map / map+reduce kernel.
//
// The reference turns a fused LocalExpr tree into *source code* and compiles it per expression
// (spartan/expr/operator/local.py:58-152 `codegen`, used with parakeet / numexpr).  The B200-native
// counterpart: the accumulator-machine code of a fused tree becomes the template argument pack of
// StaticProgram<...> (interp.cuh) and stream_kernel<T, NI, MODE, StaticProgram<...>> (stream_kernels.cuh)
// is instantiated by NVRTC for sm_100a -- the same hand-written kernel the static catalogue uses, with
// every dispatch folded away, so an arbitrary fused chain is HBM-bound instead of instruction-bound.
// The kernel sources are the very headers the library was built from (embedded at build time,
// build/jit_embedded.inc); NVRTC is dlopen()ed on first use.  If it is unavailable or a compile fails
// the caller falls back to the interpreter (DynamicProgram) -- still on the GPU.
#include "sp_common.h"
#include <cuda.h>
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "build/jit_embedded.inc"   // kJitHeaderNames[], kJitHeaderSources[], kJitNumHeaders

namespace sp {
namespace jit {

// ---- the slice of the NVRTC API that is used (nvrtc.h is not needed at build time)
typedef struct _nvrtcProgram* nvrtcProgram;
typedef int nvrtcResult;
struct Nvrtc {
  void* handle = nullptr;
  nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*);
  nvrtcResult (*DestroyProgram)(nvrtcProgram*);
  nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*);
  nvrtcResult (*AddNameExpression)(nvrtcProgram, const char*);
  nvrtcResult (*GetLoweredName)(nvrtcProgram, const char*, const char**);
  nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*);
  nvrtcResult (*GetCUBIN)(nvrtcProgram, char*);
  nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*);
  nvrtcResult (*GetProgramLog)(nvrtcProgram, char*);
  const char* (*GetErrorString)(nvrtcResult);
};

struct Driver {
  CUresult (*ModuleLoadData)(CUmodule*, const void*);
  CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*);
  CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int);
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream,
                           void**, void**);
};

static std::mutex g_mu;
static Nvrtc g_nvrtc;
static Driver g_drv;
static bool g_nvrtc_tried = false, g_drv_tried = false, g_drv_ok = false;
static int g_enabled = -1;                 // -1: read SPARTAN_JIT on first use
static std::string g_nvrtc_path;
static std::string g_log;
static int64_t g_compiled = 0, g_launches = 0, g_failures = 0;
static std::map<std::string, CUfunction> g_cache;        // key -> kernel (nullptr = compile failed, do not retry)

static bool load_nvrtc() {
  if (g_nvrtc_tried) return g_nvrtc.handle != nullptr;
  g_nvrtc_tried = true;
  std::vector<std::string> cands;
  if (!g_nvrtc_path.empty()) cands.push_back(g_nvrtc_path);
  if (const char* e = getenv("SPARTAN_NVRTC")) cands.push_back(e);
  cands.push_back("libnvrtc.so.12");
  cands.push_back("/usr/local/cuda/lib64/libnvrtc.so.12");
  cands.push_back("libnvrtc.so");
  for (const std::string& c : cands) {
    void* h = dlopen(c.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!h) continue;
    Nvrtc n;
    n.handle = h;
    bool ok = true;
#define SP_SYM(NAME)                                                        \
  *reinterpret_cast<void**>(&n.NAME) = dlsym(h, "nvrtc" #NAME);             \
  ok = ok && (n.NAME != nullptr);
    SP_SYM(CreateProgram) SP_SYM(DestroyProgram) SP_SYM(CompileProgram) SP_SYM(AddNameExpression)
    SP_SYM(GetLoweredName) SP_SYM(GetCUBINSize) SP_SYM(GetCUBIN) SP_SYM(GetProgramLogSize) SP_SYM(GetProgramLog)
    SP_SYM(GetErrorString)
#undef SP_SYM
    if (ok) { g_nvrtc = n; return true; }
    dlclose(h);
  }
  g_log = "libnvrtc.so.12 not found (set SPARTAN_NVRTC or sp_jit_set_nvrtc_path)";
  return false;
}

static bool load_driver() {
  if (g_drv_tried) return g_drv_ok;
  g_drv_tried = true;
  bool ok = true;
  auto get = [&](const char* name, void** out) {
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, out, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      ok = false;
    }
  };
  get("cuModuleLoadData", reinterpret_cast<void**>(&g_drv.ModuleLoadData));
  get("cuModuleGetFunction", reinterpret_cast<void**>(&g_drv.ModuleGetFunction));
  get("cuFuncSetAttribute", reinterpret_cast<void**>(&g_drv.FuncSetAttribute));
  get("cuLaunchKernel", reinterpret_cast<void**>(&g_drv.LaunchKernel));
  g_drv_ok = ok;
  if (!ok) g_log = "CUDA driver entry points for module loading are not available";
  return ok;
}

static bool enabled() {
  if (g_enabled < 0) {
    const char* e = getenv("SPARTAN_JIT");
    g_enabled = (e && e[0] == '0') ? 0 : 1;
  }
  return g_enabled == 1;
}

static const char* type_name(int dtype) {
  return dtype == SP_F32 ? "float" : dtype == SP_F64 ? "double" : "long long";
}

// "sp::stream::stream_kernel<float, 2, 1, sp::StaticProgram<65536, ...> >"
static std::string kernel_expr(int dtype, int ni, int mode, const uint8_t* op, const uint8_t* src, const uint8_t* arg, int n) {
  // modes 0-2: the ring kernel (map / reduce leading axis / reduce trailing axis); mode 3: the direct map kernel
  // mode 4: the direct reduce kernel (leading axis)
  std::string s = mode == 3 ? "sp::stream::direct_map_kernel<" : mode == 4 ? "sp::stream::direct_reduce_kernel<"
                                                                             : "sp::stream::stream_kernel<";
  s += type_name(dtype);
  s += ", " + std::to_string(ni) + (mode >= 3 ? std::string("") : ", " + std::to_string(mode)) + ", sp::StaticProgram<";
  for (int i = 0; i < n; ++i) {
    if (i) s += ", ";
    s += std::to_string((static_cast<int>(op[i]) << 16) | (static_cast<int>(src[i]) << 8) | arg[i]);
  }
  s += "> >";
  return s;
}

// Compiles one specialisation to a cubin.  Needs no GPU (so the build check can exercise it).
static bool compile(const std::string& expr, std::vector<char>* cubin, std::string* lowered) {
  if (!load_nvrtc()) return false;
  const std::string source = "#include \"stream_kernels.cuh\"\n";
  nvrtcProgram prog = nullptr;
  nvrtcResult r = g_nvrtc.CreateProgram(&prog, source.c_str(), "spartan_jit.cu", kJitNumHeaders, kJitHeaderSources,
                                        kJitHeaderNames);
  if (r != 0) { g_log = std::string("nvrtcCreateProgram: ") + g_nvrtc.GetErrorString(r); return false; }
  const std::string name_expr = "&" + expr;
  g_nvrtc.AddNameExpression(prog, name_expr.c_str());
  // same code generation flags as the Makefile: sm_100a, no FMA contraction (NumPy rounds after every ufunc)
  // (-default-device: the C-ABI prototypes in the embedded header are plain declarations; NVRTC accepts no host code)
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "--fmad=false", "-lineinfo", "-default-device"};
  r = g_nvrtc.CompileProgram(prog, 5, opts);
  if (r != 0) {
    size_t n = 0;
    g_nvrtc.GetProgramLogSize(prog, &n);
    std::string log(n, '\0');
    if (n) g_nvrtc.GetProgramLog(prog, &log[0]);
    g_log = std::string("nvrtcCompileProgram: ") + g_nvrtc.GetErrorString(r) + "\n" + log;
    g_nvrtc.DestroyProgram(&prog);
    return false;
  }
  const char* low = nullptr;
  r = g_nvrtc.GetLoweredName(prog, name_expr.c_str(), &low);
  size_t sz = 0;
  if (r == 0) r = g_nvrtc.GetCUBINSize(prog, &sz);
  if (r != 0 || sz == 0 || low == nullptr) {
    g_log = std::string("nvrtc lowered name / cubin: ") + g_nvrtc.GetErrorString(r);
    g_nvrtc.DestroyProgram(&prog);
    return false;
  }
  *lowered = low;
  cubin->resize(sz);
  r = g_nvrtc.GetCUBIN(prog, cubin->data());
  g_nvrtc.DestroyProgram(&prog);
  if (r != 0) { g_log = std::string("nvrtcGetCUBIN: ") + g_nvrtc.GetErrorString(r); return false; }
  return true;
}

// Returns the specialised kernel for (dtype, NI, MODE, accumulator code) or nullptr (-> use the interpreter).
static CUfunction get_kernel(int dtype, int ni, int mode, const uint8_t* op, const uint8_t* src, const uint8_t* arg,
                             int n, int smem_bytes) {
  if (!enabled()) return nullptr;
  const std::string expr = kernel_expr(dtype, ni, mode, op, src, arg, n);
  auto it = g_cache.find(expr);
  if (it != g_cache.end()) return it->second;
  CUfunction fn = nullptr;
  std::vector<char> cubin;
  std::string lowered;
  if (load_driver() && compile(expr, &cubin, &lowered)) {
    CUmodule mod = nullptr;
    CUresult cr = g_drv.ModuleLoadData(&mod, cubin.data());
    if (cr == CUDA_SUCCESS) cr = g_drv.ModuleGetFunction(&fn, mod, lowered.c_str());
    if (cr == CUDA_SUCCESS) cr = g_drv.FuncSetAttribute(fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, smem_bytes);
    if (cr != CUDA_SUCCESS) {
      g_log = "loading the specialised kernel failed with CUresult " + std::to_string(static_cast<int>(cr));
      fn = nullptr;
    }
  }
  if (fn) ++g_compiled; else ++g_failures;
  g_cache[expr] = fn;
  return fn;
}

// Launches the specialisation if there is one.  Returns 1 when launched, 0 when the caller should use the interpreter,
// < 0 on a launch error.  `params` = addresses of the kernel arguments (DevProgram, DevOperands, Plan, red_op, scratch).
int launch_stream_specialised(int dtype, int ni, int mode, const uint8_t* op, const uint8_t* src, const uint8_t* arg, int n,
                              void** params, int grid, int threads, int smem_bytes, cudaStream_t stream) {
  std::lock_guard<std::mutex> lock(g_mu);
  CUfunction fn = get_kernel(dtype, ni, mode, op, src, arg, n, smem_bytes);
  if (!fn) return 0;
  CUresult cr = g_drv.LaunchKernel(fn, static_cast<unsigned>(grid), 1, 1, static_cast<unsigned>(threads), 1, 1,
                                   static_cast<unsigned>(smem_bytes), reinterpret_cast<CUstream>(stream), params, nullptr);
  if (cr != CUDA_SUCCESS) {
    set_error("cuLaunchKernel of a specialised stream kernel failed with CUresult %d", static_cast<int>(cr));
    return SP_ERR_CUDA;
  }
  ++g_launches;
  return 1;
}

}  // namespace jit
}  // namespace sp

using namespace sp;

// 1 = specialise fused chains outside the static catalogue at run time (default; env SPARTAN_JIT=0 disables), 0 = always
// interpret them.  A tuning / test hook.
extern "C" int sp_jit_enable(int on) {
  std::lock_guard<std::mutex> lock(jit::g_mu);
  jit::g_enabled = on ? 1 : 0;
  return SP_OK;
}

// Where libnvrtc.so.12 lives, if not on the loader path (call before the first specialisation).
extern "C" int sp_jit_set_nvrtc_path(const char* path) {
  std::lock_guard<std::mutex> lock(jit::g_mu);
  SP_REQUIRE(path != nullptr, SP_ERR_INVALID, "null path");
  jit::g_nvrtc_path = path;
  jit::g_nvrtc_tried = false;
  return SP_OK;
}

extern "C" int sp_jit_stats(int64_t* compiled, int64_t* launches, int64_t* failures) {
  std::lock_guard<std::mutex> lock(jit::g_mu);
  if (compiled) *compiled = jit::g_compiled;
  if (launches) *launches = jit::g_launches;
  if (failures) *failures = jit::g_failures;
  return SP_OK;
}

// Message of the last specialisation failure (NVRTC log), "" if none.
extern "C" const char* sp_jit_last_log(void) {
  std::lock_guard<std::mutex> lock(jit::g_mu);
  return jit::g_log.c_str();
}

// Compiles (does not load or run) the specialisation of `prog` for `n_in` operands: the build-time check that the
// run-time path works on a machine without a GPU.  mode 0 = map, 1 = map+reduce over the leading axis, 2 = map+reduce over the trailing axis, 3 = the direct (non-ring) map kernel, 4 = the direct reduce kernel (leading axis).  Returns the cubin size.
namespace sp { int lower_for_jit(const sp_program* prog, uint8_t* op, uint8_t* src, uint8_t* arg, int* n); }
extern "C" int64_t sp_jit_compile_check(const sp_program* prog, int n_in, int mode) {
  SP_REQUIRE(prog != nullptr && n_in >= 0 && n_in <= SP_MAX_OPERANDS && (mode >= 0 && mode <= 4), SP_ERR_INVALID,
             "sp_jit_compile_check: bad arguments");
  uint8_t op[SP_MAX_PROGRAM], src[SP_MAX_PROGRAM], arg[SP_MAX_PROGRAM];
  int n = 0;
  int rc = lower_for_jit(prog, op, src, arg, &n);
  if (rc) return rc;
  std::lock_guard<std::mutex> lock(jit::g_mu);
  std::vector<char> cubin;
  std::string lowered;
  const std::string expr = jit::kernel_expr(prog->compute_dtype, n_in <= 2 ? 2 : 8, mode, op, src, arg, n);
  if (!jit::compile(expr, &cubin, &lowered)) {
    set_error("%s", jit::g_log.c_str());
    return SP_ERR_UNSUPPORTED;
  }
  return static_cast<int64_t>(cubin.size());
}
