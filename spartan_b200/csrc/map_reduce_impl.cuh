// Kernels for the fused element-wise map and the fused map+reduce, templated on the compute
// type T (float / double / long long), vector width V and operand capacity NI.
//
//   map_kernel         out[d0,d1,d2] = f(in...)                  tile_mapper      map.py:48-88
//   reduce_col_kernel  out[d0,d2]    = red_{d1} f(in...)         _reduce_mapper   reduce.py:21-70
//   reduce_row_kernel  out[d0]       = red_{d1} f(in...), d2==1  (axis = last / axis=None)
//   finalize_kernel    combines the per-CTA partials in fixed order (deterministic, unlike the
//                      RPC-arrival order of Tile.merge, tile.pyx:263-283) and applies the
//                      accumulate-into-existing-output semantics of the combiner.
//
// All three are HBM-bound: algorithmic bytes = sum of operand bytes read once + output written once.
#pragma once
#include "interp.cuh"
#include "stream_kernels.cuh"
#include <algorithm>
#include <string.h>
#include <type_traits>

namespace sp {

struct Dims3 {
  int64_t d0, d1, d2;
};

// ------------------------------------------------------------------------------------ map
template <typename T, int V, int NI>
__global__ void __launch_bounds__(256)
map_kernel(const DevProgram<T> prog, const DevOperands<NI> ops, const Dims3 dims, const int64_t d2v,
           const int64_t total_vecs) {
  const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const bool flat = (dims.d0 * dims.d1 == 1);
  const bool small = total_vecs < (1ll << 31);
  for (int64_t vec = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; vec < total_vecs; vec += step) {
    int64_t i0 = 0, i1 = 0, i2v = vec;
    if (!flat) {
      if (small) {
        const uint32_t uv = static_cast<uint32_t>(vec), ud2v = static_cast<uint32_t>(d2v);
        const uint32_t t = uv / ud2v;
        i2v = uv - t * ud2v;
        const uint32_t ud1 = static_cast<uint32_t>(dims.d1);
        i0 = t / ud1;
        i1 = t - static_cast<uint32_t>(i0) * ud1;
      } else {
        const int64_t t = vec / d2v;
        i2v = vec - t * d2v;
        i0 = t / dims.d1;
        i1 = t - i0 * dims.d1;
      }
    }
    const int64_t i2 = i2v * V;
    const int64_t rem = dims.d2 - i2;
    const int valid = rem >= V ? V : static_cast<int>(rem);
    T in[NI][V];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      if (i < ops.n_in) {
        const DevOperand& o = ops.in[i];
        load_operand<T, V>(o, i0 * o.stride[0] + i1 * o.stride[1] + i2 * o.stride[2], o.stride[2], valid, in[i]);
      }
    }
    T res[V], idx[V];
    if (prog.uses_index) make_index<T, V>(prog, i0, i1, i2, 2, idx);
    run_program<T, V, NI>(prog, in, idx, res);
    const DevOperand& o = ops.out;
    store_operand<T, V>(o, i0 * o.stride[0] + i1 * o.stride[1] + i2 * o.stride[2], o.stride[2], valid, res);
  }
}

// ------------------------------------------------------------------------------------ reduce over d1, d2 > 1
// block = (32, 8): x walks vectors of the inner axis (coalesced 16B loads), y walks rows of the
// reduced axis; grid.y splits the reduced axis so the grid covers >= a few waves of the 148 SMs.
template <typename T, int V, int NI>
__global__ void __launch_bounds__(256)
reduce_col_kernel(const DevProgram<T> prog, const DevOperands<NI> ops, const Dims3 dims, const int64_t d2v,
                  const int red_op, T* __restrict__ scratch) {
  __shared__ T sm[8][32][V + 1];
  const int64_t gx = blockIdx.x * 32ll + threadIdx.x;      // over d0 * d2v
  const int64_t nx = dims.d0 * d2v;
  const bool active = gx < nx;
  const int64_t i0 = active ? gx / d2v : 0;
  const int64_t i2 = active ? (gx - i0 * d2v) * V : 0;
  const int64_t rem = dims.d2 - i2;
  const int valid = active ? (rem >= V ? V : static_cast<int>(rem)) : 0;
  T acc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) acc[v] = red_identity<T>(red_op);
  if (active) {
    const int64_t rstep = 8ll * gridDim.y;
    for (int64_t r = blockIdx.y * 8ll + threadIdx.y; r < dims.d1; r += rstep) {
      T in[NI][V];
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        if (i < ops.n_in) {
          const DevOperand& o = ops.in[i];
          load_operand<T, V>(o, i0 * o.stride[0] + r * o.stride[1] + i2 * o.stride[2], o.stride[2], valid, in[i]);
        }
      }
      T res[V], idx[V];
      if (prog.uses_index) make_index<T, V>(prog, i0, r, i2, 2, idx);
      run_program<T, V, NI>(prog, in, idx, res);
#pragma unroll
      for (int v = 0; v < V; ++v)
        if (v < valid) acc[v] = red_apply<T>(red_op, acc[v], res[v]);
    }
  }
#pragma unroll
  for (int v = 0; v < V; ++v) sm[threadIdx.y][threadIdx.x][v] = acc[v];
  __syncthreads();
  if (threadIdx.y == 0 && active) {
#pragma unroll
    for (int y = 1; y < 8; ++y)
#pragma unroll
      for (int v = 0; v < V; ++v) acc[v] = red_apply<T>(red_op, acc[v], sm[y][threadIdx.x][v]);
    T* dst = scratch + (static_cast<int64_t>(blockIdx.y) * dims.d0 + i0) * dims.d2 + i2;
#pragma unroll
    for (int v = 0; v < V; ++v)
      if (v < valid) dst[v] = acc[v];
  }
}

// ------------------------------------------------------------------------------------ reduce over d1, d2 == 1
// One row (d0 index) is reduced by `segs` CTAs; vectors run along the reduced axis.
template <typename T, int V, int NI>
__global__ void __launch_bounds__(256)
reduce_row_kernel(const DevProgram<T> prog, const DevOperands<NI> ops, const Dims3 dims, const int64_t nvec,
                  const int segs, const int red_op, T* __restrict__ scratch) {
  __shared__ T sm[8];
  const int64_t b = blockIdx.x;
  const int64_t i0 = b / segs;
  const int seg = static_cast<int>(b - i0 * segs);
  T acc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) acc[v] = red_identity<T>(red_op);
  const int64_t jstep = static_cast<int64_t>(segs) * blockDim.x;
  for (int64_t jv = static_cast<int64_t>(seg) * blockDim.x + threadIdx.x; jv < nvec; jv += jstep) {
    const int64_t r = jv * V;
    const int64_t rem = dims.d1 - r;
    const int valid = rem >= V ? V : static_cast<int>(rem);
    T in[NI][V];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      if (i < ops.n_in) {
        const DevOperand& o = ops.in[i];
        load_operand<T, V>(o, i0 * o.stride[0] + r * o.stride[1], o.stride[1], valid, in[i]);
      }
    }
    T res[V], idx[V];
    if (prog.uses_index) make_index<T, V>(prog, i0, r, 0, 1, idx);
    run_program<T, V, NI>(prog, in, idx, res);
#pragma unroll
    for (int v = 0; v < V; ++v)
      if (v < valid) acc[v] = red_apply<T>(red_op, acc[v], res[v]);
  }
  T a = acc[0];
#pragma unroll
  for (int v = 1; v < V; ++v) a = red_apply<T>(red_op, a, acc[v]);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) a = red_apply<T>(red_op, a, shfl_down_t<T>(a, d));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  if (lane == 0) sm[warp] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < nwarps; ++w) a = red_apply<T>(red_op, a, sm[w]);
    scratch[i0 * segs + seg] = a;
  }
}

// ------------------------------------------------------------------------------------ finalize
// out[o] = red_s scratch[s * s_stride + o * o_stride]  (optionally combined with the old out value)
template <typename T>
__global__ void finalize_kernel(const T* __restrict__ scratch, const int64_t n_out, const int splits,
                                const int64_t s_stride, const int64_t o_stride, const DevOperand out,
                                const int64_t d2, const int red_op, const int accumulate) {
  for (int64_t o = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; o < n_out;
       o += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    T a = scratch[o * o_stride];
    for (int s = 1; s < splits; ++s) a = red_apply<T>(red_op, a, scratch[s * s_stride + o * o_stride]);
    const int64_t i0 = o / d2, i2 = o - i0 * d2;
    const int64_t idx = i0 * out.stride[0] + i2 * out.stride[2];
    if (accumulate) a = red_apply<T>(red_op, load_as<T>(out.ptr, out.dtype, idx), a);
    store_as<T>(const_cast<void*>(out.ptr), out.dtype, idx, a);
  }
}

// ------------------------------------------------------------------------------------ host launchers
template <typename T> struct TypeTag;
template <> struct TypeTag<float> { static constexpr int dtype = SP_F32; static constexpr int VA = 8, VB = 4; };
template <> struct TypeTag<double> { static constexpr int dtype = SP_F64; static constexpr int VA = 4, VB = 2; };
template <> struct TypeTag<long long> { static constexpr int dtype = SP_I64; static constexpr int VA = 4, VB = 2; };

template <typename T>
static bool convert_program(const sp_program* prog, DevProgram<T>* out) {
  memset(out, 0, sizeof(*out));
  const bool ok = lower_program<T>(prog, out);      // postfix (ABI) -> accumulator code
  for (int i = 0; i < 3; ++i) out->index_stride[i] = prog->index_stride[i];
  out->index_base = prog->index_base;
  for (int i = 0; i < SP_MAX_CONSTS; ++i) {
    if (TypeTag<T>::dtype == SP_I64) out->consts[i] = static_cast<T>(prog->iconsts[i]);
    else out->consts[i] = static_cast<T>(prog->consts[i]);
  }
  return ok;
}

// vec_axis: which of stride[] runs along the vector (2 for map / reduce_col, 1 for reduce_row)
template <typename T, int V>
static DevOperand make_operand(const sp_operand& o, int vec_axis, const int64_t dims[3]) {
  DevOperand d;
  d.ptr = o.ptr;
  d.dtype = o.dtype;
  for (int i = 0; i < 3; ++i) d.stride[i] = o.stride[i];
  const int64_t vs = o.stride[vec_axis];
  bool aligned = (reinterpret_cast<uint64_t>(o.ptr) % 16 == 0);
  for (int i = 0; i < 3; ++i) {
    if (i == vec_axis || dims[i] == 1) continue;
    if ((o.stride[i] * static_cast<int64_t>(sizeof(T))) % 16 != 0) aligned = false;
  }
  if (vs == 0) d.kind = kSplat;
  else if (vs == 1 && o.dtype == TypeTag<T>::dtype && aligned) d.kind = kVec;
  else d.kind = kGeneric;
  return d;
}


// ------------------------------------------------------------------------------------ streaming fast path
constexpr int kStreamReduceRows = 128;   // rows per reduction work unit (one partial row per unit)

template <typename T>
static int64_t stream_reduce_chunks(const int64_t dims[3]) {
  return (dims[1] + kStreamReduceRows - 1) / kStreamReduceRows;
}

// Fills `plan` and returns true when the launch can use stream_kernel.
template <typename T, int NI>
static bool plan_stream(const DevOperands<NI>& ops, const int64_t dims[3], bool has_out, int rc_rows, stream::Plan* plan) {
  const int64_t row_bytes = dims[2] * static_cast<int64_t>(sizeof(T));
  if (row_bytes < stream::kSegBytes || (row_bytes % 16) != 0) return false;
  if (has_out && ops.out.kind != kVec) return false;
  int n_stream = 0;
  for (int i = 0; i < NI; ++i) plan->stream_slot[i] = -1;
  for (int i = 0; i < ops.n_in; ++i) {
    if (ops.in[i].kind == kVec) plan->stream_slot[i] = n_stream++;
  }
  if (n_stream == 0 || n_stream > 8) return false;
  // rows per stage: 16+ warp-items (rows x 4 segments) per stage whenever a stage of <= 64 KiB can hold them
  const int rb = n_stream == 1 ? 8 : n_stream <= 4 ? 4 : 2;
  plan->stage_bytes = n_stream * rb * stream::kPanelBytes;
  // 4 x 32 KiB measured best for one and two operands (6 stages cost ~7% on the two-operand map+reduce)
  plan->n_stages = std::min(4, stream::kRingBytes / plan->stage_bytes);
  plan->d0 = dims[0]; plan->d1 = dims[1]; plan->d2 = dims[2];
  plan->n_panels = static_cast<int>((row_bytes + stream::kPanelBytes - 1) / stream::kPanelBytes);
  plan->rb = rb;
  plan->rc = rc_rows > 0 ? rc_rows : rb * 4;
  plan->n_chunks = (dims[1] + plan->rc - 1) / plan->rc;
  plan->n_units = dims[0] * plan->n_chunks * plan->n_panels;
  plan->n_stream = n_stream;
  return plan->n_units > 0;
}

template <typename T, int NI, int MODE, typename PROG>
static int launch_stream_as(const DevProgram<T>& dp, const DevOperands<NI>& ops, const stream::Plan& plan, int red_op,
                            T* scratch, cudaStream_t stream_) {
  static bool attr_set = false;
  if (!attr_set) {
    SP_CUDA_CHECK(cudaFuncSetAttribute(stream::stream_kernel<T, NI, MODE, PROG>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, stream::kSmemBytes));
    attr_set = true;
  }
  const int grid = static_cast<int>(std::min<int64_t>(plan.n_units, num_sms()));
  stream::stream_kernel<T, NI, MODE, PROG><<<grid, stream::kThreads, stream::kSmemBytes, stream_>>>(dp, ops, plan, red_op,
                                                                                                  scratch);
  SP_CUDA_CHECK(cudaGetLastError());
  return SP_OK;
}

namespace jit {
int launch_stream_specialised(int dtype, int ni, int mode, const uint8_t* op, const uint8_t* src, const uint8_t* arg, int n,
                              void** params, int grid, int threads, int smem_bytes, cudaStream_t stream);
}
constexpr int64_t kJitMinBytes = 1 << 20;

// Index of the statically compiled program equal to `dp`, or -1.
template <typename T>
static int match_static(const DevProgram<T>& dp) {
#define SP_MATCH(IDX, TYPE) if (TYPE::matches(dp.op, dp.src, dp.arg, dp.n_ops)) return IDX;
  SP_STATIC_PROGRAMS(SP_MATCH)
#undef SP_MATCH
  return -1;
}

// Returns SP_OK when a kernel was launched, < 0 on error, and kNotLaunched when `interp_ok` is false and there is no
// compiled (static or run-time specialised) instance for this program.
constexpr int kNotLaunched = 1;
template <typename T, int NI, int MODE>
static int launch_stream(const DevProgram<T>& dp, const DevOperands<NI>& ops, const stream::Plan& plan, int red_op,
                         T* scratch, cudaStream_t stream_, bool interp_ok = true) {
  // statically compiled programs exist for the float / double two-operand kernels
  if constexpr (NI == 2 && !std::is_same<T, long long>::value) {
    switch (match_static<T>(dp)) {
#define SP_LAUNCH(IDX, TYPE) case IDX: return launch_stream_as<T, NI, MODE, TYPE>(dp, ops, plan, red_op, scratch, stream_);
      SP_STATIC_PROGRAMS(SP_LAUNCH)
#undef SP_LAUNCH
      default: break;
    }
  }
  // any other fused chain: specialise the same kernel for it at run time (jit.cu) when the launch is large enough
  // for the one-off compile to pay; otherwise, or if NVRTC is unavailable, interpret it
  if (plan.d0 * plan.d1 * plan.d2 * static_cast<int64_t>(sizeof(T)) >= kJitMinBytes) {
    const int grid = static_cast<int>(std::min<int64_t>(plan.n_units, num_sms()));
    int red = red_op;
    void* params[] = {const_cast<DevProgram<T>*>(&dp), const_cast<DevOperands<NI>*>(&ops),
                      const_cast<stream::Plan*>(&plan), &red, &scratch};
    const int rc = jit::launch_stream_specialised(TypeTag<T>::dtype, NI, MODE, dp.op, dp.src, dp.arg, dp.n_ops, params,
                                                  grid, stream::kThreads, stream::kSmemBytes, stream_);
    if (rc != 0) return rc < 0 ? rc : SP_OK;
  }
  if (!interp_ok) return kNotLaunched;
  return launch_stream_as<T, NI, MODE, DynamicProgram>(dp, ops, plan, red_op, scratch, stream_);
}

// ---- direct map kernel (compiled programs only): see stream_kernels.cuh
static int g_map_kernel = -1;     // 0 = ring (bulk-copy ring), 1 = direct; env SPARTAN_MAP_KERNEL=ring|direct, default direct
static bool use_direct_map() {
  if (g_map_kernel < 0) {
    const char* e = getenv("SPARTAN_MAP_KERNEL");
    g_map_kernel = (e && e[0] == 'r') ? 0 : 1;
  }
  return g_map_kernel == 1;
}

template <typename T, int NI, typename PROG>
static int launch_direct_as(const DevProgram<T>& dp, const DevOperands<NI>& ops, const stream::Plan& plan, int64_t blocks,
                            cudaStream_t stream_) {
  stream::direct_map_kernel<T, NI, PROG><<<static_cast<unsigned>(blocks), stream::kDirectThreads, 0, stream_>>>(dp, ops, plan);
  SP_CUDA_CHECK(cudaGetLastError());
  return SP_OK;
}

// SP_OK when launched, kNotLaunched when there is no compiled instance of the program (or the shape does not fit).
template <typename T, int NI>
static int launch_direct(const DevProgram<T>& dp, const DevOperands<NI>& ops, const stream::Plan& plan, cudaStream_t stream_) {
  constexpr int VEC = 16 / sizeof(T);
  if (plan.d2 % VEC != 0) return kNotLaunched;
  for (int i = 0; i < ops.n_in; ++i)
    if (ops.in[i].kind == kGeneric) return kNotLaunched;
  const int64_t row_vecs = plan.d2 / VEC;
  const int64_t per_block = static_cast<int64_t>(stream::kDirectThreads) * stream::kDirectUnroll;
  const int64_t blocks = plan.d0 * plan.d1 * ((row_vecs + per_block - 1) / per_block);
  if (blocks <= 0 || blocks >= (1ll << 31)) return kNotLaunched;
  if constexpr (NI == 2 && !std::is_same<T, long long>::value) {
    switch (match_static<T>(dp)) {
#define SP_LAUNCH(IDX, TYPE) case IDX: return launch_direct_as<T, NI, TYPE>(dp, ops, plan, blocks, stream_);
      SP_STATIC_PROGRAMS(SP_LAUNCH)
#undef SP_LAUNCH
      default: break;
    }
  }
  if (plan.d0 * plan.d1 * plan.d2 * static_cast<int64_t>(sizeof(T)) >= kJitMinBytes) {
    void* params[] = {const_cast<DevProgram<T>*>(&dp), const_cast<DevOperands<NI>*>(&ops), const_cast<stream::Plan*>(&plan)};
    const int rc = jit::launch_stream_specialised(TypeTag<T>::dtype, NI, 3, dp.op, dp.src, dp.arg, dp.n_ops, params,
                                                  static_cast<int>(blocks), stream::kDirectThreads, 0, stream_);
    if (rc != 0) return rc < 0 ? rc : SP_OK;
  }
  return kNotLaunched;
}

static int g_reduce_kernel = -1;  // 0 = ring, 1 = direct; env SPARTAN_REDUCE_KERNEL=ring|direct, default direct
static bool use_direct_reduce() {
  if (g_reduce_kernel < 0) {
    const char* e = getenv("SPARTAN_REDUCE_KERNEL");
    g_reduce_kernel = (e && e[0] == 'r') ? 0 : 1;
  }
  return g_reduce_kernel == 1;
}

template <typename T, int NI, typename PROG>
static int launch_direct_reduce_as(const DevProgram<T>& dp, const DevOperands<NI>& ops, const stream::Plan& plan, int red_op,
                                   T* scratch, int64_t blocks, cudaStream_t stream_) {
  stream::direct_reduce_kernel<T, NI, PROG><<<static_cast<unsigned>(blocks), stream::kDirectThreads, 0, stream_>>>(
      dp, ops, plan, red_op, scratch);
  SP_CUDA_CHECK(cudaGetLastError());
  return SP_OK;
}

// Leading-axis reduction on the direct kernel.  `plan` must have been made with rc = stream::kDirectRows (one partial row
// per chunk, like MODE 1).  SP_OK when launched, kNotLaunched without a compiled instance of the program.
template <typename T, int NI>
static int launch_direct_reduce(const DevProgram<T>& dp, const DevOperands<NI>& ops, const stream::Plan& plan, int red_op,
                                T* scratch, cudaStream_t stream_) {
  constexpr int VEC = 16 / sizeof(T);
  if (plan.d2 % VEC != 0 || plan.rc != stream::kDirectRows) return kNotLaunched;
  for (int i = 0; i < ops.n_in; ++i)
    if (ops.in[i].kind == kGeneric) return kNotLaunched;
  const int64_t row_vecs = plan.d2 / VEC;
  const int64_t blocks = plan.d0 * plan.n_chunks * ((row_vecs + stream::kDirectThreads - 1) / stream::kDirectThreads);
  if (blocks <= 0 || blocks >= (1ll << 31)) return kNotLaunched;
  if constexpr (NI == 2 && !std::is_same<T, long long>::value) {
    switch (match_static<T>(dp)) {
#define SP_LAUNCH(IDX, TYPE) case IDX: return launch_direct_reduce_as<T, NI, TYPE>(dp, ops, plan, red_op, scratch, blocks, stream_);
      SP_STATIC_PROGRAMS(SP_LAUNCH)
#undef SP_LAUNCH
      default: break;
    }
  }
  if (plan.d0 * plan.d1 * plan.d2 * static_cast<int64_t>(sizeof(T)) >= kJitMinBytes) {
    int red = red_op;
    void* params[] = {const_cast<DevProgram<T>*>(&dp), const_cast<DevOperands<NI>*>(&ops), const_cast<stream::Plan*>(&plan),
                      &red, &scratch};
    const int rc = jit::launch_stream_specialised(TypeTag<T>::dtype, NI, 4, dp.op, dp.src, dp.arg, dp.n_ops, params,
                                                  static_cast<int>(blocks), stream::kDirectThreads, 0, stream_);
    if (rc != 0) return rc < 0 ? rc : SP_OK;
  }
  return kNotLaunched;
}

#ifndef SP_FLAT_ROW_BYTES
#define SP_FLAT_ROW_BYTES 16384
#endif
// A flat contiguous map (d0 = d1 = 1) is re-viewed as rows of `kFlatRow` elements so the ring carries full stages.
template <typename T, int NI>
static bool reshape_flat(DevOperands<NI>& ops, int64_t dims[3]) {
  constexpr int64_t kFlatRow = (SP_FLAT_ROW_BYTES) / sizeof(T);
  if (dims[0] * dims[1] != 1 || dims[2] < kFlatRow * 8 || dims[2] % kFlatRow != 0) return false;
  for (int i = 0; i < ops.n_in; ++i)
    if (ops.in[i].stride[2] != 0 && ops.in[i].stride[2] != 1) return false;
  if (ops.out.stride[2] != 1) return false;
  dims[1] = dims[2] / kFlatRow;
  dims[2] = kFlatRow;
  for (int i = 0; i < ops.n_in; ++i) ops.in[i].stride[1] = ops.in[i].stride[2] * kFlatRow;
  ops.out.stride[1] = kFlatRow;
  return true;
}

template <typename T, int V, int NI>
static int launch_map_v(const sp_program* prog, int n_in, const sp_operand* in, const sp_operand* out,
                        const int64_t dims[3], cudaStream_t stream) {
  DevProgram<T> dp;
  SP_REQUIRE(convert_program<T>(prog, &dp), SP_ERR_UNSUPPORTED,
             "expression needs more than %d temporaries; split it (the fusion pass does)", kMaxTmp);
  DevOperands<NI> ops;
  memset(&ops, 0, sizeof(ops));
  ops.n_in = n_in;
  for (int i = 0; i < n_in; ++i) ops.in[i] = make_operand<T, V>(in[i], 2, dims);
  ops.out = make_operand<T, V>(*out, 2, dims);
  if (ops.out.kind == kSplat) ops.out.kind = kGeneric;
  if (dims[0] * dims[1] * dims[2] == 0) return SP_OK;
  {
    // The streaming kernel works on 32-byte-per-lane vectors.  With up to two operands any program may run on it
    // (interpreted if need be); with more operands the interpreter's register footprint would spill, so only a
    // compiled instance is used there.
    const bool interp_ok = (V == 32 / sizeof(T));
    int64_t sd[3] = {dims[0], dims[1], dims[2]};
    DevOperands<NI> sops = ops;
    DevProgram<T> sdp = dp;
    if (reshape_flat<T, NI>(sops, sd)) sdp.index_stride[1] = sd[2] * sdp.index_stride[2];   // i2 = i1' * row + i2'
    stream::Plan plan;
    if (plan_stream<T, NI>(sops, sd, true, 0, &plan)) {
      if (use_direct_map()) {
        const int rc = launch_direct<T, NI>(sdp, sops, plan, stream);
        if (rc != kNotLaunched) return rc;
      }
      const int rc = launch_stream<T, NI, 0>(sdp, sops, plan, 0, nullptr, stream, interp_ok);
      if (rc != kNotLaunched) return rc;
    }
  }
  Dims3 d{dims[0], dims[1], dims[2]};
  const int64_t d2v = (dims[2] + V - 1) / V;
  const int64_t total = dims[0] * dims[1] * d2v;
  const int64_t want = (total + 255) / 256;
  const int blocks = static_cast<int>(std::min<int64_t>(want, static_cast<int64_t>(num_sms()) * 16));
  map_kernel<T, V, NI><<<blocks, 256, 0, stream>>>(dp, ops, d, d2v, total);
  SP_CUDA_CHECK(cudaGetLastError());
  return SP_OK;
}

template <typename T>
int launch_map(const sp_program* prog, int n_in, const sp_operand* in, const sp_operand* out,
               const int64_t dims[3], cudaStream_t stream) {
  if (n_in <= 2) return launch_map_v<T, TypeTag<T>::VA, 2>(prog, n_in, in, out, dims, stream);
  return launch_map_v<T, TypeTag<T>::VB, 8>(prog, n_in, in, out, dims, stream);
}

struct ReducePlan {
  bool row;        // d2 == 1: vectors along the reduced axis
  int splits;      // partials per output element
  int block;       // threads per CTA (row kernel)
  int64_t blocks_x;
};

template <int V>
static ReducePlan plan_reduce(const int64_t dims[3]) {
  ReducePlan p;
  const int64_t target = static_cast<int64_t>(num_sms()) * 8;
  p.row = (dims[2] == 1);
  if (p.row) {
    const int64_t nvec = (dims[1] + V - 1) / V;
    int block = 32;
    while (block < 256 && block < nvec) block <<= 1;
    p.block = block;
    int64_t segs = 1;
    if (dims[0] < target) {
      segs = std::min<int64_t>((target + dims[0] - 1) / dims[0], std::max<int64_t>(1, nvec / (block * 4)));
      segs = std::max<int64_t>(segs, 1);
    }
    p.splits = static_cast<int>(std::min<int64_t>(segs, 4096));
    p.blocks_x = dims[0] * p.splits;
  } else {
    const int64_t d2v = (dims[2] + V - 1) / V;
    p.blocks_x = (dims[0] * d2v + 31) / 32;
    p.block = 256;
    int64_t s = 1;
    if (p.blocks_x < target) s = (target + p.blocks_x - 1) / p.blocks_x;
    const int64_t max_s = std::max<int64_t>(1, (dims[1] + 31) / 32);   // >= 4 rows per thread
    p.splits = static_cast<int>(std::min<int64_t>(std::min<int64_t>(s, max_s), 1024));
  }
  return p;
}

template <typename T, int V>
static int64_t reduce_scratch_elems(const int64_t dims[3]) {
  const ReducePlan p = plan_reduce<V>(dims);
  const int64_t simple = static_cast<int64_t>(p.splits) * dims[0] * dims[2];
  const int64_t streamed = stream_reduce_chunks<T>(dims) * dims[0] * dims[2];
  // trailing-axis reductions on the streaming kernel (MODE 2): one partial per (row, 4 KiB panel of the reduced axis)
  const int64_t row_panels = (dims[1] * static_cast<int64_t>(sizeof(T)) + stream::kPanelBytes - 1) / stream::kPanelBytes;
  const int64_t row_streamed = p.row ? row_panels * dims[0] : 0;
  return std::max(std::max(simple, streamed), row_streamed);
}

template <typename T, int V, int NI>
static int launch_reduce_v(const sp_program* prog, int n_in, const sp_operand* in, const sp_operand* out,
                           const int64_t dims[3], int red_op, int accumulate, void* scratch, int64_t scratch_bytes,
                           cudaStream_t stream) {
  const ReducePlan p = plan_reduce<V>(dims);
  const int64_t need = static_cast<int64_t>(p.splits) * dims[0] * dims[2] * static_cast<int64_t>(sizeof(T));
  SP_REQUIRE(scratch != nullptr && scratch_bytes >= need, SP_ERR_INVALID,
             "sp_map_reduce: scratch %lld B < required %lld B", (long long)scratch_bytes, (long long)need);
  DevProgram<T> dp;
  SP_REQUIRE(convert_program<T>(prog, &dp), SP_ERR_UNSUPPORTED,
             "expression needs more than %d temporaries; split it (the fusion pass does)", kMaxTmp);
  DevOperands<NI> ops;
  memset(&ops, 0, sizeof(ops));
  ops.n_in = n_in;
  const int vec_axis = p.row ? 1 : 2;
  for (int i = 0; i < n_in; ++i) ops.in[i] = make_operand<T, V>(in[i], vec_axis, dims);
  Dims3 d{dims[0], dims[1], dims[2]};
  DevOperand o;
  o.ptr = out->ptr; o.dtype = out->dtype; o.kind = kGeneric;
  o.stride[0] = out->stride[0]; o.stride[1] = 0; o.stride[2] = out->stride[2];
  const int64_t n_out = dims[0] * dims[2];
  if (n_out == 0) return SP_OK;
  T* sc = static_cast<T*>(scratch);
  if (!p.row) {
    const bool interp_ok = (V == 32 / sizeof(T));      // see launch_map_v
    stream::Plan plan;
    if (plan_stream<T, NI>(ops, dims, false, kStreamReduceRows, &plan)) {
      const int64_t sneed = plan.n_chunks * dims[0] * dims[2] * static_cast<int64_t>(sizeof(T));
      SP_REQUIRE(scratch_bytes >= sneed, SP_ERR_INVALID, "sp_map_reduce: scratch %lld B < required %lld B",
                 (long long)scratch_bytes, (long long)sneed);
      int rc = use_direct_reduce() ? launch_direct_reduce<T, NI>(dp, ops, plan, red_op, sc, stream) : kNotLaunched;
      if (rc == kNotLaunched) rc = launch_stream<T, NI, 1>(dp, ops, plan, red_op, sc, stream, interp_ok);
      if (rc < 0) return rc;
      if (rc != kNotLaunched) {
        const int fb = static_cast<int>(std::min<int64_t>((n_out + 255) / 256, 4096));
        finalize_kernel<T><<<fb, 256, 0, stream>>>(sc, n_out, static_cast<int>(plan.n_chunks), n_out, 1, o, dims[2],
                                                   red_op, accumulate);
        SP_CUDA_CHECK(cudaGetLastError());
        return SP_OK;
      }
    }
  }
  if (p.row) {
    // Long contiguous rows: the streaming kernel with the reduced axis as its vector axis (MODE 2).  The iteration space
    // (rows, cols, 1) becomes (1, rows, cols); operand and index strides move one slot to the right.
    const bool interp_ok = (V == 32 / sizeof(T));
    int64_t sd[3] = {1, dims[0], dims[1]};
    DevOperands<NI> sops = ops;
    for (int i = 0; i < n_in; ++i) {
      sops.in[i].stride[2] = ops.in[i].stride[1];
      sops.in[i].stride[1] = ops.in[i].stride[0];
      sops.in[i].stride[0] = 0;
    }
    DevProgram<T> sdp = dp;
    sdp.index_stride[2] = dp.index_stride[1];
    sdp.index_stride[1] = dp.index_stride[0];
    sdp.index_stride[0] = 0;
    stream::Plan plan;
    // MODE 2 units are rb rows over their whole length (one stage per panel); it needs one item per consumer warp per
    // stage (rb * 4 <= 16) and rows long enough to amortise the end-of-row fold
    if (dims[1] * static_cast<int64_t>(sizeof(T)) >= 4 * stream::kPanelBytes && plan_stream<T, NI>(sops, sd, false, 0, &plan)) {
      if (plan.rb * stream::kSegsPerPanel > stream::kConsumerWarps) {      // one operand: 4 rows per stage, deeper ring
        plan.rb = stream::kConsumerWarps / stream::kSegsPerPanel;
        plan.stage_bytes = plan.n_stream * plan.rb * stream::kPanelBytes;
        plan.n_stages = std::min(stream::kMaxStages, stream::kRingBytes / plan.stage_bytes);
      }
      plan.rc = plan.rb;
      plan.n_chunks = (sd[1] + plan.rb - 1) / plan.rb;
      plan.n_units = plan.n_chunks;
      const int64_t sneed = dims[0] * static_cast<int64_t>(sizeof(T));
      SP_REQUIRE(scratch_bytes >= sneed, SP_ERR_INVALID, "sp_map_reduce: scratch %lld B < required %lld B",
                 (long long)scratch_bytes, (long long)sneed);
      int rc = launch_stream<T, NI, 2>(sdp, sops, plan, red_op, sc, stream, interp_ok);
      if (rc < 0) return rc;
      if (rc != kNotLaunched) {
        const int fb = static_cast<int>(std::min<int64_t>((n_out + 255) / 256, 4096));
        finalize_kernel<T><<<fb, 256, 0, stream>>>(sc, n_out, 1, n_out, 1, o, 1, red_op, accumulate);
        SP_CUDA_CHECK(cudaGetLastError());
        return SP_OK;
      }
    }
    SP_REQUIRE(p.blocks_x < (1ll << 31), SP_ERR_INVALID, "sp_map_reduce: too many rows (%lld)", (long long)dims[0]);
    const int64_t nvec = (dims[1] + V - 1) / V;
    reduce_row_kernel<T, V, NI><<<static_cast<unsigned>(p.blocks_x), p.block, 0, stream>>>(dp, ops, d, nvec, p.splits,
                                                                                         red_op, sc);
    SP_CUDA_CHECK(cudaGetLastError());
    const int fb = static_cast<int>(std::min<int64_t>((n_out + 255) / 256, 4096));
    finalize_kernel<T><<<fb, 256, 0, stream>>>(sc, n_out, p.splits, 1, p.splits, o, 1, red_op, accumulate);
  } else {
    SP_REQUIRE(p.blocks_x < (1ll << 31), SP_ERR_INVALID, "sp_map_reduce: output too large");
    const int64_t d2v = (dims[2] + V - 1) / V;
    dim3 grid(static_cast<unsigned>(p.blocks_x), static_cast<unsigned>(p.splits));
    dim3 block(32, 8);
    reduce_col_kernel<T, V, NI><<<grid, block, 0, stream>>>(dp, ops, d, d2v, red_op, sc);
    SP_CUDA_CHECK(cudaGetLastError());
    const int fb = static_cast<int>(std::min<int64_t>((n_out + 255) / 256, 4096));
    finalize_kernel<T><<<fb, 256, 0, stream>>>(sc, n_out, p.splits, n_out, 1, o, dims[2], red_op, accumulate);
  }
  SP_CUDA_CHECK(cudaGetLastError());
  return SP_OK;
}

template <typename T>
int launch_reduce(const sp_program* prog, int n_in, const sp_operand* in, const sp_operand* out, const int64_t dims[3],
                  int red_op, int accumulate, void* scratch, int64_t scratch_bytes, cudaStream_t stream) {
  if (n_in <= 2)
    return launch_reduce_v<T, TypeTag<T>::VA, 2>(prog, n_in, in, out, dims, red_op, accumulate, scratch, scratch_bytes, stream);
  return launch_reduce_v<T, TypeTag<T>::VB, 8>(prog, n_in, in, out, dims, red_op, accumulate, scratch, scratch_bytes, stream);
}

template <typename T>
int64_t reduce_scratch_bytes(const int64_t dims[3]) {
  // upper bound over both operand-capacity variants
  const int64_t a = reduce_scratch_elems<T, TypeTag<T>::VA>(dims);
  const int64_t b = reduce_scratch_elems<T, TypeTag<T>::VB>(dims);
  return std::max(a, b) * static_cast<int64_t>(sizeof(T)) + 256;
}

}  // namespace sp
