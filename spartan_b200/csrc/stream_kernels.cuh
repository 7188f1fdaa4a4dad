// Streaming fast path of the fused map and map+reduce(axis = leading) kernels.
//
// The simple kernels in map_reduce_impl.cuh keep only ~32 KB of loads in flight per SM (the
// interpreter's register footprint caps occupancy), which is less than half of what HBM3e needs.
// Here the loads are decoupled from the interpreter: one producer warp issues 1 KiB
// cp.async.bulk copies (TMA engine, SASS UBLKCP) into a 192 KiB shared-memory ring (3-6 stages of 32-64 KiB,
// sized so that a stage carries 16 warp-items whatever the operand count) guarded by mbarriers, sixteen consumer warps read the ring with conflict-free 16-byte LDS, run the bytecode
// and either store the result (map) or fold it into per-thread accumulators (reduce).  Up to
// 192 KiB per SM is in flight regardless of register pressure.  One CTA per SM, persistent over
// work units.
//
// Work unit = RC consecutive rows x one 4 KiB-wide column panel; one bulk copy moves one row of a panel
// (measured: with 1 KiB copies the ring stayed empty 58% of the time at 3.9 TB/s, profiles/).  A consumer
// warp handles one 1 KiB segment of a panel row at a time; a segment is
// 1024 B = 32 lanes x 2 x 16 B: lane L owns bytes [16L, 16L+16) and [512+16L, 512+16L+16), i.e.
// V = 32 / sizeof(T) elements in two contiguous halves (both LDS.128 and STG.128 stay conflict-free
// and coalesced).  Reduction partials are written per unit and combined in fixed order by
// finalize_kernel (deterministic).
//
// Eligibility (checked on the host): every array operand is either streamable (compute dtype, unit
// stride, 16-byte aligned rows) or constant along the vector axis; row length >= 1 KiB.
#pragma once
#include "interp.cuh"

namespace sp {
namespace stream {

constexpr int kRingBytes = 192 * 1024;         // shared-memory ring; a stage holds rb rows x one panel of every streamed operand
constexpr int kMaxStages = 8;
constexpr int kMaxStageBytes = 64 * 1024;
constexpr int kSegBytes = 1024;               // what one consumer warp handles per pass (32 lanes x 32 B)
constexpr int kPanelBytes = 4096;             // one bulk copy = one row x one panel (1 KiB copies starve the TMA unit)
constexpr int kSegsPerPanel = kPanelBytes / kSegBytes;
constexpr int kConsumerWarps = 16;
constexpr int kThreads = 32 * (1 + kConsumerWarps);
constexpr int kRedBytes = kConsumerWarps * 32 * 32;   // [warps][32 lanes][V * sizeof(T) = 32 B]
constexpr int kSmemBytes = kRingBytes + 1024 /*align*/ + kRedBytes + 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "SWAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni SWAIT_DONE;\n\t"
      "bra.uni SWAIT_LOOP;\n\t"
      "SWAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
// 1-D bulk async copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(32 * kConsumerWarps) : "memory"); }

struct Plan {
  int64_t d0, d1, d2;        // iteration space; vectors run along d2; reduce folds d1
  int n_panels;              // ceil(d2 * sizeof(T) / kPanelBytes)
  int rc;                    // rows per unit
  int64_t n_chunks;          // ceil(d1 / rc)      (reduce: per d0)    map: rows = d0*d1 flattened is NOT assumed
  int64_t n_units;           // d0 * n_chunks * n_panels
  int n_stream;              // streamed operands
  int rb;                    // rows per stage  (n_stream * rb * kPanelBytes = stage_bytes <= kMaxStageBytes)
  int stage_bytes;
  int n_stages;              // kRingBytes / stage_bytes, at most kMaxStages
  int stream_slot[SP_MAX_OPERANDS];   // operand -> slot in the stage, or -1 (loaded directly)
};

// Loads the two halves of a non-streamed operand (constant or generic along the vector axis).
template <typename T, int V>
__device__ __forceinline__ void load_direct_split(const DevOperand& o, int64_t base, int64_t col_a, int64_t col_b,
                                                  int valid_a, int valid_b, T (&r)[V]) {
  constexpr int H = V / 2;
  if (o.stride[2] == 0) {
    const T x = load_as<T>(o.ptr, o.dtype, base);
#pragma unroll
    for (int v = 0; v < V; ++v) r[v] = x;
    return;
  }
#pragma unroll
  for (int v = 0; v < H; ++v) {
    r[v] = (v < valid_a) ? load_as<T>(o.ptr, o.dtype, base + (col_a + v) * o.stride[2]) : T(0);
    r[H + v] = (v < valid_b) ? load_as<T>(o.ptr, o.dtype, base + (col_b + v) * o.stride[2]) : T(0);
  }
}

// MODE 2 (trailing-axis reductions): reduce over d2, the vector axis.  A work unit is `rb` consecutive rows over their WHOLE
// length: stage p of a unit carries panel p of those rows, so consumer warp w always works on the same (row w / 4, segment
// w % 4) and keeps its V-wide accumulator in registers across the panels of the row -- the per-element cost of MODE 1 --
// and only at the end of the unit are the lanes folded by shuffle and the row's four segment partials combined through
// shared memory (fixed order).  One value per row goes to scratch[d0 * d1]; nothing is left for finalize to fold.
template <typename T, int NI, typename PROG>
__device__ __forceinline__ void row_reduce_body(const DevProgram<T>& prog, const DevOperands<NI>& ops, const Plan& plan,
                                                const int red_op, T* __restrict__ scratch, const uint32_t smem_base,
                                                uint8_t* smem_gen, T* red_smem, const uint32_t bar_base) {
  constexpr int V = 32 / sizeof(T);
  constexpr int H = V / 2;
  constexpr int EPS = kSegBytes / sizeof(T);
  constexpr int EPP = kPanelBytes / sizeof(T);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  const int n_stages = plan.n_stages;
  const int stage_bytes = plan.stage_bytes;
  const int rb = plan.rb;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t n_units = plan.d0 * plan.n_chunks;           // n_chunks = ceil(d1 / rb)
  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    const int slot = lane / rb;
    const int r_in = lane - slot * rb;
    int my_op = -1;
    for (int i = 0; i < NI; ++i)
      if (i < ops.n_in && plan.stream_slot[i] == slot) my_op = i;
    for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x) {
      const int64_t chunk = u % plan.n_chunks;
      const int64_t i0 = u / plan.n_chunks;
      const int64_t row0 = chunk * rb;
      const int rows = static_cast<int>(min(static_cast<int64_t>(rb), plan.d1 - row0));
      for (int panel = 0; panel < plan.n_panels; ++panel) {
        const int64_t col0 = static_cast<int64_t>(panel) * EPP;
        const uint32_t seg_bytes = static_cast<uint32_t>(min(static_cast<int64_t>(EPP), plan.d2 - col0) * sizeof(T));
        mbar_wait(empty_bar(stage), phase ^ 1);
        if (lane == 0) mbar_expect_tx(full_bar(stage), seg_bytes * rows * plan.n_stream);
        __syncwarp();
        if (my_op >= 0 && slot < plan.n_stream && r_in < rows) {
          const DevOperand& o = ops.in[my_op];
          const T* src = static_cast<const T*>(o.ptr) + i0 * o.stride[0] + (row0 + r_in) * o.stride[1] + col0;
          bulk_g2s(smem_base + stage * stage_bytes + (slot * rb + r_in) * kPanelBytes, src, seg_bytes, full_bar(stage));
        }
        if (++stage == n_stages) { stage = 0; phase ^= 1; }
      }
    }
    return;
  }
  const int cw = warp - 1;
  int stage = 0;
  uint32_t phase = 0;
  for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x) {
    const int64_t chunk = u % plan.n_chunks;
    const int64_t i0 = u / plan.n_chunks;
    const int64_t row0 = chunk * rb;
    const int rows = static_cast<int>(min(static_cast<int64_t>(rb), plan.d1 - row0));
    T acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = red_identity<T>(red_op);
    for (int panel = 0; panel < plan.n_panels; ++panel) {
      const int64_t col0 = static_cast<int64_t>(panel) * EPP;
      const int64_t cols = min(static_cast<int64_t>(EPP), plan.d2 - col0);
      mbar_wait(full_bar(stage), phase);
      const uint8_t* sbase = smem_gen + stage * stage_bytes;
      bool released = false;
      for (int item = cw; item < rows * kSegsPerPanel; item += kConsumerWarps) {
        const int r = item / kSegsPerPanel;
        const int q = item % kSegsPerPanel;
        const int64_t col_a = col0 + q * EPS + lane * H;
        const int64_t col_b = col_a + EPS / 2;
        const int valid_a = static_cast<int>(max(static_cast<int64_t>(0), min(static_cast<int64_t>(H), col0 + cols - col_a)));
        const int valid_b = static_cast<int>(max(static_cast<int64_t>(0), min(static_cast<int64_t>(H), col0 + cols - col_b)));
        T in[NI][V];
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          if (i < ops.n_in) {
            const int sl = plan.stream_slot[i];
            if (sl >= 0) {
              const uint8_t* seg = sbase + (sl * rb + r) * kPanelBytes + q * kSegBytes;
              T tmp[V];
              *reinterpret_cast<int4*>(&tmp[0]) = *reinterpret_cast<const int4*>(seg + 16 * lane);
              *reinterpret_cast<int4*>(&tmp[H]) = *reinterpret_cast<const int4*>(seg + 512 + 16 * lane);
#pragma unroll
              for (int v = 0; v < V; ++v) in[i][v] = tmp[v];
            } else {
              const DevOperand& o = ops.in[i];
              load_direct_split<T, V>(o, i0 * o.stride[0] + (row0 + r) * o.stride[1], col_a, col_b, valid_a, valid_b, in[i]);
            }
          }
        }
        if (item + kConsumerWarps >= rows * kSegsPerPanel) {
          __syncwarp();
          if (lane == 0) mbar_arrive(empty_bar(stage));
          released = true;
        }
        T res[V], idx[V];
        if (prog.uses_index) {
          const long long base = prog.index_base + i0 * prog.index_stride[0] + (row0 + r) * prog.index_stride[1];
#pragma unroll
          for (int v = 0; v < H; ++v) {
            idx[v] = static_cast<T>(base + (col_a + v) * prog.index_stride[2]);
            idx[H + v] = static_cast<T>(base + (col_b + v) * prog.index_stride[2]);
          }
        }
        PROG::template run<T, V, NI>(prog, in, idx, res);
        if (valid_a == H && valid_b == H) {
#pragma unroll
          for (int v = 0; v < V; ++v) acc[v] = red_apply<T>(red_op, acc[v], res[v]);
        } else {                       // the ragged end of a row: elements past it hold no data
#pragma unroll
          for (int v = 0; v < H; ++v) {
            if (v < valid_a) acc[v] = red_apply<T>(red_op, acc[v], res[v]);
            if (v < valid_b) acc[H + v] = red_apply<T>(red_op, acc[H + v], res[H + v]);
          }
        }
      }
      if (!released) {
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar(stage));
      }
      if (++stage == n_stages) { stage = 0; phase ^= 1; }
    }
    // end of the unit: lanes -> one value per (row, segment); segments -> one value per row
    // (rows * 4 <= 16 items per stage: warp w owned item w of every stage, i.e. row w / 4, segment w % 4)
    T a = acc[0];
#pragma unroll
    for (int v = 1; v < V; ++v) a = red_apply<T>(red_op, a, acc[v]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) a = red_apply<T>(red_op, a, shfl_down_t<T>(a, d));
    if (lane == 0) red_smem[cw] = a;
    consumer_bar();
    if (cw == 0 && lane < rows) {
      T r = red_smem[lane * kSegsPerPanel];
#pragma unroll
      for (int sg = 1; sg < kSegsPerPanel; ++sg) r = red_apply<T>(red_op, r, red_smem[lane * kSegsPerPanel + sg]);
      scratch[i0 * plan.d1 + row0 + lane] = r;
    }
    consumer_bar();
  }
}

// MODE 0: map (store), MODE 1: reduce over d1 (partials to scratch[chunk][d0][d2]), MODE 2: see row_reduce_body.
template <typename T, int NI, int MODE, typename PROG>
__global__ void __launch_bounds__(kThreads, 1)
stream_kernel(const DevProgram<T> prog, const DevOperands<NI> ops, const Plan plan, const int red_op,
              T* __restrict__ scratch) {
  constexpr int V = 32 / sizeof(T);
  constexpr int H = V / 2;
  constexpr int EPS = kSegBytes / sizeof(T);      // elements per 1 KiB segment
  constexpr int EPP = kPanelBytes / sizeof(T);    // elements per panel row
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + kRingBytes + kRedBytes;
  T* red_smem = reinterpret_cast<T*>(smem_gen + kRingBytes);   // [warps][32 lanes][V]
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  const int n_stages = plan.n_stages;
  const int stage_bytes = plan.stage_bytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), kConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if constexpr (MODE == 2) {
    row_reduce_body<T, NI, PROG>(prog, ops, plan, red_op, scratch, smem_base, smem_gen, red_smem, bar_base);
    return;
  }
  const int rb = plan.rb;
  const int stages_per_unit = (plan.rc + rb - 1) / rb;

  if (warp == 0) {
    // ===================== producer: one bulk copy per (operand, row) per lane =====================
    int stage = 0;
    uint32_t phase = 0;
    const int slot = lane / rb;            // which streamed operand this lane copies
    const int r_in = lane - slot * rb;     // which row of the stage
    int my_op = -1;
    for (int i = 0; i < NI; ++i)
      if (i < ops.n_in && plan.stream_slot[i] == slot) my_op = i;
    for (int64_t u = blockIdx.x; u < plan.n_units; u += gridDim.x) {
      const int panel = static_cast<int>(u % plan.n_panels);
      const int64_t t = u / plan.n_panels;
      const int64_t chunk = t % plan.n_chunks;
      const int64_t i0 = t / plan.n_chunks;
      const int64_t col0 = static_cast<int64_t>(panel) * EPP;
      const int64_t cols = min(static_cast<int64_t>(EPP), plan.d2 - col0);
      const uint32_t seg_bytes = static_cast<uint32_t>(cols * sizeof(T));
      const int64_t row_end = min(plan.d1, (chunk + 1) * plan.rc);
      for (int st = 0; st < stages_per_unit; ++st) {
        const int64_t row0 = chunk * plan.rc + static_cast<int64_t>(st) * rb;
        if (row0 >= row_end) break;
        const int rows = static_cast<int>(min(static_cast<int64_t>(rb), row_end - row0));
        mbar_wait(empty_bar(stage), phase ^ 1);
        if (lane == 0) mbar_expect_tx(full_bar(stage), seg_bytes * rows * plan.n_stream);
        __syncwarp();
        if (my_op >= 0 && slot < plan.n_stream && r_in < rows) {
          const DevOperand& o = ops.in[my_op];
          const T* src = static_cast<const T*>(o.ptr) + i0 * o.stride[0] + (row0 + r_in) * o.stride[1] + col0;
          const uint32_t dst = smem_base + stage * stage_bytes + (slot * rb + r_in) * kPanelBytes;
          bulk_g2s(dst, src, seg_bytes, full_bar(stage));
        }
        if (++stage == n_stages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================== consumers =====================
    const int cw = warp - 1;
    int stage = 0;
    uint32_t phase = 0;
    for (int64_t u = blockIdx.x; u < plan.n_units; u += gridDim.x) {
      const int panel = static_cast<int>(u % plan.n_panels);
      const int64_t t = u / plan.n_panels;
      const int64_t chunk = t % plan.n_chunks;
      const int64_t i0 = t / plan.n_chunks;
      const int64_t col0 = static_cast<int64_t>(panel) * EPP;
      const int64_t cols = min(static_cast<int64_t>(EPP), plan.d2 - col0);
      const int q = cw % kSegsPerPanel;                      // the 1 KiB segment of the panel this warp owns
      const int64_t col_a = col0 + q * EPS + lane * H;       // first half of this lane's vector
      const int64_t col_b = col_a + EPS / 2;                 // second half
      const int valid_a = static_cast<int>(max(static_cast<int64_t>(0), min(static_cast<int64_t>(H), col0 + cols - col_a)));
      const int valid_b = static_cast<int>(max(static_cast<int64_t>(0), min(static_cast<int64_t>(H), col0 + cols - col_b)));
      const int64_t row_end = min(plan.d1, (chunk + 1) * plan.rc);
      T acc[V];
      if (MODE == 1) {
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v] = red_identity<T>(red_op);
      }
      for (int st = 0; st < stages_per_unit; ++st) {
        const int64_t row0 = chunk * plan.rc + static_cast<int64_t>(st) * rb;
        if (row0 >= row_end) break;
        const int rows = static_cast<int>(min(static_cast<int64_t>(rb), row_end - row0));
        mbar_wait(full_bar(stage), phase);
        const uint8_t* sbase = smem_gen + stage * stage_bytes;
        bool released = false;
        for (int item = cw; item < rows * kSegsPerPanel; item += kConsumerWarps) {
          const int r = item / kSegsPerPanel;                // item % kSegsPerPanel == q for every item of this warp
          T in[NI][V];
#pragma unroll
          for (int i = 0; i < NI; ++i) {
            if (i < ops.n_in) {
              const int sl = plan.stream_slot[i];
              if (sl >= 0) {
                const uint8_t* seg = sbase + (sl * rb + r) * kPanelBytes + q * kSegBytes;
                const int4 a = *reinterpret_cast<const int4*>(seg + 16 * lane);
                const int4 b = *reinterpret_cast<const int4*>(seg + 512 + 16 * lane);
                T tmp[V];
                *reinterpret_cast<int4*>(&tmp[0]) = a;
                *reinterpret_cast<int4*>(&tmp[H]) = b;
#pragma unroll
                for (int v = 0; v < V; ++v) in[i][v] = tmp[v];
              } else {
                const DevOperand& o = ops.in[i];
                load_direct_split<T, V>(o, i0 * o.stride[0] + (row0 + r) * o.stride[1], col_a, col_b, valid_a, valid_b,
                                        in[i]);
              }
            }
          }
          // The operands of this warp's last item of the stage are in registers: hand the stage back to the producer
          // now, so the refill is in flight while the program runs and the results are stored.
          if (item + kConsumerWarps >= rows * kSegsPerPanel) {
            __syncwarp();
            if (lane == 0) mbar_arrive(empty_bar(stage));
            released = true;
          }
          T res[V], idx[V];
          if (prog.uses_index) {       // two halves of the lane's vector sit EPS/2 elements apart
            const long long base = prog.index_base + i0 * prog.index_stride[0] + (row0 + r) * prog.index_stride[1];
#pragma unroll
            for (int v = 0; v < H; ++v) {
              idx[v] = static_cast<T>(base + (col_a + v) * prog.index_stride[2]);
              idx[H + v] = static_cast<T>(base + (col_b + v) * prog.index_stride[2]);
            }
          }
          PROG::template run<T, V, NI>(prog, in, idx, res);
          if (MODE == 0) {
            const DevOperand& o = ops.out;
            T* dst = static_cast<T*>(const_cast<void*>(o.ptr)) + i0 * o.stride[0] + (row0 + r) * o.stride[1];
            if (valid_a == H) *reinterpret_cast<int4*>(dst + col_a) = *reinterpret_cast<const int4*>(&res[0]);
            else
              for (int v = 0; v < valid_a; ++v) dst[col_a + v] = res[v];
            if (valid_b == H) *reinterpret_cast<int4*>(dst + col_b) = *reinterpret_cast<const int4*>(&res[H]);
            else
              for (int v = 0; v < valid_b; ++v) dst[col_b + v] = res[H + v];
          } else {
#pragma unroll
            for (int v = 0; v < V; ++v) acc[v] = red_apply<T>(red_op, acc[v], res[v]);
          }
        }
        if (!released) {               // a warp without an item in this stage
          __syncwarp();
          if (lane == 0) mbar_arrive(empty_bar(stage));
        }
        if (++stage == n_stages) { stage = 0; phase ^= 1; }
      }
      if (MODE == 1) {
        // fold the 8 consumer warps (fixed order), warp 0 of the consumers writes the unit's partial
#pragma unroll
        for (int v = 0; v < V; ++v) red_smem[(cw * 32 + lane) * V + v] = acc[v];
        consumer_bar();
        if (cw < kSegsPerPanel) {       // warp q folds the warps that worked on segment q, in fixed order
#pragma unroll
          for (int w = cw + kSegsPerPanel; w < kConsumerWarps; w += kSegsPerPanel)
#pragma unroll
            for (int v = 0; v < V; ++v) acc[v] = red_apply<T>(red_op, acc[v], red_smem[(w * 32 + lane) * V + v]);
          T* dst = scratch + (chunk * plan.d0 + i0) * plan.d2;
          for (int v = 0; v < valid_a; ++v) dst[col_a + v] = acc[v];
          for (int v = 0; v < valid_b; ++v) dst[col_b + v] = acc[H + v];
        }
        consumer_bar();
      }
    }
  }
}

// ------------------------------------------------------------------------------------ direct map
// The pure map writes as many bytes as it reads per operand, and with stores in the mix the bulk-copy ring above loses
// ~10% to its consumers waiting on late stages (ncu: 43% of samples at the full barrier).  For compiled programs (static
// catalogue or run-time specialised -- their register footprint is small) the map therefore goes the plain way: a
// non-persistent grid, every thread issues kDirectUnroll independent 16-byte streaming loads per operand (coalesced: the
// vectors of one step of a block are contiguous), runs the program on them and stores with streaming 16-byte stores.
// A CTA covers 256 x kDirectUnroll vectors (16 KiB) of one row; memory-level parallelism comes from 8 resident CTAs per SM.
constexpr int kDirectThreads = 256;
constexpr int kDirectUnroll = 4;

template <typename T, int NI, typename PROG>
__global__ void __launch_bounds__(kDirectThreads)
direct_map_kernel(const DevProgram<T> prog, const DevOperands<NI> ops, const Plan plan) {
  constexpr int VEC = 16 / sizeof(T);
  constexpr int V = VEC * kDirectUnroll;
  const int64_t row_vecs = plan.d2 / VEC;                       // rows are whole vectors (checked on the host)
  const int64_t blocks_per_row = (row_vecs + kDirectThreads * kDirectUnroll - 1) / (kDirectThreads * kDirectUnroll);
  const int64_t b = blockIdx.x;
  const int64_t row = b / blocks_per_row;
  const int64_t vec0 = (b - row * blocks_per_row) * (kDirectThreads * kDirectUnroll) + threadIdx.x;
  const int64_t i0 = row / plan.d1;
  const int64_t i1 = row - i0 * plan.d1;
  T in[NI][V];
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    if (i < ops.n_in) {
      const DevOperand& o = ops.in[i];
      const T* base = static_cast<const T*>(o.ptr) + i0 * o.stride[0] + i1 * o.stride[1];
      if (o.stride[2] == 0) {
        const T x = load_as<T>(o.ptr, o.dtype, i0 * o.stride[0] + i1 * o.stride[1]);
#pragma unroll
        for (int v = 0; v < V; ++v) in[i][v] = x;
      } else {
#pragma unroll
        for (int u = 0; u < kDirectUnroll; ++u) {
          const int64_t vu = vec0 + u * kDirectThreads;
          int4 q = make_int4(0, 0, 0, 0);
          if (vu < row_vecs) q = __ldcs(reinterpret_cast<const int4*>(base) + vu);
          *reinterpret_cast<int4*>(&in[i][u * VEC]) = q;
        }
      }
    }
  }
  T res[V], idx[V];
  if (prog.uses_index) {
    const long long ib = prog.index_base + i0 * prog.index_stride[0] + i1 * prog.index_stride[1];
#pragma unroll
    for (int u = 0; u < kDirectUnroll; ++u)
#pragma unroll
      for (int v = 0; v < VEC; ++v)
        idx[u * VEC + v] = static_cast<T>(ib + ((vec0 + u * kDirectThreads) * VEC + v) * prog.index_stride[2]);
  }
  PROG::template run<T, V, NI>(prog, in, idx, res);
  T* dst = static_cast<T*>(const_cast<void*>(ops.out.ptr)) + i0 * ops.out.stride[0] + i1 * ops.out.stride[1];
#pragma unroll
  for (int u = 0; u < kDirectUnroll; ++u) {
    const int64_t vu = vec0 + u * kDirectThreads;
    if (vu < row_vecs) __stcs(reinterpret_cast<int4*>(dst) + vu, *reinterpret_cast<const int4*>(&res[u * VEC]));
  }
}

// ------------------------------------------------------------------------------------ direct reduce (leading axis)
// The same plain formulation for the fused map+reduce over the leading axis, again for compiled programs only: a thread
// owns 16 bytes of columns, walks a chunk of kDirectRows rows and keeps kDirectUnroll rows' worth of independent 16-byte
// loads per operand in flight; a CTA covers 256 x 16 B of columns.  No shared memory, no barriers: ~0.13 issued
// instructions per element against ~0.6 in the ring kernel (ncu: the ring's consumers sit at 46-54 % issue utilisation,
// which is what made chains with more arithmetic lose bandwidth).  Partials go to scratch[chunk][d0][d2] like MODE 1.
constexpr int kDirectRows = 128;

template <typename T, int NI, typename PROG>
__global__ void __launch_bounds__(kDirectThreads)
direct_reduce_kernel(const DevProgram<T> prog, const DevOperands<NI> ops, const Plan plan, const int red_op,
                     T* __restrict__ scratch) {
  constexpr int VEC = 16 / sizeof(T);
  constexpr int U = kDirectUnroll;
  const int64_t row_vecs = plan.d2 / VEC;
  const int64_t col_blocks = (row_vecs + kDirectThreads - 1) / kDirectThreads;
  int64_t b = blockIdx.x;
  const int64_t cb = b % col_blocks; b /= col_blocks;
  const int64_t chunk = b % plan.n_chunks;
  const int64_t i0 = b / plan.n_chunks;
  const int64_t vec = cb * kDirectThreads + threadIdx.x;
  if (vec >= row_vecs) return;
  const int64_t row0 = chunk * kDirectRows;
  const int64_t row_end = min(plan.d1, row0 + kDirectRows);
  T acc[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) acc[v] = red_identity<T>(red_op);
  for (int64_t r = row0; r < row_end; r += U) {
    T in[NI][VEC * U];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      if (i < ops.n_in) {
        const DevOperand& o = ops.in[i];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int64_t row = min(r + u, row_end - 1);         // rows past the chunk re-read its last row, unused below
          if (o.stride[2] == 0) {
            const T x = load_as<T>(o.ptr, o.dtype, i0 * o.stride[0] + row * o.stride[1]);
#pragma unroll
            for (int v = 0; v < VEC; ++v) in[i][u * VEC + v] = x;
          } else {
            const int4 q = __ldcs(reinterpret_cast<const int4*>(static_cast<const T*>(o.ptr) + i0 * o.stride[0] +
                                                              row * o.stride[1]) + vec);
            *reinterpret_cast<int4*>(&in[i][u * VEC]) = q;
          }
        }
      }
    }
    T res[VEC * U], idx[VEC * U];
    if (prog.uses_index) {
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int v = 0; v < VEC; ++v)
          idx[u * VEC + v] = static_cast<T>(prog.index_base + i0 * prog.index_stride[0] +
                                            min(r + u, row_end - 1) * prog.index_stride[1] +
                                            (vec * VEC + v) * prog.index_stride[2]);
    }
    PROG::template run<T, VEC * U, NI>(prog, in, idx, res);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (r + u < row_end) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = red_apply<T>(red_op, acc[v], res[u * VEC + v]);
      }
    }
  }
  T* dst = scratch + (chunk * plan.d0 + i0) * plan.d2 + vec * VEC;
  *reinterpret_cast<int4*>(dst) = *reinterpret_cast<const int4*>(&acc[0]);
}

}  // namespace stream
}  // namespace sp
