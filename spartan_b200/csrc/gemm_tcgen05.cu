// Dense fp32 contraction for spartan.dot on B200 (sm_100a).
//
// Replaces the per-tile `tiles[0].dot(tiles[1])` BLAS call of the reference
// (spartan/expr/dot.py:195-217 dot_map2_mapper, :222-238 dot_outer_mapper) and
// the `np.add` combiner that merges the rank-k partials
// (spartan/array/tile.pyx:263-268) with one persistent, warp-specialised
// tcgen05 kernel:
//
//   C[m, n] (+)= sum over segments s, terms t, k :  A_{s,t}[m, k] * Bt_{s,t}[n, k]
//
// * operands are *prepared* copies (prep kernels below): fp32 values rounded
//   (round-to-nearest; the tensor core would otherwise truncate) to TF32 or
//   split into bf16 pieces, B transposed so both operands are K-major, K
//   zero-padded to the k-block so TMA strides are 16B aligned;
// * "segments" are the (A strip, B strip) pairs of a tiled dot, "terms" the
//   precision split (x3: lo*hi + hi*lo + hi*hi), so a whole tiled contraction
//   is ONE launch;
// * TMA (cp.async.bulk.tensor, 128B swizzle) -> 4-stage smem ring -> one
//   elected thread issues tcgen05.mma (kind::tf32 M=128,N=256,K=8 or kind::f16
//   bf16 K=16) into a double-buffered 2x256-column TMEM accumulator;
// * the tensor core's fp32 accumulator TRUNCATES on every accumulate (measured
//   on B200: relative bias ~ -6e-9 per K element, profiles/r01_gemm_probe_v1.log),
//   so TMEM only ever holds a short K chunk: every `chunk_kb` k-blocks the 8
//   epilogue warps drain the chunk with tcgen05.ld and add it to fp32 REGISTER
//   accumulators with round-to-nearest (128 per thread; registers are moved from
//   the producer warpgroup to the two epilogue warpgroups with setmaxnreg);
// * precision modes: tf32x1 (1 pass), tf32x3 / bf16x3 (hi/lo split, 3 passes).
//
// Roofline: tensor pipe. Algorithmic work = 2*M*N*K flop per dot.
#include "sp_common.h"
#include <cuda.h>
#include <cuda_bf16.h>
#include <algorithm>
#include <string.h>

namespace sp {
namespace gemm {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * 128;      // 128 rows x 128 B (one swizzle row of K)
constexpr int B_STAGE_BYTES = BN * 128;
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;   // 48 KiB
constexpr int NUM_ACC = 2;
constexpr int TMEM_COLS = NUM_ACC * BN;      // 512
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 128 + 32 * NUM_EPI_WARPS;   // WG0: TMA / MMA / TMEM-alloc / idle, WG1-2: epilogue
constexpr int EPI_WARP0 = 4;
constexpr int GROUP_M = 16;                  // tile rasterisation group (L2 reuse)
constexpr int MAX_SEGMENTS = SP_GEMM_MAX_TERMS;
constexpr int MAX_MAPS = 4 * 8;              // per launch: <= 8 segments x (A_hi, A_lo, B_hi, B_lo)
constexpr int EPI_EXCHANGE_BYTES = 2048;     // epilogue mode 2: per-row (min, arg min) of the upper column half + the 128 row labels
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + EPI_EXCHANGE_BYTES;
constexpr int REGS_PRODUCER = 40;
constexpr int REGS_EPILOGUE = 232;
// CTA-pair variant (cta_group::2): the two CTAs of a cluster own one 256 x 256 tile of C.  Each CTA loads its own 128 rows
// of A and 128 of the 256 rows of Bt, every operand tile ONCE per k-block (hi and lo copies side by side in the stage), and
// the leader issues M=256 MMAs that read both CTAs' shared memory: 64 KiB of L2->smem traffic per CTA per k-block instead
// of 144 KiB, and half the B bytes per MMA out of each SM's shared memory.
constexpr int PAIR_TILE_BYTES = 128 * 128;                 // 128 rows x one 128 B swizzle row of K
constexpr int PAIR_RING_BYTES = 12 * PAIR_TILE_BYTES;      // 192 KiB: 3 stages x 4 tiles (split modes), 6 x 2 (tf32x1)
constexpr int PAIR_SMEM_BYTES = PAIR_RING_BYTES + 1024 + 256 + EPI_EXCHANGE_BYTES;

struct Segment {
  int k_blocks;
  int n_terms;
  int a_map[3];
  int b_map[3];
};

struct Params {
  CUtensorMap maps[MAX_MAPS];
  Segment segs[8];
  int n_segs;
  int chunk_kb;          // k-blocks accumulated in TMEM before promotion to registers
  int M, N;
  int m_blocks, n_blocks;
  int accumulate;
  float* C;
  int64_t ldc;
  // epilogue mode 1 (k-means): instead of storing C, each (row, 128-column half tile) emits
  //   min_j (col_bias[j] - 2 * C[row, j]) and its column index  ->  part_val / part_idx [M][n_blocks * 2]
  int epi_mode;
  const float* col_bias;
  float* part_val;
  int* part_idx;
  // pair kernel: round counter that keeps the clusters' K loops within a few k-blocks of each other, so that the operand
  // tiles several clusters share are still in L2 when the next one asks for them (nullptr = free running)
  unsigned int* round_sync;
  int sync_kb;           // additional check-ins every sync_kb k-blocks inside a tile (0 = only at tile starts)
  int group_m;           // tile rasterisation: tile rows per group
  // epilogue mode 2 (k-means, fused): tiles are walked n-fastest per cluster (raster 1) so that a CTA sees every column
  // tile of its 128 rows back to back; the running (min, arg min) stays in registers, and after the last column tile the
  // epilogue warps write the labels and add the rows of `pts` (the points in fp32) into sums[label] / counts[label]
  int raster;            // 0: grouped rasterisation (tile_coords), 1: all column tiles of a row tile consecutively
  const float* pts;
  int64_t ldp;
  int pts_d;
  int* labels;
  float* sums;
  unsigned long long* counts;
  // gated segments (multi-GPU dot): segment s may only be read once *seg_flag[s] has reached seg_flag_value[s] -- the word
  // a peer's copy engine writes after it has pushed the operand strip of that segment into this GPU's memory (peer.cu)
  const unsigned int* seg_flag[8];
  unsigned int seg_flag_value[8];
  unsigned int* gate_status;   // incremented when a gate times out (the launch then finishes on whatever is there)
};

// ----------------------------------------------------------------------------
// PTX wrappers
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
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
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc]; fp32 accumulate.  KIND 0: tf32, 1: bf16 (kind::f16).
template <int KIND>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
  if constexpr (KIND == 0) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum)
        : "memory");
  } else {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum)
        : "memory");
  }
}
// Arrives on `bar` once every previously issued tcgen05.mma of this thread retired.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

// ---- CTA-pair (cta_group::2) forms ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t num_clusters_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on a barrier that lives in another CTA of the cluster (address from map_to_cta)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  // default semantics (what CUTLASS' ClusterBarrier::arrive issues): the TMEM hand-over is ordered by
  // tcgen05.fence::before_thread_sync; an explicit .release.cluster here compiled to MEMBAR.ALL + ERRBAR + CGAERRBAR per
  // chunk and cost the argmin epilogue ~12% of its time (ncu source view of the k-means GEMM, profiles/)
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
// TMA load whose completion bytes are credited to a barrier in the pair's leader CTA
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t cluster_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(cluster_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
template <int KIND>
__device__ __forceinline__ void umma_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
  if constexpr (KIND == 0) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum)
        : "memory");
  } else {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum)
        : "memory");
  }
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs of the pair once the pair's MMAs retired
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"(static_cast<uint16_t>(3)) : "memory");
}

// K-major, 128B-swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart.
// Field layout follows the sm_100 shared-memory matrix descriptor
// (start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout [61,64)).
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(1) << 16;              // LBO (unused for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;      // SBO: next 8-row group
  d |= static_cast<uint64_t>(1) << 46;              // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;              // SWIZZLE_128B
  return d;
}

// instruction descriptor: D fp32, A/B format fmt (0 f16, 1 bf16, 2 tf32), both K-major
__host__ __device__ constexpr uint32_t make_idesc(int fmt, int m, int n) {
  return (1u << 4) | (static_cast<uint32_t>(fmt) << 7) | (static_cast<uint32_t>(fmt) << 10) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

__device__ __forceinline__ void tile_coords(int tile, int m_blocks, int n_blocks, int group_m, int& m_blk, int& n_blk) {
  const int tiles_per_group = group_m * n_blocks;
  const int group = tile / tiles_per_group;
  const int first_m = group * group_m;
  const int rows_in_group = min(group_m, m_blocks - first_m);
  const int in_group = tile - group * tiles_per_group;
  m_blk = first_m + in_group % rows_in_group;
  n_blk = in_group / rows_in_group;
}

// The v-th tile of this CTA / cluster, or false when it has none left.
template <typename P>
__device__ __forceinline__ bool tile_at(const P& p, int v, int first, int step, int m_tiles, int& m_blk, int& n_blk) {
  if (p.raster == 1) {
    const int mt = first + (v / p.n_blocks) * step;
    if (mt >= m_tiles) return false;
    m_blk = mt;
    n_blk = v % p.n_blocks;
    return true;
  }
  const int tile = first + v * step;
  if (tile >= m_tiles * p.n_blocks) return false;
  tile_coords(tile, m_tiles, p.n_blocks, p.group_m, m_blk, n_blk);
  return true;
}

// One check-in of the soft round barrier: `participants` clusters pass this point; wait until all have -- for at most
// ~100 us (in step, clusters arrive within a few us of each other).  A cluster that times out stops waiting for the rest
// of the launch (it keeps checking in): when not every cluster of the grid is resident -- another kernel, e.g. an NCCL
// collective, holds some SMs -- the barrier costs one time-out per cluster instead of one per round, and nothing can
// deadlock.
__device__ __forceinline__ void round_checkin(unsigned int* counter, unsigned int& target, int participants, bool& wait) {
  target += static_cast<unsigned int>(participants);
  atomicAdd(counter, 1u);
  if (!wait) return;
  for (int spin = 0; spin < 160; ++spin) {
    unsigned int seen;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
    if (seen >= target) return;
    __nanosleep(256);
  }
  wait = false;
}

// Gate of a K segment whose operand strip is pushed into this GPU's memory by a peer (peer.cu: data copy, then a flag word,
// both by the copy engine, in stream order).  The TMA thread polls the flag with a system-scope acquire load; the fence
// orders that generic-proxy read before the async-proxy (TMA) reads of the strip.  Bounded: after ~4 s the launch goes on
// and reports through gate_status, so a lost peer cannot hang the GPU.
__device__ __forceinline__ void gate_wait(const unsigned int* flag, unsigned int value, unsigned int* status) {
  unsigned long long t0 = 0;
  for (unsigned int spin = 0;; ++spin) {
    unsigned int seen;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
    if (static_cast<int>(seen - value) >= 0) break;
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (spin == 0) t0 = t;
    if (t - t0 > 4000000000ull) {
      if (status != nullptr) atomicAdd(status, 1u);
      break;
    }
    __nanosleep(128);
  }
  asm volatile("fence.proxy.async.global;" ::: "memory");
}

// ----------------------------------------------------------------------------
// The kernel.  KIND 0: tf32 operands (BK = 32 elements), KIND 1: bf16 (BK = 64).
// PAIR false: one CTA per 128 x 256 tile, a pipeline stage = the (A, Bt) tiles of one precision term.
// PAIR true : launched as clusters of 2; the pair owns a 256 x 256 tile (this CTA: rows rank*128 .. +128), a stage =
//             every distinct operand tile of one k-block ([A copies][Bt-half copies], 16 KiB each); only the leader
//             (cluster rank 0) issues MMAs (cta_group::2, M = 256) and its commits arrive on both CTAs' barriers.
// ----------------------------------------------------------------------------
constexpr int MAX_STAGES = 6;

template <int KIND, bool PAIR>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_kernel(const __grid_constant__ Params p) {
  constexpr int ELEM = (KIND == 0) ? 4 : 2;
  constexpr int BK = 128 / ELEM;               // elements per 128B swizzle row
  constexpr int UMMA_K = 32 / ELEM;            // 32 bytes of K per MMA
  constexpr uint32_t IDESC = make_idesc(KIND == 0 ? 2 : 1, PAIR ? 2 * BM : BM, BN);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int copies = (p.segs[0].n_terms == 3) ? 2 : 1;
  const int stage_bytes = PAIR ? 2 * copies * PAIR_TILE_BYTES : STAGE_BYTES;
  const int n_stages = PAIR ? PAIR_RING_BYTES / stage_bytes : STAGES;
  const uint32_t bar_base = smem_base + (PAIR ? PAIR_RING_BYTES : STAGES * STAGE_BYTES);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * MAX_STAGES + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * MAX_STAGES + NUM_ACC + a); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * MAX_STAGES + 2 * NUM_ACC);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;          // 0 = leader of the pair
  const int first_tile = PAIR ? static_cast<int>(cluster_id_x()) : static_cast<int>(blockIdx.x);
  const int tile_step = PAIR ? static_cast<int>(num_clusters_x()) : static_cast<int>(gridDim.x);
  const int m_tiles = PAIR ? (p.m_blocks + 1) / 2 : p.m_blocks;  // tile rows (256 rows each in pair mode)
  const int num_tiles = m_tiles * p.n_blocks;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.n_segs; ++s)
      for (int t = 0; t < (PAIR ? copies : p.segs[s].n_terms); ++t) {
        tma_prefetch_desc(&p.maps[p.segs[s].a_map[t]]);
        tma_prefetch_desc(&p.maps[p.segs[s].b_map[t]]);
      }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < NUM_ACC; ++a) {
      mbar_init(tmem_full_bar(a), 1);
      // one arrive per epilogue warp; in pair mode the leader's barrier also collects the peer's warps
      mbar_init(tmem_empty_bar(a), PAIR ? 2 * NUM_EPI_WARPS : NUM_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if constexpr (PAIR) tmem_alloc_pair(tmem_ptr_smem, TMEM_COLS);
    else tmem_alloc(tmem_ptr_smem, TMEM_COLS);
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();      // the peer's barriers must exist before anything arrives on them
  else __syncthreads();
  tc_fence_after();

  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  if (warp < EPI_WARP0) {
    reg_dec<REGS_PRODUCER>();
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        unsigned int round_target = 0;
        bool round_wait = true;
        unsigned int gate_seen = 0;      // segments whose gate this CTA has passed
        int m_blk, n_blk;
        for (int v = 0; tile_at(p, v, first_tile, tile_step, m_tiles, m_blk, n_blk); ++v) {
          const int tile = first_tile + v * tile_step;      // (raster 0; the round barrier is off for raster 1)
          if constexpr (PAIR) {
            // Soft round barrier (leaders only; the peer is throttled through the ring): every cluster with a tile in
            // this round checks in, then waits -- for a bounded time, so nothing can deadlock -- until all have.
            if (p.round_sync != nullptr && rank == 0 && tile != first_tile)
              round_checkin(p.round_sync, round_target, min(tile_step, num_tiles - (tile - first_tile)), round_wait);
          }
          for (int s = 0; s < p.n_segs; ++s) {
            const Segment seg = p.segs[s];
            if (p.seg_flag[s] != nullptr && !((gate_seen >> s) & 1u)) {
              gate_wait(p.seg_flag[s], p.seg_flag_value[s], p.gate_status);
              gate_seen |= 1u << s;
            }
            for (int kb = 0; kb < seg.k_blocks; ++kb) {
              if constexpr (PAIR) {
                if (p.round_sync != nullptr && rank == 0 && p.sync_kb > 0 && kb > 0 && kb % p.sync_kb == 0)
                  round_checkin(p.round_sync, round_target, min(tile_step, num_tiles - (tile - first_tile)), round_wait);
                mbar_wait(empty_bar(stage), phase ^ 1);
                // the leader's barrier counts the bytes of both CTAs' loads of this stage
                if (rank == 0) mbar_expect_tx(full_bar(stage), 2u * stage_bytes);
                const uint32_t leader_full = map_to_cta(full_bar(stage), 0);
                const uint32_t dst = smem_base + stage * stage_bytes;
                const int a_row = (m_blk * 2 + static_cast<int>(rank)) * BM;
                const int b_row = n_blk * BN + static_cast<int>(rank) * (BN / 2);
                for (int c = 0; c < copies; ++c)
                  tma_load_2d_pair(dst + c * PAIR_TILE_BYTES, &p.maps[seg.a_map[c]], leader_full, kb * BK, a_row);
                for (int c = 0; c < copies; ++c)
                  tma_load_2d_pair(dst + (copies + c) * PAIR_TILE_BYTES, &p.maps[seg.b_map[c]], leader_full, kb * BK, b_row);
                if (++stage == n_stages) { stage = 0; phase ^= 1; }
              } else {
                for (int t = 0; t < seg.n_terms; ++t) {
                  mbar_wait(empty_bar(stage), phase ^ 1);
                  const uint32_t a_dst = smem_base + stage * STAGE_BYTES;
                  const uint32_t b_dst = a_dst + A_STAGE_BYTES;
                  mbar_expect_tx(full_bar(stage), STAGE_BYTES);
                  tma_load_2d(a_dst, &p.maps[seg.a_map[t]], full_bar(stage), kb * BK, m_blk * BM);
                  tma_load_2d(b_dst, &p.maps[seg.b_map[t]], full_bar(stage), kb * BK, n_blk * BN);
                  if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
              }
            }
          }
        }
      }
      __syncwarp();
    } else if (warp == 1 && rank == 0) {
      // ===================== MMA issuer =====================
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        int mma_m, mma_n;
        for (int v = 0; tile_at(p, v, first_tile, tile_step, m_tiles, mma_m, mma_n); ++v) {
          int in_chunk = 0;          // k-blocks issued into the current TMEM chunk
          uint32_t accum = 0;
          bool chunk_open = false;
          for (int s = 0; s < p.n_segs; ++s) {
            const int k_blocks = p.segs[s].k_blocks;
            const int n_terms = p.segs[s].n_terms;
            for (int kb = 0; kb < k_blocks; ++kb) {
              if (!chunk_open) {
                mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1);   // epilogue(s) drained this buffer
                tc_fence_after();
                chunk_open = true;
                accum = 0;
                in_chunk = 0;
              }
              const uint32_t tmem_d = tmem_base + acc * BN;
              if constexpr (PAIR) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                const uint32_t base = smem_base + stage * stage_bytes;
                for (int t = 0; t < n_terms; ++t) {
                  // small cross terms first, dominant term last: (lo,hi) (hi,lo) (hi,hi); single copy: (hi,hi)
                  const int ai = (n_terms == 3 && t == 0) ? 1 : 0;
                  const int bi = (n_terms == 3 && t == 1) ? 1 : 0;
                  const uint32_t a_addr = base + ai * PAIR_TILE_BYTES;
                  const uint32_t b_addr = base + (copies + bi) * PAIR_TILE_BYTES;
#pragma unroll
                  for (int k = 0; k < BK / UMMA_K; ++k) {
                    umma_pair<KIND>(tmem_d, make_kmajor_sw128_desc(a_addr + k * 32), make_kmajor_sw128_desc(b_addr + k * 32),
                                    IDESC, accum);
                    accum = 1;
                  }
                }
                umma_commit_pair(empty_bar(stage));   // frees this stage in both CTAs when these MMAs retire
                if (++stage == n_stages) { stage = 0; phase ^= 1; }
              } else {
                for (int t = 0; t < n_terms; ++t) {
                  mbar_wait(full_bar(stage), phase);
                  tc_fence_after();
                  const uint32_t a_addr = smem_base + stage * STAGE_BYTES;
                  const uint32_t b_addr = a_addr + A_STAGE_BYTES;
#pragma unroll
                  for (int k = 0; k < BK / UMMA_K; ++k) {
                    const uint64_t da = make_kmajor_sw128_desc(a_addr + k * 32);
                    const uint64_t db = make_kmajor_sw128_desc(b_addr + k * 32);
                    umma<KIND>(tmem_d, da, db, IDESC, accum);
                    accum = 1;
                  }
                  umma_commit(empty_bar(stage));      // frees the smem slot when these MMAs retire
                  if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
              }
              if (++in_chunk == p.chunk_kb) {
                // chunk complete -> epilogue promotes it
                if constexpr (PAIR) umma_commit_pair(tmem_full_bar(acc)); else umma_commit(tmem_full_bar(acc));
                chunk_open = false;
                if (++acc == NUM_ACC) { acc = 0; acc_phase ^= 1; }
              }
            }
          }
          if (chunk_open) {
            if constexpr (PAIR) umma_commit_pair(tmem_full_bar(acc)); else umma_commit(tmem_full_bar(acc));
            if (++acc == NUM_ACC) { acc = 0; acc_phase ^= 1; }
          }
        }
      }
      __syncwarp();
    }
  } else {
    // ===================== Epilogue: TMEM chunk -> fp32 register accumulators -> C =====================
    reg_inc<REGS_EPILOGUE>();
    const int q = warp & 3;                        // TMEM lane quarter this warp may read (warp % 4)
    const int half = (warp - EPI_WARP0) >> 2;      // which 128-column half of the tile
    int total_kb = 0;
    for (int s = 0; s < p.n_segs; ++s) total_kb += p.segs[s].k_blocks;
    const int n_chunks = (total_kb + p.chunk_kb - 1) / p.chunk_kb;
    int acc = 0;
    uint32_t acc_phase = 0;
    float run_best = 3.402823466e+38f;     // epilogue mode 2: minimum over the column tiles seen so far of this row tile
    int run_best_j = 0x7fffffff;
    int m_blk, n_blk;
    for (int v = 0; tile_at(p, v, first_tile, tile_step, m_tiles, m_blk, n_blk); ++v) {
      if constexpr (PAIR) m_blk = m_blk * 2 + static_cast<int>(rank);     // this CTA's 128-row block
      if (p.epi_mode == 0 && p.accumulate) {
        // C (+)= : pull this thread's 512 bytes of the old C into L2 while the tile's K loop runs
        const int prow = m_blk * BM + q * 32 + lane;
        if (prow < p.M) {
          const char* ptr = reinterpret_cast<const char*>(p.C + static_cast<int64_t>(prow) * p.ldc + n_blk * BN + half * 128);
#pragma unroll
          for (int i = 0; i < 4; ++i) asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr + 128 * i));
        }
      }
      if (p.epi_mode == 2 && n_blk == 0) {
        // the fp32 rows this CTA will add into the centroids after its last column tile: pull them into L2 now
        const int lines_per_row = (p.pts_d * 4 + 127) / 128;
        const int t = (warp - EPI_WARP0) * 32 + lane;
        for (int i = t; i < BM * lines_per_row; i += 32 * NUM_EPI_WARPS) {
          const int64_t grow = static_cast<int64_t>(m_blk) * BM + i / lines_per_row;
          if (grow < p.M) {
            const char* ptr = reinterpret_cast<const char*>(p.pts + grow * p.ldp) + (i % lines_per_row) * 128;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
          }
        }
      }
      float sum[128];
#pragma unroll
      for (int j = 0; j < 128; ++j) sum[j] = 0.0f;
      for (int c = 0; c < n_chunks; ++c) {
        mbar_wait(tmem_full_bar(acc), acc_phase);
        tc_fence_after();
        const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + half * 128;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(tbase + g * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) sum[g * 32 + j] += __uint_as_float(r[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (PAIR) mbar_arrive_cluster(map_to_cta(tmem_empty_bar(acc), 0));   // the issuer lives in the leader
          else mbar_arrive(tmem_empty_bar(acc));
        }
        if (++acc == NUM_ACC) { acc = 0; acc_phase ^= 1; }
      }
      const int row = m_blk * BM + q * 32 + lane;
      const int col0 = n_blk * BN + half * 128;
      if (p.epi_mode != 0) {
        float best = 3.402823466e+38f;
        int best_j = 0x7fffffff;
        if (col0 + 128 <= p.N && (reinterpret_cast<uint64_t>(p.col_bias) & 15) == 0) {
          // full half tile: 16-byte bias loads (uniform address: one transaction per warp), bias - 2*sum as one FMA (the
          // doubling is exact, so it rounds once like FMUL + FADD), four independent (min, arg min) chains over j mod 4
          // merged at the end -- the scalar loop below had the epilogue, not the tensor cores, bound this kernel
          const float4* b4 = reinterpret_cast<const float4*>(p.col_bias + col0);
          float bv[4] = {best, best, best, best};
          int bj[4] = {best_j, best_j, best_j, best_j};
#pragma unroll
          for (int j4 = 0; j4 < 32; ++j4) {     // fully unrolled: sum[] must stay in registers
            const float4 b = __ldg(b4 + j4);
            const float v0 = __fmaf_rn(-2.0f, sum[4 * j4], b.x), v1 = __fmaf_rn(-2.0f, sum[4 * j4 + 1], b.y);
            const float v2 = __fmaf_rn(-2.0f, sum[4 * j4 + 2], b.z), v3 = __fmaf_rn(-2.0f, sum[4 * j4 + 3], b.w);
            if (v0 < bv[0]) { bv[0] = v0; bj[0] = 4 * j4; }
            if (v1 < bv[1]) { bv[1] = v1; bj[1] = 4 * j4 + 1; }
            if (v2 < bv[2]) { bv[2] = v2; bj[2] = 4 * j4 + 2; }
            if (v3 < bv[3]) { bv[3] = v3; bj[3] = 4 * j4 + 3; }
          }
#pragma unroll
          for (int c = 0; c < 4; ++c)           // ties resolve to the smallest index, like np.argmin
            if (bv[c] < best || (bv[c] == best && bj[c] < best_j)) { best = bv[c]; best_j = bj[c]; }
          best_j += col0;
        } else {
#pragma unroll
          for (int j = 0; j < 128; ++j) {       // fully unrolled: sum[] must stay in registers
            const int col = col0 + j;
            if (col < p.N) {
              const float v = __ldg(p.col_bias + col) - 2.0f * sum[j];
              if (v < best) { best = v; best_j = col; }        // first minimum: ties resolve to the smallest index
            }
          }
        }
        if (p.epi_mode == 1) {
          if (row < p.M) {
            const int64_t o = static_cast<int64_t>(row) * (p.n_blocks * 2) + n_blk * 2 + half;
            p.part_val[o] = best;
            p.part_idx[o] = best_j;
          }
        } else {
          // fused k-means: keep the running minimum of this row over the column tiles (columns ascend, so a strict
          // comparison keeps the smallest index among equal values)
          if (best < run_best) { run_best = best; run_best_j = best_j; }
          if (n_blk == p.n_blocks - 1) {
            // combine the two column halves of the row (the two epilogue warpgroups), publish the labels of the CTA's
            // 128 rows in shared memory, then all 8 epilogue warps add the rows into their centroids
            float* ex_val = reinterpret_cast<float*>(smem_raw + (bar_base + 256u - smem_u32(smem_raw)));
            int* ex_idx = reinterpret_cast<int*>(ex_val + 128);
            int* row_label = ex_idx + 128;
            const int r_in = q * 32 + lane;
            if (half == 1) { ex_val[r_in] = run_best; ex_idx[r_in] = run_best_j; }
            asm volatile("bar.sync 2, 256;" ::: "memory");
            if (half == 0) {
              const float ov = ex_val[r_in];
              const int oj = ex_idx[r_in];
              if (ov < run_best) { run_best = ov; run_best_j = oj; }     // half 1 holds the larger indices
              row_label[r_in] = run_best_j;
              if (row < p.M) {
                p.labels[row] = run_best_j;
                atomicAdd(p.counts + run_best_j, 1ull);
              }
            }
            asm volatile("bar.sync 2, 256;" ::: "memory");
            const int ew = warp - EPI_WARP0;                 // 0..7: rows ew*16 .. ew*16+15 of the CTA's block
            const int d4 = p.pts_d >> 2;
            // 8 rows at a time: all their loads first (16 independent 16-byte loads per lane in flight; the rows were
            // prefetched into L2 when this row tile began), then their atomics -- a load directly followed by its atomic
            // would serialise 32 memory round trips per warp and stall the tensor cores behind the epilogue
            for (int jb = 0; jb < d4; jb += 64) {
              const int ja = jb + lane, jc = jb + 32 + lane;
#pragma unroll
              for (int hb = 0; hb < 2; ++hb) {
                float4 buf[16];
#pragma unroll
                for (int rr = 0; rr < 8; ++rr) {
                  const int64_t grow = static_cast<int64_t>(m_blk) * BM + ew * 16 + hb * 8 + rr;
                  const float4* src = reinterpret_cast<const float4*>(p.pts + grow * p.ldp);
                  buf[2 * rr] = (grow < p.M && ja < d4) ? __ldcs(src + ja) : make_float4(0.f, 0.f, 0.f, 0.f);
                  buf[2 * rr + 1] = (grow < p.M && jc < d4) ? __ldcs(src + jc) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int rr = 0; rr < 8; ++rr) {
                  const int rl = ew * 16 + hb * 8 + rr;
                  const int64_t grow = static_cast<int64_t>(m_blk) * BM + rl;
                  if (grow < p.M) {
                    float4* dst = reinterpret_cast<float4*>(p.sums + static_cast<int64_t>(row_label[rl]) * p.pts_d);
                    if (ja < d4) atomicAdd(dst + ja, buf[2 * rr]);                 // red.global.add.v4.f32
                    if (jc < d4) atomicAdd(dst + jc, buf[2 * rr + 1]);
                  }
                }
              }
            }
            asm volatile("bar.sync 2, 256;" ::: "memory");   // row_label is rewritten for the next row tile
            run_best = 3.402823466e+38f;
            run_best_j = 0x7fffffff;
          }
        }
      } else if (row < p.M) {
        float* c_row = p.C + static_cast<int64_t>(row) * p.ldc;
        const bool vec_ok = ((reinterpret_cast<uint64_t>(p.C) & 15) == 0) && ((p.ldc & 3) == 0) && (col0 + 128 <= p.N);
        if (vec_ok) {
          float4* dst = reinterpret_cast<float4*>(c_row + col0);
          if (p.accumulate) {
            // C (+)= : the old values were prefetched into L2 when the tile began; read them 8 x 16 B at a time BEFORE
            // the stores of the batch (a load directly followed by its store serialises 32 memory round trips per
            // thread and the tensor cores end up waiting for this epilogue)
#pragma unroll
            for (int jb = 0; jb < 32; jb += 8) {
              float4 o[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) o[j] = __ldcs(dst + jb + j);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int jj = jb + j;
                __stcs(dst + jj, make_float4(sum[4 * jj] + o[j].x, sum[4 * jj + 1] + o[j].y, sum[4 * jj + 2] + o[j].z,
                                             sum[4 * jj + 3] + o[j].w));
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              dst[j] = make_float4(sum[4 * j], sum[4 * j + 1], sum[4 * j + 2], sum[4 * j + 3]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 128; ++j) {
            if (col0 + j < p.N) {
              float v = sum[j];
              if (p.accumulate) v += c_row[col0 + j];
              c_row[col0 + j] = v;
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();      // nobody leaves while the pair may still touch its shared memory / TMEM
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ----------------------------------------------------------------------------
// Operand preparation (HBM-bound: reads the fp32 operand once, writes 1 or 2 rounded copies).
//   SPLIT 0: tf32 hi            SPLIT 1: tf32 hi + lo
//   SPLIT 2: bf16 hi + lo
// Output rows have pitch Kp (the caller zero-fills what no strip covers); B is written transposed ([N, Kp]).
// ----------------------------------------------------------------------------
__device__ __forceinline__ float rn_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

template <int SPLIT>
__device__ __forceinline__ void split_store(float a, void* hi, void* lo, int64_t idx) {
  if constexpr (SPLIT == 2) {
    const __nv_bfloat16 h = __float2bfloat16_rn(a);
    static_cast<__nv_bfloat16*>(hi)[idx] = h;
    static_cast<__nv_bfloat16*>(lo)[idx] = __float2bfloat16_rn(a - __bfloat162float(h));
  } else {
    const float h = rn_tf32(a);
    static_cast<float*>(hi)[idx] = h;
    if constexpr (SPLIT == 1) static_cast<float*>(lo)[idx] = rn_tf32(a - h);
  }
}

// A: [M, K] (lda) -> hi/lo rows of pitch Kp.  One thread per 4 consecutive k (float4 load when aligned).
template <int SPLIT>
__global__ void prep_a_kernel(const float* __restrict__ A, int64_t lda, int M, int K, int Kw, int64_t Kp, void* hi,
                              void* lo) {
  const int kq = Kw / 4;       // Kw: width written per row (K rounded up to 4, zero tail); Kp: destination row pitch
  const int64_t total = static_cast<int64_t>(M) * kq;
  const bool vec = ((reinterpret_cast<uint64_t>(A) & 15) == 0) && ((lda & 3) == 0);
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t m = i / kq;
    const int k = static_cast<int>(i - m * kq) * 4;
    float v[4];
    if (vec && k + 4 <= K) {
      const float4 x = *reinterpret_cast<const float4*>(A + m * lda + k);
      v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = (k + j < K) ? A[m * lda + k + j] : 0.0f;
    }
    // four consecutive k of one row: one 8-byte (bf16) or 16-byte (tf32) store per copy (Kp and k are multiples of 4
    // and the operand buffers are 128-byte aligned, so the stores are aligned)
    const int64_t o = m * Kp + k;
    if constexpr (SPLIT == 2) {
      __nv_bfloat16 h[4], l[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        h[j] = __float2bfloat16_rn(v[j]);
        l[j] = __float2bfloat16_rn(v[j] - __bfloat162float(h[j]));
      }
      *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(hi) + o) = *reinterpret_cast<const uint2*>(h);
      *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(lo) + o) = *reinterpret_cast<const uint2*>(l);
    } else {
      float h[4], l[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        h[j] = rn_tf32(v[j]);
        l[j] = rn_tf32(v[j] - h[j]);
      }
      *reinterpret_cast<float4*>(static_cast<float*>(hi) + o) = make_float4(h[0], h[1], h[2], h[3]);
      if constexpr (SPLIT == 1) *reinterpret_cast<float4*>(static_cast<float*>(lo) + o) = make_float4(l[0], l[1], l[2], l[3]);
    }
  }
}

// B: [K, N] (ldb) -> hiT/loT [N, Kp] through a 32x33 shared tile (coalesced both ways).
template <int SPLIT>
__global__ void prep_bt_kernel(const float* __restrict__ B, int64_t ldb, int K, int N, int64_t Kp, void* hiT, void* loT) {
  __shared__ float tile[32][33];
  const int n0 = blockIdx.x * 32;
  const int k0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int k = k0 + r, n = n0 + threadIdx.x;
    tile[r][threadIdx.x] = (k < K && n < N) ? B[static_cast<int64_t>(k) * ldb + n] : 0.0f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int n = n0 + r, k = k0 + threadIdx.x;
    if (n < N && k < K) split_store<SPLIT>(tile[threadIdx.x][r], hiT, loT, static_cast<int64_t>(n) * Kp + k);
  }
}

// ----------------------------------------------------------------------------
// Host side
// ----------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess) {
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// 2-D K-major tensor [rows, Kp] of `elem`-byte elements, box = [box_rows, 128 B of K], 128B swizzle.
static int make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t Kp, int box_rows, int elem) {
  EncodeTiledFn enc = get_encode_fn();
  SP_REQUIRE(enc != nullptr, SP_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(Kp), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(Kp) * elem};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / elem), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, elem == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SP_REQUIRE(r == CUDA_SUCCESS, SP_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
  return SP_OK;
}

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

struct Mode {
  int kind;      // kernel KIND
  int elem;      // operand element bytes
  int bk;        // elements per k-block
  int copies;    // prepared copies per operand
  int terms;
  int split;     // prep SPLIT
  int chunk_kb;  // default promotion interval
};

static bool mode_of(int precision, Mode* m) {
  switch (precision) {
    case SP_GEMM_TF32X1: *m = Mode{0, 4, 32, 1, 1, 0, 8}; return true;
    case SP_GEMM_TF32X3: *m = Mode{0, 4, 32, 2, 3, 1, 8}; return true;
    case SP_GEMM_BF16X3: *m = Mode{1, 2, 64, 2, 3, 2, 8}; return true;
    default: return false;
  }
}

static int g_chunk_override = 0;
static int g_variant = 0;      // 0: choose, 1: one CTA per tile, 2: CTA pairs
static int g_sync_kb = 0;
static int g_group_m = 8;       // 256-row tile rows per rasterisation group (pair kernel); the one-CTA kernel uses twice as many 128-row blocks
static int g_round_sync = 1;   // pair kernel: keep the clusters' tile rounds in step (L2 reuse of shared operand tiles)
static int g_round_sync_min_kb = 64;   // round barrier only for contractions at least this many k-blocks deep
static unsigned int* g_round_counter = nullptr;

}  // namespace gemm
}  // namespace sp

using namespace sp;
using namespace sp::gemm;

// Test / tuning hook: k-blocks accumulated inside TMEM before promotion (0 = per-mode default).
extern "C" int sp_gemm_set_chunk_kblocks(int kb) {
  SP_REQUIRE(kb >= 0 && kb <= (1 << 20), SP_ERR_INVALID, "bad chunk size %d", kb);
  g_chunk_override = kb;
  return SP_OK;
}

// Test / tuning hook: 0 = choose by shape (default), 1 = one CTA per 128x256 tile, 2 = CTA pairs on 256x256 tiles.
extern "C" int sp_gemm_set_variant(int variant) {
  SP_REQUIRE(variant >= 0 && variant <= 2, SP_ERR_INVALID, "bad gemm variant %d", variant);
  g_variant = variant;
  return SP_OK;
}

// Test / tuning hook: 1 = the CTA pairs of one launch start every tile round together (soft barrier), 0 = free running.
extern "C" int sp_gemm_set_round_sync(int on) {
  g_round_sync = on ? 1 : 0;
  return SP_OK;
}
// Test / tuning hooks: extra check-ins every `sync_kb` k-blocks (0 = tile starts only); tile rows per rasterisation group.
extern "C" int sp_gemm_set_tuning(int sync_kb, int group_m) {
  SP_REQUIRE(sync_kb >= 0 && group_m >= 1 && group_m <= 1024, SP_ERR_INVALID, "bad tuning values");
  g_sync_kb = sync_kb;
  g_group_m = group_m;
  return SP_OK;
}

extern "C" int64_t sp_gemm_kpad(int64_t K, int precision) {
  Mode md;
  if (!mode_of(precision, &md)) return -1;
  return round_up(K, md.bk);
}

// Bytes of one prepared operand with `rows` rows (M for A, N for B) and padded depth Kp: [copies][rows][Kp].
extern "C" int64_t sp_gemm_prepared_bytes(int64_t rows, int64_t Kp, int precision) {
  Mode md;
  if (!mode_of(precision, &md)) return -1;
  return rows * Kp * md.elem * md.copies;
}

static int check_prep(const void* out, int64_t out_bytes, int64_t rows, int64_t Kp, int precision, Mode* md) {
  SP_REQUIRE(mode_of(precision, md), SP_ERR_INVALID, "unknown precision %d", precision);
  SP_REQUIRE(Kp > 0 && Kp % md->bk == 0, SP_ERR_INVALID, "Kp=%lld is not a multiple of the k-block %d", (long long)Kp, md->bk);
  SP_REQUIRE(out != nullptr && (reinterpret_cast<uint64_t>(out) % 128) == 0, SP_ERR_INVALID,
             "prepared operand buffer must be 128-byte aligned");
  SP_REQUIRE(out_bytes >= rows * Kp * md->elem * md->copies, SP_ERR_INVALID, "prepared operand buffer too small");
  return SP_OK;
}

static int check_prep_rows(const void* out, int64_t copy_stride, int64_t rows, int64_t Kp, int precision, Mode* md) {
  SP_REQUIRE(mode_of(precision, md), SP_ERR_INVALID, "unknown precision %d", precision);
  SP_REQUIRE(Kp > 0 && Kp % md->bk == 0, SP_ERR_INVALID, "Kp=%lld is not a multiple of the k-block %d", (long long)Kp, md->bk);
  SP_REQUIRE(out != nullptr && (reinterpret_cast<uint64_t>(out) % 128) == 0, SP_ERR_INVALID,
             "prepared operand buffer must be 128-byte aligned");
  SP_REQUIRE(md->copies == 1 || (copy_stride >= rows * Kp * md->elem && copy_stride % 128 == 0), SP_ERR_INVALID,
             "bad copy stride %lld", (long long)copy_stride);
  return SP_OK;
}

// Row-range forms: `out` points at the first destination row inside copy 0 of a larger prepared operand
// [copies][rows_total][Kp]; copy 1 (the lo halves) starts copy_stride bytes after copy 0.  They let a pipelined
// dot prepare one strip at a time into a full-size operand and contract any row range of it.
extern "C" int sp_gemm_prepare_a_rows(const float* A, int64_t lda, int64_t M, int64_t K, int precision, void* out,
                                      int64_t copy_stride, int64_t Kp, int64_t k_offset, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  Mode md;
  int rc = check_prep_rows(out, copy_stride, M, Kp, precision, &md);
  if (rc) return rc;
  SP_REQUIRE(k_offset >= 0 && k_offset % 4 == 0 && k_offset + K <= Kp, SP_ERR_INVALID, "bad k_offset %lld", (long long)k_offset);
  uint8_t* hi = static_cast<uint8_t*>(out) + k_offset * md.elem;
  uint8_t* lo = md.copies == 2 ? hi + copy_stride : nullptr;
  // rows of the destination are Kp apart; the kernel pads K up to a multiple of 4 within its strip
  const int64_t Kq = std::min<int64_t>(round_up(K, 4), Kp - k_offset);
  const int64_t total = M * (Kq / 4);
  const int threads = 256;
  const int blocks = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((total + threads - 1) / threads, num_sms() * 32)));
  const int Mi = static_cast<int>(M), Ki = static_cast<int>(K), Kqi = static_cast<int>(Kq);
  if (md.split == 0) prep_a_kernel<0><<<blocks, threads, 0, stream>>>(A, lda, Mi, Ki, Kqi, Kp, hi, lo);
  else if (md.split == 1) prep_a_kernel<1><<<blocks, threads, 0, stream>>>(A, lda, Mi, Ki, Kqi, Kp, hi, lo);
  else prep_a_kernel<2><<<blocks, threads, 0, stream>>>(A, lda, Mi, Ki, Kqi, Kp, hi, lo);
  SP_CUDA_CHECK(cudaGetLastError());
  return SP_OK;
}

extern "C" int sp_gemm_prepare_b_rows(const float* B, int64_t ldb, int64_t K, int64_t N, int precision, void* out,
                                      int64_t copy_stride, int64_t Kp, int64_t k_offset, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  Mode md;
  int rc = check_prep_rows(out, copy_stride, N, Kp, precision, &md);
  if (rc) return rc;
  SP_REQUIRE(k_offset >= 0 && k_offset + K <= Kp, SP_ERR_INVALID, "bad k_offset %lld", (long long)k_offset);
  uint8_t* hi = static_cast<uint8_t*>(out) + k_offset * md.elem;
  uint8_t* lo = md.copies == 2 ? hi + copy_stride : nullptr;
  dim3 grid(static_cast<unsigned>((N + 31) / 32), static_cast<unsigned>((K + 31) / 32));
  dim3 block(32, 8);
  const int Ni = static_cast<int>(N), Ki = static_cast<int>(K);
  if (md.split == 0) prep_bt_kernel<0><<<grid, block, 0, stream>>>(B, ldb, Ki, Ni, Kp, hi, lo);
  else if (md.split == 1) prep_bt_kernel<1><<<grid, block, 0, stream>>>(B, ldb, Ki, Ni, Kp, hi, lo);
  else prep_bt_kernel<2><<<grid, block, 0, stream>>>(B, ldb, Ki, Ni, Kp, hi, lo);
  SP_CUDA_CHECK(cudaGetLastError());
  return SP_OK;
}

// A strip [M, K] (leading dim lda) -> out[copies][M][Kp] at depth offset k_offset (columns k_offset .. k_offset+K).
// The caller zero-fills the buffer when the strips it writes do not cover [0, Kp).
extern "C" int sp_gemm_prepare_a(const float* A, int64_t lda, int64_t M, int64_t K, int precision, void* out,
                                 int64_t Kp, int64_t k_offset, int64_t out_bytes, void* stream_) {
  Mode md;
  int rc = check_prep(out, out_bytes, M, Kp, precision, &md);
  if (rc) return rc;
  return sp_gemm_prepare_a_rows(A, lda, M, K, precision, out, M * Kp * md.elem, Kp, k_offset, stream_);
}

// B strip [K, N] (leading dim ldb) -> transposed out[copies][N][Kp] at depth offset k_offset.
extern "C" int sp_gemm_prepare_b(const float* B, int64_t ldb, int64_t K, int64_t N, int precision, void* out,
                                 int64_t Kp, int64_t k_offset, int64_t out_bytes, void* stream_) {
  Mode md;
  int rc = check_prep(out, out_bytes, N, Kp, precision, &md);
  if (rc) return rc;
  return sp_gemm_prepare_b_rows(B, ldb, K, N, precision, out, N * Kp * md.elem, Kp, k_offset, stream_);
}

struct Gates {
  const uint32_t* const* flags;    // per segment: word to poll (NULL = not gated)
  const uint32_t* values;          // per segment: value the word must reach
  uint32_t* status;                // time-out counter (device memory) or NULL
};

struct FusedKmeans {
  const float* pts;            // the points in fp32 (rows of the A operand before preparation)
  int64_t ldp;
  int d;
  int32_t* labels;
  float* sums;
  unsigned long long* counts;
};

static int launch_prepared(int n_seg, const sp_gemm_prepared_view* segs, float* C, int64_t ldc, int64_t M, int64_t N,
                           int accumulate, int precision, int epi_mode, const float* col_bias, float* part_val,
                           int* part_idx, void* stream_, const Gates* gates = nullptr, const FusedKmeans* fused = nullptr) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  Mode md;
  SP_REQUIRE(mode_of(precision, &md), SP_ERR_INVALID, "sp_gemm_prepared: unknown precision %d", precision);
  SP_REQUIRE(n_seg >= 1 && n_seg <= 8, SP_ERR_INVALID, "sp_gemm_prepared: %d segments (limit 8 per launch)", n_seg);
  SP_REQUIRE(M > 0 && N > 0 && M < (1ll << 31) && N < (1ll << 31), SP_ERR_INVALID, "bad M/N %lld %lld",
             (long long)M, (long long)N);
  // CTA pairs pay off as soon as there are two 128-row blocks to pair up
  const bool pair = g_variant == 2 || (g_variant == 0 && M > BM);
  static bool attr_set = false;
  static int max_clusters = 0;
  if (!attr_set) {
    SP_CUDA_CHECK(cudaFuncSetAttribute(gemm_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    SP_CUDA_CHECK(cudaFuncSetAttribute(gemm_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    SP_CUDA_CHECK(cudaFuncSetAttribute(gemm_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM_BYTES));
    SP_CUDA_CHECK(cudaFuncSetAttribute(gemm_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM_BYTES));
    cudaLaunchConfig_t qc = {};
    cudaLaunchAttribute qa[1];
    qa[0].id = cudaLaunchAttributeClusterDimension;
    qa[0].val.clusterDim.x = 2; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
    qc.gridDim = dim3(static_cast<unsigned>(num_sms() / 2 * 2));
    qc.blockDim = dim3(NUM_THREADS);
    qc.dynamicSmemBytes = PAIR_SMEM_BYTES;
    qc.attrs = qa; qc.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, gemm_kernel<1, true>, &qc) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = num_sms() / 2;
    }
    max_clusters = n;
    attr_set = true;
  }
  Params p;
  memset(&p, 0, sizeof(p));
  int n_maps = 0;
  const int b_box = pair ? BN / 2 : BN;
  for (int s = 0; s < n_seg; ++s) {
    const sp_gemm_prepared_view& g = segs[s];
    SP_REQUIRE(g.Kp > 0 && g.Kp % md.bk == 0 && g.A != nullptr && g.B != nullptr, SP_ERR_INVALID, "bad prepared segment %d", s);
    SP_REQUIRE(md.copies == 1 || (g.a_copy_stride >= M * g.Kp * md.elem && g.b_copy_stride >= N * g.Kp * md.elem), SP_ERR_INVALID,
               "prepared segment %d: copy stride smaller than the row range", s);
    const uint8_t* a_hi = static_cast<const uint8_t*>(g.A);
    const uint8_t* b_hi = static_cast<const uint8_t*>(g.B);
    Segment& sg = p.segs[s];
    sg.k_blocks = static_cast<int>(g.Kp / md.bk);
    sg.n_terms = md.terms;
    const int ia_hi = n_maps++, ib_hi = n_maps++;
    int rc = make_map(&p.maps[ia_hi], a_hi, M, g.Kp, BM, md.elem);
    if (rc) return rc;
    rc = make_map(&p.maps[ib_hi], b_hi, N, g.Kp, b_box, md.elem);
    if (rc) return rc;
    if (md.copies == 2) {
      const int ia_lo = n_maps++, ib_lo = n_maps++;
      rc = make_map(&p.maps[ia_lo], a_hi + g.a_copy_stride, M, g.Kp, BM, md.elem);
      if (rc) return rc;
      rc = make_map(&p.maps[ib_lo], b_hi + g.b_copy_stride, N, g.Kp, b_box, md.elem);
      if (rc) return rc;
      if (pair) {
        // the pair kernel indexes the copies: [0] = hi, [1] = lo
        sg.a_map[0] = ia_hi; sg.b_map[0] = ib_hi;
        sg.a_map[1] = ia_lo; sg.b_map[1] = ib_lo;
      } else {
        // per-term operand pairs: small cross terms first, dominant term last
        sg.a_map[0] = ia_lo; sg.b_map[0] = ib_hi;
        sg.a_map[1] = ia_hi; sg.b_map[1] = ib_lo;
        sg.a_map[2] = ia_hi; sg.b_map[2] = ib_hi;
      }
    } else {
      sg.a_map[0] = ia_hi; sg.b_map[0] = ib_hi;
    }
  }
  p.n_segs = n_seg;
  if (gates != nullptr) {
    for (int s = 0; s < n_seg; ++s) {
      p.seg_flag[s] = gates->flags ? gates->flags[s] : nullptr;
      p.seg_flag_value[s] = gates->values ? gates->values[s] : 0u;
    }
    p.gate_status = gates->status;
  }
  p.chunk_kb = g_chunk_override > 0 ? g_chunk_override : md.chunk_kb;
  p.M = static_cast<int>(M);
  p.N = static_cast<int>(N);
  p.m_blocks = static_cast<int>((M + BM - 1) / BM);
  p.n_blocks = static_cast<int>((N + BN - 1) / BN);
  p.accumulate = accumulate;
  p.C = C;
  p.ldc = ldc;
  p.group_m = pair ? g_group_m : 2 * g_group_m;
  p.sync_kb = g_sync_kb;
  p.epi_mode = epi_mode;
  if (fused != nullptr) {
    p.raster = 1;
    p.pts = fused->pts; p.ldp = fused->ldp; p.pts_d = fused->d;
    p.labels = fused->labels; p.sums = fused->sums; p.counts = fused->counts;
  }
  p.col_bias = col_bias;
  p.part_val = part_val;
  p.part_idx = part_idx;
  if (pair) {
    const int tiles = (p.m_blocks + 1) / 2 * p.n_blocks;
    // The round barrier pays when a tile round lasts long enough for the clusters to drift apart (deep K: the operand
    // tiles they share would fall out of L2); with a shallow K a round is a few microseconds and a grid-wide check-in per
    // tile costs more than it saves (k-means: 4 k-blocks per tile, tensor pipe 62% busy with it, ncu r2c_apps).
    int total_kb = 0;
    for (int s = 0; s < n_seg; ++s) total_kb += p.segs[s].k_blocks;
    if (g_round_sync && tiles > max_clusters && total_kb >= g_round_sync_min_kb) {
      if (g_round_counter == nullptr) SP_CUDA_CHECK(cudaMalloc(&g_round_counter, sizeof(unsigned int)));
      SP_CUDA_CHECK(cudaMemsetAsync(g_round_counter, 0, sizeof(unsigned int), stream));
      p.round_sync = g_round_counter;
    }
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    const int units = p.raster == 1 ? (p.m_blocks + 1) / 2 : tiles;      // raster 1: a cluster takes whole row tiles
    cfg.gridDim = dim3(static_cast<unsigned>(2 * std::min(units, max_clusters)));
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = PAIR_SMEM_BYTES;
    cfg.stream = stream;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (md.kind == 0) SP_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm_kernel<0, true>, p));
    else SP_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm_kernel<1, true>, p));
  } else {
    const int tiles = p.raster == 1 ? p.m_blocks : p.m_blocks * p.n_blocks;
    const int grid = std::min(tiles, num_sms());
    if (md.kind == 0) gemm_kernel<0, false><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(p);
    else gemm_kernel<1, false><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(p);
  }
  SP_CUDA_CHECK(cudaGetLastError());
  return SP_OK;
}

static int to_views(int n_seg, const sp_gemm_prepared_segment* segs, int64_t M, int64_t N, int precision,
                    sp_gemm_prepared_view* out) {
  Mode md;
  SP_REQUIRE(mode_of(precision, &md), SP_ERR_INVALID, "unknown precision %d", precision);
  SP_REQUIRE(n_seg >= 1 && n_seg <= 8, SP_ERR_INVALID, "%d segments (limit 8 per launch)", n_seg);
  for (int s = 0; s < n_seg; ++s) {
    out[s].A = segs[s].A; out[s].B = segs[s].B; out[s].Kp = segs[s].Kp;
    out[s].a_copy_stride = M * segs[s].Kp * md.elem;
    out[s].b_copy_stride = N * segs[s].Kp * md.elem;
  }
  return SP_OK;
}

// C[M,N] (+)= sum_s A_s . B_s over PREPARED operands (sp_gemm_prepare_a / _b), one launch.
extern "C" int sp_gemm_prepared(int n_seg, const sp_gemm_prepared_segment* segs, float* C, int64_t ldc, int64_t M,
                                int64_t N, int accumulate, int precision, void* stream_) {
  SP_REQUIRE(C != nullptr, SP_ERR_INVALID, "sp_gemm_prepared: null C");
  sp_gemm_prepared_view v[8];
  int rc = to_views(n_seg, segs, M, N, precision, v);
  if (rc) return rc;
  return launch_prepared(n_seg, v, C, ldc, M, N, accumulate, precision, 0, nullptr, nullptr, nullptr, stream_);
}

// Same over row ranges of larger prepared operands (sp_gemm_prepare_a_rows / _b_rows).
extern "C" int sp_gemm_prepared_views(int n_seg, const sp_gemm_prepared_view* segs, float* C, int64_t ldc, int64_t M,
                                      int64_t N, int accumulate, int precision, void* stream_) {
  SP_REQUIRE(C != nullptr, SP_ERR_INVALID, "sp_gemm_prepared_views: null C");
  return launch_prepared(n_seg, segs, C, ldc, M, N, accumulate, precision, 0, nullptr, nullptr, nullptr, stream_);
}

// Same, with gated segments: segment s is read only after *ready_flag[s] has reached ready_value[s] (wrap-safe compare;
// ready_flag[s] == NULL: not gated).  The flags are words in this GPU's memory that peers write behind the operand strips
// they push (sp_peer_push): the contraction starts on the strips that are present and picks the others up as they land,
// inside ONE launch -- the device side of the dot row x column shuffle (dot.py:195-238, map.py:243-286 join_mapper).
// A gate that stays closed for ~4 s is given up and counted in *gate_status (device uint32, may be NULL).
extern "C" int sp_gemm_prepared_views_gated(int n_seg, const sp_gemm_prepared_view* segs, const uint32_t* const* ready_flag,
                                            const uint32_t* ready_value, uint32_t* gate_status, float* C, int64_t ldc,
                                            int64_t M, int64_t N, int accumulate, int precision, void* stream_) {
  SP_REQUIRE(C != nullptr, SP_ERR_INVALID, "sp_gemm_prepared_views_gated: null C");
  Gates g{ready_flag, ready_value, gate_status};
  return launch_prepared(n_seg, segs, C, ldc, M, N, accumulate, precision, 0, nullptr, nullptr, nullptr, stream_, &g);
}

// Fused row-argmin epilogue (k-means assignment, k_means_.py:61-66): for every row and every 128-column half tile
// emits  min_j (col_bias[j] - 2 * (A.B)[row, j])  and the arg min; parts = sp_gemm_argmin_parts(N) entries per row.
extern "C" int64_t sp_gemm_argmin_parts(int64_t N) { return ((N + BN - 1) / BN) * 2; }

extern "C" int sp_gemm_prepared_argmin(int n_seg, const sp_gemm_prepared_segment* segs, int64_t M, int64_t N,
                                       const float* col_bias, float* part_val, int32_t* part_idx, int precision,
                                       void* stream_) {
  SP_REQUIRE(col_bias && part_val && part_idx, SP_ERR_INVALID, "sp_gemm_prepared_argmin: null pointer");
  sp_gemm_prepared_view v[8];
  int rc = to_views(n_seg, segs, M, N, precision, v);
  if (rc) return rc;
  return launch_prepared(n_seg, v, nullptr, 0, M, N, 0, precision, 1, col_bias, part_val, part_idx, stream_);
}

// k-means assignment fused end to end (k_means_.py:61-97): labels[i] = argmin_j (col_bias[j] - 2 (A.B)[i, j]) over ALL
// columns, counts[labels[i]] += 1, sums[labels[i], :] += pts[i, :] -- the GEMM walks the column tiles of a row tile back to
// back, keeps the running arg min in registers and its epilogue warps do the accumulation while the tensor cores are
// already on the next row tile.  d must be a multiple of 4, pts / sums 16-byte aligned.
extern "C" int sp_gemm_prepared_kmeans(const sp_gemm_prepared_segment* seg, int64_t M, int64_t N, const float* col_bias,
                                       const float* pts, int64_t ldp, int64_t d, int32_t* labels, float* sums,
                                       int64_t* counts, int precision, void* stream_) {
  SP_REQUIRE(seg && col_bias && pts && labels && sums && counts, SP_ERR_INVALID, "sp_gemm_prepared_kmeans: null pointer");
  SP_REQUIRE(d > 0 && (d & 3) == 0 && (ldp & 3) == 0 && d < (1ll << 31) &&
             ((reinterpret_cast<uint64_t>(pts) | reinterpret_cast<uint64_t>(sums)) & 15) == 0, SP_ERR_INVALID,
             "sp_gemm_prepared_kmeans: d / ldp must be multiples of 4 and pts / sums 16-byte aligned");
  sp_gemm_prepared_view v[1];
  int rc = to_views(1, seg, M, N, precision, v);
  if (rc) return rc;
  FusedKmeans fk{pts, ldp, static_cast<int>(d), labels, sums, reinterpret_cast<unsigned long long*>(counts)};
  return launch_prepared(1, v, nullptr, 0, M, N, 0, precision, 2, col_bias, nullptr, nullptr, stream_, nullptr, &fk);
}

extern "C" int64_t sp_gemm_f32_workspace_bytes(int64_t M, int64_t N, int n_seg, const int64_t* seg_k, int precision) {
  Mode md;
  if (!mode_of(precision, &md)) return -1;
  int64_t total = 0;
  for (int s = 0; s < n_seg; ++s) {
    const int64_t Kp = round_up(seg_k[s], md.bk);
    total += round_up(M * Kp * md.elem * md.copies, 1024) + round_up(N * Kp * md.elem * md.copies, 1024);
  }
  return total + 1024;
}

// Convenience: prepare + contract in one call (workspace holds the prepared operands).
extern "C" int sp_gemm_f32_segments(int n_seg, const sp_gemm_segment* segs, float* C, int64_t ldc, int64_t M,
                                     int64_t N, int accumulate, int precision, void* workspace,
                                     int64_t workspace_bytes, void* stream_) {
  Mode md;
  SP_REQUIRE(mode_of(precision, &md), SP_ERR_INVALID, "sp_gemm_f32_segments: unknown precision %d", precision);
  SP_REQUIRE(n_seg >= 1 && n_seg <= 8, SP_ERR_INVALID, "sp_gemm_f32_segments: %d segments (limit 8 per launch)", n_seg);
  SP_REQUIRE(M > 0 && N > 0, SP_ERR_INVALID, "bad M/N");
  {
    int64_t ks[8];
    for (int s = 0; s < n_seg; ++s) ks[s] = segs[s].K;
    const int64_t need = sp_gemm_f32_workspace_bytes(M, N, n_seg, ks, precision);
    SP_REQUIRE(workspace != nullptr && workspace_bytes >= need, SP_ERR_INVALID,
               "sp_gemm_f32_segments: workspace %lld B < required %lld B", (long long)workspace_bytes, (long long)need);
  }
  uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uint64_t>(workspace) + 1023) & ~1023ull);
  sp_gemm_prepared_segment ps[8];
  for (int s = 0; s < n_seg; ++s) {
    const sp_gemm_segment& g = segs[s];
    SP_REQUIRE(g.K > 0, SP_ERR_INVALID, "segment %d has K=%lld", s, (long long)g.K);
    const int64_t Kp = round_up(g.K, md.bk);
    const int64_t a_bytes = round_up(M * Kp * md.elem * md.copies, 1024);
    const int64_t b_bytes = round_up(N * Kp * md.elem * md.copies, 1024);
    uint8_t* a_buf = ws;
    uint8_t* b_buf = ws + a_bytes;
    ws += a_bytes + b_bytes;
    if (Kp != round_up(g.K, 4)) SP_CUDA_CHECK(cudaMemsetAsync(a_buf, 0, a_bytes, static_cast<cudaStream_t>(stream_)));
    if (Kp != g.K) SP_CUDA_CHECK(cudaMemsetAsync(b_buf, 0, b_bytes, static_cast<cudaStream_t>(stream_)));
    int rc = sp_gemm_prepare_a(g.A, g.lda, M, g.K, precision, a_buf, Kp, 0, a_bytes, stream_);
    if (rc) return rc;
    rc = sp_gemm_prepare_b(g.B, g.ldb, g.K, N, precision, b_buf, Kp, 0, b_bytes, stream_);
    if (rc) return rc;
    ps[s].A = a_buf; ps[s].B = b_buf; ps[s].Kp = Kp;
  }
  return sp_gemm_prepared(n_seg, ps, C, ldc, M, N, accumulate, precision, stream_);
}

// C (+)= op(A) . op(B) for operands that may be TRANSPOSED VIEWS (spartan/expr/operator/transpose.py:27-67):
// a_trans != 0: A is given as At [K, M] row-major (lda = its leading dimension); b_trans != 0: B is given as
// Bt [N, K] row-major.  Nothing is materialised: a K-major Bt is exactly what the tensor path consumes (it goes
// through the A-style preparation), an M-major At goes through the transposing B-style preparation.
extern "C" int sp_gemm_f32_ex(const float* A, int64_t lda, int a_trans, const float* B, int64_t ldb, int b_trans,
                               float* C, int64_t ldc, int64_t M, int64_t N, int64_t K, int accumulate, int precision,
                               void* workspace, int64_t workspace_bytes, void* stream_) {
  Mode md;
  SP_REQUIRE(mode_of(precision, &md), SP_ERR_INVALID, "sp_gemm_f32_ex: unknown precision %d", precision);
  SP_REQUIRE(M > 0 && N > 0 && K > 0 && A && B && C, SP_ERR_INVALID, "sp_gemm_f32_ex: bad arguments");
  const int64_t need = sp_gemm_f32_workspace_bytes(M, N, 1, &K, precision);
  SP_REQUIRE(workspace != nullptr && workspace_bytes >= need, SP_ERR_INVALID,
             "sp_gemm_f32_ex: workspace %lld B < required %lld B", (long long)workspace_bytes, (long long)need);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uint64_t>(workspace) + 1023) & ~1023ull);
  const int64_t Kp = round_up(K, md.bk);
  const int64_t a_bytes = round_up(M * Kp * md.elem * md.copies, 1024);
  const int64_t b_bytes = round_up(N * Kp * md.elem * md.copies, 1024);
  uint8_t* a_buf = ws;
  uint8_t* b_buf = ws + a_bytes;
  if (Kp != K) {
    SP_CUDA_CHECK(cudaMemsetAsync(a_buf, 0, a_bytes, stream));
    SP_CUDA_CHECK(cudaMemsetAsync(b_buf, 0, b_bytes, stream));
  }
  int rc = a_trans ? sp_gemm_prepare_b(A, lda, K, M, precision, a_buf, Kp, 0, a_bytes, stream_)
                   : sp_gemm_prepare_a(A, lda, M, K, precision, a_buf, Kp, 0, a_bytes, stream_);
  if (rc) return rc;
  rc = b_trans ? sp_gemm_prepare_a(B, ldb, N, K, precision, b_buf, Kp, 0, b_bytes, stream_)
               : sp_gemm_prepare_b(B, ldb, K, N, precision, b_buf, Kp, 0, b_bytes, stream_);
  if (rc) return rc;
  sp_gemm_prepared_segment seg;
  seg.A = a_buf; seg.B = b_buf; seg.Kp = Kp;
  return sp_gemm_prepared(1, &seg, C, ldc, M, N, accumulate, precision, stream_);
}

extern "C" int sp_gemm_f32(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
                            int64_t M, int64_t N, int64_t K, int accumulate, int precision, void* workspace,
                            int64_t workspace_bytes, void* stream) {
  sp_gemm_segment seg;
  seg.A = A; seg.lda = lda; seg.B = B; seg.ldb = ldb; seg.K = K;
  return sp_gemm_f32_segments(1, &seg, C, ldc, M, N, accumulate, precision, workspace, workspace_bytes, stream);
}
