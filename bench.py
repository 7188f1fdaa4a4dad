#!/usr/bin/env python
"""bench.py -- the hot path of BASELINE.json on N B200s of one node.

  python bench.py --gpus 1 --steps 5 --warmup 3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference ...        # the reference's CPU path (NumPy oracle) on the host cores

Headline workload (BASELINE.json configs[1]): spartan.dot of two 32768x32768 fp32 tiled arrays,
tile_hint=(4096,4096); metric = GFLOP/s with 2*N^3 algorithmic flops.  A "step" is one evaluation of the
DotExpr over inputs already resident in HBM (``value``) or starting from host buffers through the public
API with the H2D/D2H copies inside the timed region (``e2e``).  The same run also measures, each with its own
roofline and parity: the fused map+reduce (configs[2]) and its neighbours (a chain outside the kernel catalogue, a
pure map, trailing-axis and flat reductions), k-means (configs[3]) and the PageRank SpMV (configs[4]).
One JSON line on stdout (rank 0).
"""
import os
import sys


def _host_cores():
  try:
    return len(os.sched_getaffinity(0))
  except Exception:
    return os.cpu_count() or 1


if '--impl' in sys.argv and sys.argv[sys.argv.index('--impl') + 1:][:1] == ['reference']:
  # The reference arm is BLAS on the host cores.  torchrun exports OMP_NUM_THREADS=1 to every rank; the arm must not
  # inherit that, so the thread counts are pinned here, before NumPy (and its OpenBLAS) is loaded.
  for _v in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
    os.environ[_v] = str(_host_cores())

import argparse
import json
import subprocess
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
  if p not in sys.path:
    sys.path.insert(0, p)

TOL = 1e-5       # north_star: results within 1e-5 (normwise: max|got - ref| <= TOL * max|ref|)


def parse_args():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=5)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
  ap.add_argument('--n', type=int, default=32768, help='matrix order of the dot workload')
  ap.add_argument('--tile', type=int, default=4096)
  ap.add_argument('--precision', default=os.environ.get('SPARTAN_DOT_PRECISION', 'bf16x3'))
  ap.add_argument('--mr-log2', type=int, default=30, help='log2(#elements) of the map+reduce workload')
  ap.add_argument('--km-n', type=int, default=10000000, help='points of the k-means workload (x 256 dims, k = 1024)')
  ap.add_argument('--pr-n', type=int, default=10000000, help='pages of the PageRank SpMV workload (10 outlinks each)')
  ap.add_argument('--skip-e2e', action='store_true')
  ap.add_argument('--skip-dot', action='store_true', help='tuning aid: no dot workload')
  ap.add_argument('--skip-variants', action='store_true', help='no tf32x3 / uncached / randn legs of the dot workload')
  ap.add_argument('--skip-mapreduce', action='store_true')
  ap.add_argument('--skip-apps', action='store_true', help='no k-means / SpMV workloads')
  ap.add_argument('--skip-cpu', action='store_true')
  ap.add_argument('--no-graph', action='store_true', help='time the map+reduce step eagerly only')
  return ap.parse_args()


def measured_traffic(key):
  for name in ('r02_traffic.json', 'r01_traffic.json'):
    try:
      with open(os.path.join(ROOT, 'profiles', name)) as f:
        v = json.load(f).get(key)
      if v is not None:
        return v
    except Exception:
      pass
  return None


def measured_peaks():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(path):
    with open(path) as f:
      d = json.load(f)
    return {'hbm_gbs': d['hbm_gbs'], 'bf16_tflops': d['bf16_tflops'],
            'bf16_tflops_sustained': d.get('bf16_tflops_sustained', d['bf16_tflops']), 'source': 'measured'}
  return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'source': 'fallback'}


def workload_config(n, tile):
  """The ``config`` object, identical in both arms."""
  return {'workload': 'spartan.dot %dx%d fp32, tile_hint=(%d,%d)' % (n, n, tile, tile),
          'l2': 'operands (%.1f GiB) larger than L2 / than the host caches' % (2 * n * n * 4 / 2 ** 30)}


def blas_threads():
  try:
    import threadpoolctl
    return max([p.get('num_threads', 1) for p in threadpoolctl.threadpool_info() if p.get('user_api') == 'blas'] or [1])
  except Exception:
    return None


class ClockSampler(object):
  """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""
  Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
       'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
       'clocks_event_reasons.sw_power_cap')

  def __init__(self, index):
    self.index = index
    self.lines = []
    self.proc = None

  def start(self):
    try:
      self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                    '--format=csv,noheader,nounits', '-lms', '20'], stdout=subprocess.PIPE,
                                   stderr=subprocess.DEVNULL, text=True)
      self.t = threading.Thread(target=self._read, daemon=True)
      self.t.start()
    except Exception:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.lines.append((time.perf_counter(), line.strip()))

  def mark_begin(self):
    self.t0 = time.perf_counter()

  def mark_end(self):
    self.t1 = time.perf_counter()

  def stop(self):
    if self.proc is None:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
    time.sleep(0.05)
    self.proc.terminate()
    try:
      self.proc.wait(timeout=2)
    except Exception:
      self.proc.kill()
    sm, mx, reasons = [], [], set()
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    t0, t1 = getattr(self, 't0', 0.0), getattr(self, 't1', float('inf'))
    inside = [ln for ts, ln in self.lines if t0 <= ts <= t1 + 0.03]
    if not inside:                       # very short timed region: fall back to the samples nearest to it
      inside = [ln for ts, ln in self.lines if ts >= t0 - 0.05][:3]
    for ln in inside:
      f = [x.strip() for x in ln.split(',')]
      if len(f) < 9:
        continue
      try:
        sm.append(float(f[1])); mx.append(float(f[2]))
      except ValueError:
        continue
      for name, val in zip(names, f[5:9]):
        if val.lower().startswith('active'):
          reasons.add(name)
    if not sm:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
    return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons),
            'samples': len(sm)}


# ----------------------------------------------------------------------------------------------- reference arm
def reference_dot_setup(n, tile):
  """Oracle arrays of the configured shape: the A tiles of ONE tile row (what a bounded step joins), all of B, an empty
  C.  Values repeat one random tile (BLAS time does not depend on them)."""
  import spartan_oracle
  from spartan_oracle import distarray as odist, extent as oext
  spartan_oracle.initialize(8)                       # num_workers := 8 ranks of the box (placement only; one process)
  rng = np.random.default_rng(0)
  t = min(tile, n)
  blk_a = rng.random((t, t), dtype=np.float32)
  blk_b = rng.random((t, t), dtype=np.float32)
  A = odist.create((n, n), np.float32, tile_hint=(t, t))
  B = odist.create((n, n), np.float32, tile_hint=(t, t))
  C = odist.create((n, n), np.float32, reducer=np.add, tile_hint=(t, t))
  a_tiles = []
  for ex in A.tiles:
    if ex.ul[0] == 0:
      A.update(ex, blk_a[:ex.shape[0], :ex.shape[1]])
      a_tiles.append(ex)
  for ex in B.tiles:
    B.update(ex, blk_b[:ex.shape[0], :ex.shape[1]])
  return A, B, C, a_tiles


def reference_dot_step(A, B, C, a_tiles):
  """One bounded step of the reference's tiled dot: the K-joins of every A tile of one tile row -- per tile: stitch the
  B row block (DistArrayImpl.fetch), BLAS sgemm (dot_map2_mapper, dot.py:195-217), np.add-merge of the partial into the
  C tiles (Tile.merge, tile.pyx:263-268).  1/(n/tile) of the whole job; returns its flop count."""
  from spartan_oracle import expr as oexpr
  for ex in a_tiles:
    oexpr.dot_grid_join_mapper(ex, (A, B), C)
  n = A.shape[0]
  return 2.0 * a_tiles[0].shape[0] * n * n


def run_reference(args):
  """``--impl reference``: the reference's own CPU implementation of the path (oracle restatement, NumPy/BLAS on every
  host core), honouring --steps / --warmup.  Under torchrun only rank 0 works."""
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  n, tile = args.n, args.tile
  cores = _host_cores()
  try:
    import threadpoolctl
    limiter = threadpoolctl.threadpool_limits(limits=cores, user_api='blas')
  except Exception:
    limiter = None
  A, B, C, a_tiles = reference_dot_setup(n, tile)
  for _ in range(args.warmup):
    reference_dot_step(A, B, C, a_tiles)
  t0 = time.perf_counter()
  flops = 0.0
  for _ in range(args.steps):
    flops += reference_dot_step(A, B, C, a_tiles)
  t = (time.perf_counter() - t0) / max(1, args.steps)
  flops /= max(1, args.steps)
  val = flops / t / 1e9
  threads = blas_threads()
  sample = ('one tile row of the tiled dot per step: %d K-joins, each (%d x %d).(%d x %d) sgemm + B row-block stitch + '
            'np.add merge into %d C tiles = 1/%d of the job; oracle restatement of dot.py:195-217 / tile.pyx:263-268, '
            'OpenBLAS threads = %s' % (len(a_tiles), a_tiles[0].shape[0], a_tiles[0].shape[1], a_tiles[0].shape[1], n,
                                       len(a_tiles), max(1, n // max(1, a_tiles[0].shape[0])), threads))
  out = {'metric': 'spartan.dot fp32 GFLOP/s', 'value': val, 'unit': 'GFLOP/s', 'n_gpus': args.gpus,
         'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': t * 1e3, 'higher_is_better': True,
         'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'impl': 'reference',
         'config': workload_config(n, tile),
         'cpu_baseline': {'value': val, 'unit': 'GFLOP/s', 'cores': threads or cores, 'kind': 'port', 'sample': sample},
         'e2e': {'value': val, 'unit': 'GFLOP/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
         'gpu_launches': 0, 'host_cores': cores}
  del A, B, C
  out['map_reduce'] = cpu_map_reduce_baseline()
  if limiter is not None:
    limiter.restore_original_limits()
  print(json.dumps(out), flush=True)


def cpu_map_reduce_baseline(log2_elems=27, tiles=8):
  """The reference's CPU path for configs[2] on a bounded sample: the oracle's own evaluator (tile_mapper + _reduce_mapper
  + Tile.merge: one NumPy temporary per ufunc, map.py:48-88, reduce.py:21-70) on one thread, and the same per-tile chain
  on a pool of `tiles` threads (emulates num_workers=8; NumPy releases the GIL)."""
  from concurrent.futures import ThreadPoolExecutor
  import spartan_oracle
  from spartan_oracle import expr as oexpr
  cols = 32768
  rows = (1 << log2_elems) // cols
  rng = np.random.default_rng(2)
  x = rng.random((rows, cols), dtype=np.float32); y = rng.random((rows, cols), dtype=np.float32)
  trow = rows // tiles
  spartan_oracle.initialize(tiles)
  oexpr.eval_cache.clear()
  X = oexpr.from_numpy(x, tile_hint=(trow, cols)).evaluate(); Y = oexpr.from_numpy(y, tile_hint=(trow, cols)).evaluate()
  t0 = time.perf_counter()
  ref = (oexpr.lazify(X) * 2 + oexpr.lazify(Y)).sum(axis=0).optimized().glom()
  t_single = time.perf_counter() - t0

  def tile_fn(i):
    xs, ys = x[i * trow:(i + 1) * trow], y[i * trow:(i + 1) * trow]
    return (xs * np.float32(2) + ys).sum(axis=0)

  def run(pool):
    parts = list(pool.map(tile_fn, range(tiles)))
    out = parts[0]
    for p in parts[1:]:
      out = np.add(out, p)
    return out
  nbytes = 2.0 * 4.0 * rows * cols
  with ThreadPoolExecutor(tiles) as pool:
    run(pool)
    t0 = time.perf_counter(); got = run(pool); t_pool = time.perf_counter() - t0
  assert np.allclose(got, ref, rtol=1e-5)
  return {'value': nbytes / t_pool / 1e9, 'unit': 'GB/s', 'cores': min(tiles, _host_cores()), 'kind': 'port',
          'single_thread_gbs': nbytes / t_single / 1e9,
          'sample': '(x*2+y).sum(axis=0) over 2^%d fp32 elements in %d row tiles: oracle evaluator on one thread '
                    '(single_thread_gbs) and the per-tile NumPy chain on %d tile threads (value)' % (log2_elems, tiles, tiles)}


# ----------------------------------------------------------------------------------------------- b200 arm
class Bench(object):
  def __init__(self, args):
    import torch
    import torch.distributed as dist
    import spartan_b200 as sp
    from spartan_b200 import comm
    self.torch, self.dist, self.sp, self.comm = torch, dist, sp, comm
    self.args = args
    self.ctx = sp.initialize()
    self.rank, self.world = self.ctx.worker_id, self.ctx.num_workers
    assert self.world == args.gpus or self.world == 1, 'launch with torchrun for --gpus %d' % args.gpus
    self.peaks = measured_peaks()
    self.holder = {}

  # ---------------------------------------------------------------- timing helpers
  def sync(self):
    self.comm.barrier()
    self.torch.cuda.synchronize()

  def maxreduce(self, ms):
    if self.world == 1:
      return ms
    t = self.torch.tensor([ms], device=self.ctx.device, dtype=self.torch.float64)
    self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
    return float(t.item())

  def timed(self, fn, steps, warmup):
    """ms per step: W untimed steps, barrier + synchronize, K steps between two CUDA events on the launching stream,
    barrier + synchronize, max over ranks."""
    torch = self.torch
    for _ in range(warmup):
      fn()
    self.sync()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
      fn()
    ev1.record()
    self.sync()
    return self.maxreduce(ev0.elapsed_time(ev1)) / steps

  def to_host(self, t):
    return None if t is None else t.cpu().numpy()

  # ---------------------------------------------------------------- dot
  def dot_parity(self, A, B, C, label):
    """Rank 0 gathers 64 rows of A and C (16 groups of 4, spread over the matrix) and 8 groups of 128 columns of B (one
    group inside every rank's column block where there are up to 8 ranks) over NCCL point-to-point and compares
    C[rows, cols] with a float64 product on the host."""
    sp = self.sp
    n = A.shape[0]
    rng = np.random.default_rng(12345)
    row_groups = sorted(int(r) for r in rng.integers(0, max(1, n - 4), 16))
    ngroups = 8
    col_groups = [int(g * (n // ngroups) + rng.integers(0, max(1, n // ngroups - 128))) for g in range(ngroups)]
    gw = min(128, n)
    a_rows, c_rows = [], []
    for r in row_groups:
      ex = sp.extent.create((r, 0), (min(n, r + 4), n), (n, n))
      a_rows.append(self.to_host(A.fetch(ex, dst=0)))
      c_rows.append(self.to_host(C.fetch(ex, dst=0)))
    b_cols = [self.to_host(B.fetch(sp.extent.create((0, c), (n, min(n, c + gw)), (n, n)), dst=0)) for c in col_groups]
    if self.rank != 0:
      return None
    a64 = np.concatenate(a_rows).astype(np.float64)
    c_got = np.concatenate(c_rows)
    err, ref_max = 0.0, 0.0
    for c0, b in zip(col_groups, b_cols):
      ref = a64 @ b.astype(np.float64)
      err = max(err, float(np.abs(c_got[:, c0:c0 + b.shape[1]] - ref).max()))
      ref_max = max(ref_max, float(np.abs(ref).max()))
    rel = err / ref_max
    return {'data': label, 'max_rel_err_vs_fp64': rel, 'tolerance': TOL, 'ok': bool(rel <= TOL),
            'sample': '%d rows x %d columns of C (16 row groups, %d column groups over all column blocks) vs float64 '
                      'np.dot on the host' % (a64.shape[0], gw * len(col_groups), len(col_groups))}

  def run_dot(self, out):
    torch, sp, args, ctx = self.torch, self.sp, self.args, self.ctx
    from spartan_b200.expr.base import eval_cache, lazify
    from spartan_b200 import device_ops
    n, tile, world = args.n, args.tile, self.world
    peaks = self.peaks
    sp.FLAGS.dot_precision = args.precision
    A = sp.rand(n, n, seed=0, dtype=np.float32, tile_hint=(tile, tile)).evaluate()
    B = sp.rand(n, n, seed=1, dtype=np.float32, tile_hint=(tile, tile)).evaluate()
    holder = self.holder

    def dot_step(a=A, b=B):
      holder['C'] = sp.dot(lazify(a), lazify(b), tile_hint=(tile, tile)).evaluate()

    sampler = ClockSampler(int(os.environ.get('LOCAL_RANK', 0)))
    sampler.start()                        # nvidia-smi needs ~100 ms to start: launch it before the warm-up
    for _ in range(args.warmup):
      dot_step()
    self.sync()
    launches_w = ctx.kernel_launches
    sampler.mark_begin()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
      dot_step()
    ev1.record()
    self.sync()
    sampler.mark_end()
    clocks = sampler.stop()
    ms = self.maxreduce(ev0.elapsed_time(ev1)) / args.steps
    launches = ctx.kernel_launches - launches_w
    flops = 2.0 * n ** 3
    gflops = flops / ms / 1e6
    tf = gflops / 1e3 / world
    passes = 3 if 'x3' in args.precision else 1
    mma_peak = peaks['bf16_tflops_sustained'] / (2 if args.precision.startswith('tf32') else 1)
    parity = {'uniform': self.dot_parity(A, B, holder['C'], 'uniform[0,1) (the timed inputs)')}
    out.update({
      'metric': 'spartan.dot fp32 GFLOP/s', 'value': gflops, 'unit': 'GFLOP/s', 'n_gpus': world, 'steps': args.steps,
      'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
      'dtype': 'f32', 'data': 'synthetic (device Philox, seeds 0/1)', 'config': workload_config(n, tile),
      'precision': '%s: fp32 operands split into bf16 hi+lo, 3 tensor-core passes, fp32 accumulate; 16 operand mantissa '
                   'bits vs sgemm\'s 24 -- about 4x the error of np.dot, inside the 1e-5 bar (see parity, both datasets); '
                   'tf32x3 timed beside it (dot_variants)' % args.precision if args.precision == 'bf16x3' else args.precision,
      'operand_cache': 'prepared (split / transposed) operands of unchanged arrays are reused across steps '
                       '(FLAGS.dot_prepared_cache); dot_variants.uncached re-prepares every step; the cross-GPU operand '
                       'exchange runs every step either way',
      'roofline': {'bound': 'tensor', 'achieved': tf, 'peak': peaks['bf16_tflops_sustained'], 'unit': 'TFLOP/s',
                   'frac': tf / peaks['bf16_tflops_sustained'],
                   'traffic': measured_traffic('gemm_kernel_bf16x3_dot32768') if (world == 1 and n == 32768 and args.precision == 'bf16x3') else None,
                   'executed_mma_frac': tf * passes / mma_peak,
                   'peak_source': peaks['source'] + ' bf16 sustained (tf32 runs at half the bf16 rate); executed_mma_frac '
                                  'counts the 3 MMA passes of the split modes and can exceed 1: the sustained cuBLAS figure '
                                  'is itself limited by the power cap',
                   'per_gpu': True},
      'clocks': clocks, 'gpu_launches': launches, 'parity': parity})

    variants = {}
    if not args.skip_variants:
      k = max(1, min(args.steps, 5))
      # operands re-prepared every step (no cache)
      sp.FLAGS.dot_prepared_cache = False
      ms_u = self.timed(dot_step, k, 1)
      sp.FLAGS.dot_prepared_cache = True
      variants['uncached'] = {'precision': args.precision, 'ms_per_step': ms_u, 'value': flops / ms_u / 1e6, 'unit': 'GFLOP/s',
                              'steps': k}
      # zero-mean inputs: the dataset that separates the precision modes
      A2 = sp.randn(n, n, seed=10, dtype=np.float32, tile_hint=(tile, tile)).evaluate()
      B2 = sp.randn(n, n, seed=11, dtype=np.float32, tile_hint=(tile, tile)).evaluate()
      dot_step(A2, B2)
      parity['randn'] = self.dot_parity(A2, B2, holder['C'], 'standard normal (zero mean)')
      # tf32x3 beside bf16x3, same inputs, same protocol
      if args.precision != 'tf32x3':
        device_ops.prepared_cache.clear()
        sp.FLAGS.dot_precision = 'tf32x3'
        ms_t = self.timed(dot_step, k, 2)
        tft = flops / ms_t / 1e9 / world
        variants['tf32x3'] = {'precision': 'tf32x3', 'ms_per_step': ms_t, 'value': flops / ms_t / 1e6, 'unit': 'GFLOP/s',
                              'steps': k, 'roofline_frac': tft / peaks['bf16_tflops_sustained'],
                              'executed_mma_frac': tft * 3 / (peaks['bf16_tflops_sustained'] / 2),
                              'parity_uniform': self.dot_parity(A, B, holder['C'], 'uniform[0,1)')}
        dot_step(A2, B2)
        variants['tf32x3']['parity_randn'] = self.dot_parity(A2, B2, holder['C'], 'standard normal (zero mean)')
        sp.FLAGS.dot_precision = args.precision
      del A2, B2
    out['dot_variants'] = variants
    if world > 1 and ctx.peer.available():
      out['peer_memory'] = {'used': True, 'gate_timeouts': ctx.peer.gate_timeouts()}

    # ---------------- e2e: host buffers -> from_numpy (H2D) -> dot -> read-back of C (D2H)
    e2e = None
    if not args.skip_e2e:
      import psutil
      need = 3 * n * n * 4 * world       # every rank holds pinned a, b and a read-back buffer
      fits = torch.tensor([1 if psutil.virtual_memory().available >= 2 * need else 0], device=ctx.device)
      if world > 1:
        self.dist.all_reduce(fits, op=self.dist.ReduceOp.MIN)      # one decision for the whole job
      if not bool(fits.item()):
        e2e = {'skipped': 'host memory: need %.0f GiB pinned across ranks' % (need / 2 ** 30)}
      else:
        device_ops.prepared_cache.clear()
        a_host = torch.empty((n, n), dtype=torch.float32, pin_memory=True)
        b_host = torch.empty((n, n), dtype=torch.float32, pin_memory=True)
        out_host = torch.empty((n, n), dtype=torch.float32, pin_memory=True)
        a_np, b_np, out_np = a_host.numpy(), b_host.numpy(), out_host.numpy()
        # same values as the resident run: every rank needs (and uploads) only its own blocks of the host operands
        A.read_local_into(a_np); B.read_local_into(b_np)
        torch.cuda.synchronize()
        dot_step()
        resident = holder['C']
        out_bytes = [0]

        def e2e_step():
          eval_cache.clear()
          e = sp.dot(sp.from_numpy(a_np, tile_hint=(tile, tile)), sp.from_numpy(b_np, tile_hint=(tile, tile)),
                     tile_hint=(tile, tile))
          c = e.evaluate()
          out_bytes[0] = c.read_local_into(out_np)          # D2H of this rank's share of C into pinned memory
          torch.cuda.current_stream().synchronize()

        steps_e = max(1, min(args.steps, 5))
        ms_e = self.timed(e2e_step, steps_e, 2)
        # the read-back against the resident result (same inputs; the streamed schedule sums K in another order)
        diff = 0.0
        scale = 1.0
        for block in resident.local_blocks()[:4]:
          r0 = block.ul[0]
          sl = (slice(r0, min(block.lr[0], r0 + 8)), slice(block.ul[1], block.lr[1]))
          want = resident.fetch(sp.extent.create((sl[0].start, sl[1].start), (sl[0].stop, sl[1].stop), (n, n))).cpu().numpy()
          diff = max(diff, float(np.abs(out_np[sl] - want).max()))
          scale = max(scale, float(np.abs(want).max()))
        timeline = None
        if world > 1 and os.environ.get('SPARTAN_BENCH_TRACE'):
          sp.FLAGS.dot_trace = True
          eval_cache.clear()
          self.sync()
          c = sp.dot(sp.from_numpy(a_np, tile_hint=(tile, tile)), sp.from_numpy(b_np, tile_hint=(tile, tile)),
                     tile_hint=(tile, tile)).evaluate()
          tr = getattr(c, 'trace', None)
          c.read_local_into(out_np)
          torch.cuda.synchronize()
          mine = tr.report() if tr is not None else None
          sp.FLAGS.dot_trace = False
          gathered = [None] * world
          self.dist.all_gather_object(gathered, mine)
          timeline = {'rank0': gathered[0], 'rank%d' % (world - 1): gathered[-1]}
          del c
        e2e = {'value': flops / ms_e / 1e6, 'unit': 'GFLOP/s', 'ms_per_step': ms_e, 'steps': steps_e, 'timeline_ms': timeline,
               'h2d_bytes_per_step': int(2 * n * n * 4 // world), 'd2h_bytes_per_step': int(out_bytes[0]),
               'max_rel_diff_vs_resident': self.maxreduce(diff / scale),
               'numa_bound_cpus': (len(ctx.numa_cpus) if ctx.numa_cpus else None),
               'note': 'sp.dot(sp.from_numpy(a), sp.from_numpy(b)).evaluate() + read-back of C; pinned host buffers; bytes '
                       'are per rank; upload, contraction and read-back are pipelined strip by strip '
                       '(FLAGS.dot_stream_host_operands)'}
        del a_host, b_host, out_host
    out['e2e'] = e2e
    if self.rank == 0 and world == 1 and not args.skip_cpu:
      out['cpu_baseline'] = self.cpu_dot_baseline(n, tile)
    del A, B
    holder.clear(); eval_cache.clear(); device_ops.prepared_cache.clear(); torch.cuda.empty_cache()

  def cpu_dot_baseline(self, n, tile):
    """Rank 0, N=1: the oracle's tiled dot on a bounded sample (ONE K-join of the reference arm's step)."""
    import threadpoolctl
    cores = _host_cores()
    with threadpoolctl.threadpool_limits(limits=cores, user_api='blas'):
      A, B, C, a_tiles = reference_dot_setup(n, tile)
      from spartan_oracle import expr as oexpr
      oexpr.dot_grid_join_mapper(a_tiles[0], (A, B), C)
      t0 = time.perf_counter()
      for ex in a_tiles[1:3]:
        oexpr.dot_grid_join_mapper(ex, (A, B), C)
      t = time.perf_counter() - t0
      threads = blas_threads()
    k = len(a_tiles[1:3])
    flops = 2.0 * a_tiles[0].shape[0] * a_tiles[0].shape[1] * n * k
    return {'value': flops / t / 1e9, 'unit': 'GFLOP/s', 'cores': threads or cores, 'kind': 'port',
            'sample': '%d K-joins of the oracle tiled dot: (%d x %d).(%d x %d) sgemm + B row-block stitch + np.add merge '
                      'each; OpenBLAS threads = %s' % (k, a_tiles[0].shape[0], a_tiles[0].shape[1], a_tiles[0].shape[1], n, threads)}

  # ---------------------------------------------------------------- map / reduce family
  def run_map_reduce(self, out):
    torch, sp, args, ctx, world = self.torch, self.sp, self.args, self.ctx, self.world
    from spartan_b200.expr.base import eval_cache, lazify
    peaks, holder = self.peaks, self.holder
    total = 1 << args.mr_log2
    cols = 32768
    rows = total // cols
    trow = max(1, rows // 8)
    X = sp.rand(rows, cols, seed=2, dtype=np.float32, tile_hint=(trow, cols)).evaluate()
    Y = sp.rand(rows, cols, seed=3, dtype=np.float32, tile_hint=(trow, cols)).evaluate()
    steps, warm = max(args.steps, 5), max(args.warmup, 3)
    # host samples for parity: 256 columns and 8 rows of both operands, gathered on rank 0
    c0, r0 = 4096 + 17, min(rows - 8, 3 * trow + 5)
    col_ex = sp.extent.create((0, c0), (rows, c0 + 256), (rows, cols))
    row_ex = sp.extent.create((r0, 0), (r0 + 8, cols), (rows, cols))
    xc, yc = self.to_host(X.fetch(col_ex, dst=0)), self.to_host(Y.fetch(col_ex, dst=0))
    xr, yr = self.to_host(X.fetch(row_ex, dst=0)), self.to_host(Y.fetch(row_ex, dst=0))

    def general(x, y):
      return sp.abs(x - y) * x + sp.maximum(y, 0.5)

    def general_np(x, y):
      return np.abs(x - y) * x + np.maximum(y, 0.5)

    cases = [
      ('sum_axis0', '(x*2+y).sum(axis=0)', lambda: (lazify(X) * 2 + lazify(Y)).sum(axis=0).optimized().evaluate(), 8.0,
       'kernel catalogue entry x*c+y'),
      ('general_sum_axis0', '(|x-y|*x + max(y,0.5)).sum(axis=0)',
       lambda: general(lazify(X), lazify(Y)).sum(axis=0).optimized().evaluate(), 8.0, 'chain outside the catalogue: '
       'specialised at run time through NVRTC'),
      ('map', 'x*2+y -> new array', lambda: (lazify(X) * 2 + lazify(Y)).optimized().evaluate(), 12.0, 'pure map'),
      ('sum_axis1', '(x*2+y).sum(axis=1)', lambda: (lazify(X) * 2 + lazify(Y)).sum(axis=1).optimized().evaluate(), 8.0,
       'trailing-axis reduction'),
      ('sum_all', '(x*2+y).sum()', lambda: (lazify(X) * 2 + lazify(Y)).sum().optimized().evaluate(), 8.0,
       'flat reduction'),
    ]
    results = {}
    for name, text, fn, bytes_per_elem, note in cases:
      def eager(fn=fn):
        holder['S'] = fn()
      ms_eager = self.timed(eager, steps, warm)
      ms, launch = ms_eager, 'eager'
      if not args.no_graph:
        rep = sp.replayable(fn)

        def replay(rep=rep):
          holder['S'] = rep()
        ms = self.timed(replay, steps, warm)
        launch = 'cuda-graph replay of evaluate() (%d library kernels per step)' % rep.kernel_launches
      nbytes = bytes_per_elem * total
      gbs = nbytes / ms / 1e6
      S = holder['S']
      # parity (every N): against float64 (reductions) / the exact fp32 chain (map) on the gathered samples
      err = None
      if name in ('sum_axis0', 'general_sum_axis0'):
        got = S.glom()
        if self.rank == 0:
          f = general_np if name.startswith('general') else (lambda a, b: a * 2 + b)
          ref = f(xc.astype(np.float64), yc.astype(np.float64)).sum(axis=0)
          err = float(np.abs(got[c0:c0 + 256] - ref).max() / np.abs(ref).max())
        if name == 'sum_axis0':
          holder['colsum64'] = float(got.astype(np.float64).sum())
      elif name == 'sum_axis1':
        got = S.glom()
        if self.rank == 0:
          ref = (xr.astype(np.float64) * 2 + yr).sum(axis=1)
          err = float(np.abs(got[r0:r0 + 8] - ref).max() / np.abs(ref).max())
      elif name == 'sum_all':
        got = float(S.glom())
        if self.rank == 0 and 'colsum64' in holder:      # a checksum of checksums: sum() == sum of the column sums
          err = abs(got - holder['colsum64']) / abs(holder['colsum64'])
      else:
        got = self.to_host(S.fetch(row_ex, dst=0))
        if self.rank == 0:
          err = 0.0 if np.array_equal(got, xr * np.float32(2) + yr) else float(np.abs(got - (xr * np.float32(2) + yr)).max())
      results[name] = {'expr': text, 'note': note, 'value': gbs, 'unit': 'GB/s', 'ms_per_step': ms, 'launch': launch,
                       'ms_per_step_eager': ms_eager, 'value_eager': nbytes / ms_eager / 1e6,
                       'algorithmic_bytes': nbytes, 'frac_of_hbm_per_gpu': gbs / world / peaks['hbm_gbs'],
                       'max_rel_err': err, 'ok': (None if err is None else bool(err <= TOL))}
      holder.pop('S', None)
      rep = None
      eval_cache.clear()
    head = results['sum_axis0']
    mr = {'metric': 'fused (x*2+y).sum(axis=0) GB/s', 'value': head['value'], 'unit': 'GB/s', 'ms_per_step': head['ms_per_step'],
          'launch': head['launch'], 'ms_per_step_eager': head['ms_per_step_eager'], 'value_eager': head['value_eager'],
          'elements': total, 'algorithmic_bytes': head['algorithmic_bytes'],
          'roofline': {'bound': 'hbm', 'achieved': head['value'] / world, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                       'frac': head['value'] / world / peaks['hbm_gbs'],
                       'traffic': measured_traffic('direct_reduce_kernel_mapreduce_2p30') if (world == 1 and args.mr_log2 == 30) else None,
                       'peak_source': peaks['source'] + ' copy bandwidth (read+write); a read-only stream can exceed it'},
          'max_rel_err_vs_fp64': head['max_rel_err'], 'tolerance': TOL, 'variants': results}
    if self.rank == 0 and world == 1 and not args.skip_cpu:
      mr['cpu_baseline'] = cpu_map_reduce_baseline()
    out['map_reduce'] = mr
    del X, Y
    holder.clear(); eval_cache.clear(); torch.cuda.empty_cache()

  # ---------------------------------------------------------------- k-means (configs[3])
  def run_kmeans(self, out):
    torch, sp, args, ctx, world, rank = self.torch, self.sp, self.args, self.ctx, self.world, self.rank
    from spartan_b200 import device_ops
    from spartan_b200.expr.base import eval_cache
    n, d, k = args.km_n, 256, 1024
    n = max(8 * 128, n // 8 * 8)
    tile_rows = n // 8
    X = sp.rand(n, d, seed=4, dtype=np.float32, tile_hint=(tile_rows, d)).evaluate()
    c0 = X.fetch(sp.extent.create((0, 0), (k, d), (n, d)), dst=0)
    c0 = c0.contiguous() if rank == 0 else torch.empty((k, d), dtype=torch.float32, device=ctx.device)
    self.comm.broadcast(c0, 0)
    c0 = c0.cpu().numpy()
    iters = 5                                                # iterations per fit() in the timed steps
    res = {}

    def step(n_iter=iters):
      res['centers'], res['labels'] = sp.KMeans(n_clusters=k, n_iter=n_iter).fit(X, centers=c0)

    device_ops.prepared_cache.clear()
    step(1)                                                  # sizes scratch, creates communicators
    device_ops.prepared_cache.clear()
    ms_first = self.timed(lambda: step(1), 1, 0)             # a one-iteration fit incl. laying the points out for the tensor cores
    steps = max(2, min(self.args.steps, 5))
    launches0 = ctx.kernel_launches
    ms = self.timed(step, steps, 1) / iters
    launches = (ctx.kernel_launches - launches0) // ((steps + 1) * iters)
    step(1)                                                  # parity below: the labels / centres of ONE iteration from c0
    flops = 2.0 * n * d * k
    tf = flops / ms / 1e9
    # parity at every N: labels of 2048 sampled points of every rank against float64 distances on the host; the counts
    # and the per-centre sums through conserved totals (sum of counts = n; column sums of the centre sums = column sums
    # of X, computed by the fused reduce kernel)
    labels = res['labels']
    bad = torch.zeros(2, dtype=torch.float64, device=ctx.device)
    for block in X.local_blocks()[:1]:
      r0 = block.ul[0]
      m = min(2048, block.lr[0] - r0)
      xs = X.fetch(sp.extent.create((r0, 0), (r0 + m, d), (n, d))).cpu().numpy().astype(np.float64)
      lab = labels.fetch(sp.extent.create((r0,), (r0 + m,), (n,))).cpu().numpy()
      c64 = c0.astype(np.float64)
      d2 = (xs * xs).sum(1)[:, None] - 2.0 * xs @ c64.T + (c64 * c64).sum(1)[None, :]
      ref = d2.argmin(1)
      wrong = lab != ref
      # a disagreement only counts when the two candidates are not a near tie in float64
      gap = d2[np.arange(m), lab.clip(0, k - 1)] - d2[np.arange(m), ref]
      bad[0] += float(np.count_nonzero(wrong & (gap > 1e-5 * np.abs(d2[np.arange(m), ref]))))
      bad[1] += m
    if world > 1:
      self.dist.all_reduce(bad)
    # the accumulate + all-reduce + divide: 8 sampled centres recomputed from the labels in float64 (torch as the checker)
    sample_c = [int(j) for j in np.random.default_rng(5).integers(0, k, 8)]
    acc = torch.zeros((len(sample_c), d + 1), dtype=torch.float64, device=ctx.device)
    for block in X.local_blocks():
      xb = X.fetch(block)
      lb = labels.fetch(sp.extent.create((block.ul[0],), (block.lr[0],), (n,)))
      for i, j in enumerate(sample_c):
        sel = lb == j
        acc[i, :d] += xb[sel].to(torch.float64).sum(0)
        acc[i, d] += sel.sum()
    if world > 1:
      self.dist.all_reduce(acc)
    parity = None
    if rank == 0:
      acc = acc.cpu().numpy()
      cerr = 0.0
      for i, j in enumerate(sample_c):
        if acc[i, d] > 0:
          want = acc[i, :d] / acc[i, d]
          cerr = max(cerr, float(np.abs(res['centers'][j] - want).max() / np.abs(want).max()))
      parity = {'labels_checked': int(bad[1].item()), 'labels_wrong': int(bad[0].item()),
                'centres_checked': len(sample_c), 'centres_max_rel_err_vs_fp64': cerr, 'tolerance': TOL,
                'ok': bool(bad[0].item() == 0 and cerr <= TOL)}
    km_out = {'config': 'k-means %d x %d fp32, k=%d: one iteration = assign (distance GEMM + argmin) + accumulate + '
                        'all-reduce of sums/counts + centre update on the host' % (n, d, k),
              'n_gpus': world, 'ms_per_iter': ms, 'ms_first_iter_incl_point_preparation': ms_first, 'steps': steps,
              'iterations_per_step': iters,
              'value': tf, 'unit': 'TFLOP/s (2*n*d*k distance flops)', 'points_gbs': n * d * 4 / ms / 1e6,
              'gpu_launches_per_iter': launches,
              'roofline': {'bound': 'tensor', 'achieved': tf / world, 'peak': self.peaks['bf16_tflops_sustained'],
                           'unit': 'TFLOP/s', 'frac': tf / world / self.peaks['bf16_tflops_sustained'],
                           'executed_mma_frac': 3 * tf / world / self.peaks['bf16_tflops_sustained'], 'per_gpu': True,
                           'traffic': measured_traffic('kmeans_iteration_10Mx256')},
              'parity': parity}
    if rank == 0 and world == 1 and not args.skip_cpu:
      from spartan_oracle import apps
      ms_ = 4000
      xs = X.fetch(sp.extent.create((0, 0), (ms_, d), (n, d))).cpu().numpy()
      t0 = time.perf_counter(); ref = apps.kmeans_dist_mapper(xs, c0); t1 = time.perf_counter()
      lab = labels.fetch(sp.extent.create((0,), (ms_,), (n,))).cpu().numpy()
      km_out['cpu_baseline'] = {'kind': 'port', 'cores': 1, 'unit': 'TFLOP/s', 'value': 2.0 * ms_ * d * k / (t1 - t0) / 1e12,
                                'sample': 'oracle kmeans_dist_mapper (scipy cdist + argmin, k_means_.py:61-66) on %d points; '
                                          'labels agree with the device on %.4f of them' % (ms_, float((lab == ref).mean()))}
    out['kmeans'] = km_out
    del X, labels
    res.clear(); eval_cache.clear(); device_ops.prepared_cache.clear(); torch.cuda.empty_cache()

  # ---------------------------------------------------------------- PageRank SpMV (configs[4])
  def run_spmv(self, out):
    torch, sp, args, ctx, world, rank = self.torch, self.sp, self.args, self.ctx, self.world, self.rank
    from spartan_b200.examples import pagerank
    from spartan_b200.expr.base import eval_cache, lazify
    N = max(8 * 1024, args.pr_n // 8 * 8)
    strip = N // 8
    wts = pagerank.make_weights(N, strip, seed=5)
    p = sp.rand(N, 1, seed=6, dtype=np.float32, tile_hint=(strip, 1)).evaluate()
    holder = self.holder

    def ev():
      return sp.dot(wts, lazify(p)).evaluate()

    def eager():
      holder['y'] = ev()
    steps, warm = max(args.steps, 5), max(args.warmup, 3)
    ms_eager = self.timed(eager, steps, warm)
    ms, launch = ms_eager, 'eager'
    if not args.no_graph:
      rep = sp.replayable(ev)

      def replay():
        holder['y'] = rep()
      ms = self.timed(replay, steps, warm)
      launch = 'cuda-graph replay of evaluate() (%d library kernels per step)' % rep.kernel_launches
    nnz = wts.val.nnz
    bytes_alg = 8.0 * nnz + 4.0 * (N + 1) + 4.0 * N + 4.0 * N
    gbs = bytes_alg / ms / 1e6
    # parity at every N: 4096 sampled rows; each rank sums its strips' contributions in float64 on the host
    y = holder['y'].glom().reshape(-1)
    rng = np.random.default_rng(77)
    rows = np.unique(rng.integers(0, N, 4096))
    part = np.zeros(rows.shape[0], dtype=np.float64)
    rows_t = torch.from_numpy(rows).to(ctx.device)
    for c0, c1, owner, dev, nnz_b in wts.val.blocks:
      if owner != rank:
        continue
      rowptr, col, val = dev
      lo = rowptr[rows_t].to(torch.int64).cpu().numpy(); hi = rowptr[rows_t + 1].to(torch.int64).cpu().numpy()
      xs = p.fetch(sp.extent.create((c0, 0), (c1, 1), (N, 1))).reshape(-1)
      lens = hi - lo
      if lens.sum():
        idx = np.concatenate([np.arange(a, b) for a, b in zip(lo, hi)])
        idx_t = torch.from_numpy(idx).to(ctx.device)
        contrib = (val[idx_t].to(torch.float64) * xs[col[idx_t].to(torch.int64)].to(torch.float64)).cpu().numpy()
        np.add.at(part, np.repeat(np.arange(rows.shape[0]), lens), contrib)
    pt = torch.from_numpy(part).to(ctx.device)
    if world > 1:
      self.dist.all_reduce(pt)
    ref = pt.cpu().numpy()
    err = float(np.abs(y[rows] - ref).max() / max(1e-30, np.abs(ref).max()))
    sp_out = {'config': 'PageRank SpMV: %d x %d fp32, %d outlinks/page (nnz = %d), column strips of %d (reference layout), '
                        'y all-reduced' % (N, N, pagerank.OUTLINKS_PER_PAGE, nnz, strip),
              'n_gpus': world, 'ms_per_step': ms, 'launch': launch, 'ms_per_step_eager': ms_eager, 'value': gbs, 'unit': 'GB/s',
              'algorithmic_bytes': bytes_alg,
              'roofline': {'bound': 'hbm', 'achieved': gbs / world, 'peak': self.peaks['hbm_gbs'], 'unit': 'GB/s',
                           'frac': gbs / world / self.peaks['hbm_gbs'], 'per_gpu': True,
                           'traffic': measured_traffic('spmv_csr_10M')},
              'parity': {'max_rel_err_vs_fp64': err, 'tolerance': TOL, 'ok': bool(err <= TOL),
                         'sample': '%d random rows, contributions of every strip summed in float64' % rows.shape[0]}}
    if rank == 0 and world == 1 and not args.skip_cpu:
      import scipy.sparse
      from spartan_oracle import apps
      c0, c1, owner, dev, nnz_b = wts.val.blocks[0]
      rowptr, col, val = dev
      # the first two reference strips as a host matrix (bounded sample), multiplied the reference's way
      keep = col < 2 * strip
      counts = (rowptr[1:] - rowptr[:-1]).to(torch.int64)
      rr = torch.repeat_interleave(torch.arange(N, device=ctx.device), counts)[keep].cpu().numpy()
      m = scipy.sparse.coo_matrix((val[keep].cpu().numpy(), (rr, col[keep].cpu().numpy())), shape=(N, 2 * strip)).tocsc()
      xv = p.fetch(sp.extent.create((0, 0), (2 * strip, 1), (N, 1))).reshape(-1).cpu().numpy()
      t0 = time.perf_counter(); apps.spmv_strips(m, xv, strip); t1 = time.perf_counter()
      b = 8.0 * m.nnz + 12.0 * N * 2
      sp_out['cpu_baseline'] = {'kind': 'port', 'cores': 1, 'value': b / (t1 - t0) / 1e9, 'unit': 'GB/s',
                                'sample': 'oracle spmv_strips (per strip tocsr().dot + np.add merge, dot.py:213-217) on the '
                                          'first 2 of 8 column strips (%d non-zeros)' % m.nnz}
    out['spmv'] = sp_out
    holder.clear(); eval_cache.clear(); torch.cuda.empty_cache()


def run_b200(args):
  b = Bench(args)
  torch, dist = b.torch, b.dist
  out = {}
  if args.skip_dot:
    out.update({'metric': 'spartan.dot fp32 GFLOP/s', 'value': None, 'unit': 'GFLOP/s', 'n_gpus': b.world,
                'steps': args.steps, 'warmup': args.warmup, 'skipped': 'dot (tuning run)'})
  else:
    b.run_dot(out)
  if not args.skip_mapreduce:
    b.run_map_reduce(out)
  if not args.skip_apps:
    b.run_kmeans(out)
    b.run_spmv(out)
  if b.rank == 0:
    print(json.dumps(out), flush=True)
  if b.world > 1:
    # captured NCCL work must be released before the communicator is torn down (destroy would wait on it forever)
    b.holder.clear()
    import gc
    gc.collect()
    torch.cuda.synchronize()
    b.comm.barrier()
    b.sp.shutdown()
    dist.destroy_process_group()


if __name__ == '__main__':
  if os.environ.get('SP_BENCH_WATCHDOG'):          # debugging aid: dump every thread's stack and exit if the run hangs
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ['SP_BENCH_WATCHDOG']), exit=True)
  a = parse_args()
  if a.impl == 'reference':
    run_reference(a)
  else:
    run_b200(a)
