#!/usr/bin/env python
"""bench.py -- the hot path of BASELINE.json on N B200s of one node.

  python bench.py --gpus 1 --steps 5 --warmup 3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference ...        # the reference's CPU path (NumPy oracle) on the host cores

Headline workload (BASELINE.json configs[1]): spartan.dot of two 32768x32768 fp32 tiled arrays,
tile_hint=(4096,4096); metric = GFLOP/s with 2*N^3 algorithmic flops.  A "step" is one evaluation of the
DotExpr over inputs already resident in HBM (``value``) or starting from host buffers through the public
API with the H2D/D2H copies inside the timed region (``e2e``).  The fused map+reduce workload
(configs[2], (x*2+y).sum(axis=0) over 2^30 fp32 elements) is measured in the same run and reported under
``map_reduce`` with its own HBM roofline.  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
  if p not in sys.path:
    sys.path.insert(0, p)


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
  ap.add_argument('--skip-e2e', action='store_true')
  ap.add_argument('--skip-dot', action='store_true', help='tuning aid: only the map+reduce workload')
  ap.add_argument('--skip-mapreduce', action='store_true')
  ap.add_argument('--skip-cpu', action='store_true')
  ap.add_argument('--no-graph', action='store_true', help='time the map+reduce step eagerly only')
  return ap.parse_args()


def measured_traffic(key):
  path = os.path.join(ROOT, 'profiles', 'r01_traffic.json')
  try:
    with open(path) as f:
      return json.load(f).get(key)
  except Exception:
    return None


def measured_peaks():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(path):
    with open(path) as f:
      d = json.load(f)
    return {'hbm_gbs': d['hbm_gbs'], 'bf16_tflops': d['bf16_tflops'],
            'bf16_tflops_sustained': d.get('bf16_tflops_sustained', d['bf16_tflops']), 'source': 'measured'}
  return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'source': 'fallback'}


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
def run_reference(args):
  """The reference's CPU implementation of the path = its per-tile NumPy calls (np.dot -> BLAS sgemm, all
  host threads), restated by the oracle; a step is a bounded sample of the workload: one 4096-row strip of C
  (dot_map2_mapper on one A row strip, spartan/expr/dot.py:195-217)."""
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  n, tile = args.n, args.tile
  rows = min(tile, n)
  rng = np.random.default_rng(0)
  a = rng.random((rows, n), dtype=np.float32)
  b = rng.random((n, min(n, tile)), dtype=np.float32)
  cores = os.cpu_count() or 1
  flops = 2.0 * rows * n * b.shape[1]
  for _ in range(max(1, min(args.warmup, 1))):
    np.dot(a, b)
  times = []
  for _ in range(max(1, min(args.steps, 3))):
    t0 = time.perf_counter()
    np.dot(a, b)
    times.append(time.perf_counter() - t0)
  t = float(np.median(times))
  val = flops / t / 1e9
  out = {'metric': 'spartan.dot fp32 GFLOP/s', 'value': val, 'unit': 'GFLOP/s', 'n_gpus': args.gpus,
         'steps': len(times), 'warmup': 1, 'ms_per_step': t * 1e3, 'higher_is_better': True, 'scaling': 'strong',
         'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'impl': 'reference',
         'config': {'workload': 'spartan.dot %dx%d fp32, tile_hint=(%d,%d)' % (n, n, tile, tile)},
         'cpu_baseline': {'value': val, 'unit': 'GFLOP/s', 'cores': cores, 'kind': 'port',
                          'sample': 'np.dot of one C strip: (%d x %d) . (%d x %d) fp32, NumPy/BLAS, all host threads'
                                    % (rows, n, n, b.shape[1])},
         'e2e': {'value': val, 'unit': 'GFLOP/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
         'gpu_launches': 0}
  print(json.dumps(out), flush=True)


def cpu_map_reduce_baseline(log2_elems=27, tiles=8):
  """The reference's CPU path for configs[2] on a bounded sample: per row tile the NumPy chain of tile_mapper +
  _reduce_mapper ((x*2) -> (+y) -> .sum(axis=0), one temporary per ufunc; map.py:48-88, reduce.py:21-70), tiles evaluated
  by a pool of `tiles` threads (emulates num_workers=8; NumPy releases the GIL), partials merged with np.add
  (tile.pyx:263-283)."""
  from concurrent.futures import ThreadPoolExecutor
  cols = 32768
  rows = (1 << log2_elems) // cols
  rng = np.random.default_rng(2)
  x = rng.random((rows, cols), dtype=np.float32); y = rng.random((rows, cols), dtype=np.float32)
  trow = rows // tiles

  def tile_fn(i):
    xs, ys = x[i * trow:(i + 1) * trow], y[i * trow:(i + 1) * trow]
    return (xs * np.float32(2) + ys).sum(axis=0)

  def run(pool):
    parts = list(pool.map(tile_fn, range(tiles))) if pool else [tile_fn(i) for i in range(tiles)]
    out = parts[0]
    for p in parts[1:]:
      out = np.add(out, p)
    return out
  nbytes = 2.0 * 4.0 * rows * cols
  t0 = time.perf_counter(); run(None); t_single = time.perf_counter() - t0
  with ThreadPoolExecutor(tiles) as pool:
    run(pool)
    t0 = time.perf_counter(); run(pool); t_pool = time.perf_counter() - t0
  return {'value': nbytes / t_pool / 1e9, 'unit': 'GB/s', 'cores': min(tiles, os.cpu_count() or 1), 'kind': 'port',
          'single_thread_gbs': nbytes / t_single / 1e9,
          'sample': '(x*2+y).sum(axis=0) over 2^%d fp32 elements in %d row tiles, NumPy ufunc chain per tile, '
                    '%d tile threads' % (log2_elems, tiles, tiles)}


# ----------------------------------------------------------------------------------------------- b200 arm
def timed(fn, steps, warmup, sync, maxreduce):
  import torch
  for _ in range(warmup):
    fn()
  sync()
  ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  ev0.record()
  for _ in range(steps):
    fn()
  ev1.record()
  sync()
  return maxreduce(ev0.elapsed_time(ev1)) / steps


def run_b200(args):
  import torch
  import torch.distributed as dist
  import spartan_b200 as sp
  from spartan_b200 import comm
  from spartan_b200.expr.base import eval_cache, lazify

  ctx = sp.initialize()
  rank, world = ctx.worker_id, ctx.num_workers
  assert world == args.gpus or world == 1, 'launch with torchrun for --gpus %d' % args.gpus
  peaks = measured_peaks()
  n, tile = args.n, args.tile
  sp.FLAGS.dot_precision = args.precision

  def sync():
    comm.barrier()
    torch.cuda.synchronize()

  def maxreduce(ms):
    if world == 1:
      return ms
    t = torch.tensor([ms], device=ctx.device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

  # ---------------- inputs, generated on the device (Philox) and kept resident in HBM
  if args.skip_dot:
    n = args.n = 2048
    args.skip_e2e = True
  A = sp.rand(n, n, seed=0, dtype=np.float32, tile_hint=(tile, tile)).evaluate()
  B = sp.rand(n, n, seed=1, dtype=np.float32, tile_hint=(tile, tile)).evaluate()
  holder = {}

  def dot_step():
    e = sp.dot(lazify(A), lazify(B), tile_hint=(tile, tile))
    holder['C'] = e.evaluate()

  sampler = ClockSampler(int(os.environ.get('LOCAL_RANK', 0)))
  sampler.start()                        # nvidia-smi needs ~100 ms to start: launch it before the warm-up
  launches0 = ctx.kernel_launches
  for _ in range(args.warmup):
    dot_step()
  sync()
  launches_w = ctx.kernel_launches
  sampler.mark_begin()
  ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  ev0.record()
  for _ in range(args.steps):
    dot_step()
  ev1.record()
  sync()
  sampler.mark_end()
  clocks = sampler.stop()
  ms = maxreduce(ev0.elapsed_time(ev1)) / args.steps
  launches = (ctx.kernel_launches - launches_w)
  flops = 2.0 * n ** 3
  gflops = flops / ms / 1e6

  # ---------------- parity on the same inputs (outside the timed region): C[rows] vs float64 np.dot
  parity = None
  if world == 1:
    rows = 8
    a_rows = A.fetch(sp.extent.create((0, 0), (rows, n), (n, n))).cpu().numpy()
    b_all = B.glom() if n <= 8192 else None
    C = holder['C']
    c_rows = C.fetch(sp.extent.create((0, 0), (rows, n), (n, n))).cpu().numpy()
    if b_all is None:
      cols = 512
      b_cols = B.fetch(sp.extent.create((0, 0), (n, cols), (n, n))).cpu().numpy()
      ref = np.dot(a_rows.astype(np.float64), b_cols.astype(np.float64))
      c_rows = c_rows[:, :cols]
    else:
      ref = np.dot(a_rows.astype(np.float64), b_all.astype(np.float64))
    err = float(np.abs(c_rows - ref).max() / np.abs(ref).max())
    parity = {'max_rel_err_vs_fp64': err, 'tolerance': 1e-5, 'ok': bool(err <= 1e-5),
              'sample': 'first %d rows of C vs float64 np.dot' % rows}

  # ---------------- e2e: host buffers -> from_numpy (H2D) -> dot -> glom of a C strip (D2H)
  e2e = None
  if not args.skip_e2e:
    import psutil
    ne = n
    need = 3 * ne * ne * 4 * world       # every rank holds pinned a, b and a read-back buffer
    if psutil.virtual_memory().available < 2 * need:
      e2e = {'skipped': 'host memory: need %.0f GiB pinned across ranks' % (need / 2 ** 30)}
    else:
      a_host = torch.empty((ne, ne), dtype=torch.float32, pin_memory=True)
      b_host = torch.empty((ne, ne), dtype=torch.float32, pin_memory=True)
      a_host.uniform_(0, 1); b_host.uniform_(0, 1)
      a_np, b_np = a_host.numpy(), b_host.numpy()
      out_host = torch.empty((ne, ne), dtype=torch.float32, pin_memory=True)
      out_np = out_host.numpy()
      out_bytes = [0]

      def e2e_step():
        eval_cache.clear()
        e = sp.dot(sp.from_numpy(a_np, tile_hint=(tile, tile)), sp.from_numpy(b_np, tile_hint=(tile, tile)),
                   tile_hint=(tile, tile))
        c = e.evaluate()
        out_bytes[0] = c.read_local_into(out_np)          # D2H of this rank's share of C into pinned memory
        torch.cuda.current_stream().synchronize()

      steps_e = max(1, min(args.steps, 3))
      ms_e = timed(e2e_step, steps_e, 1, sync, maxreduce)
      e2e = {'value': flops / ms_e / 1e6, 'unit': 'GFLOP/s', 'ms_per_step': ms_e,
             'h2d_bytes_per_step': int(2 * ne * ne * 4 // world), 'd2h_bytes_per_step': int(out_bytes[0]),
             'note': 'sp.dot(sp.from_numpy(a), sp.from_numpy(b)).evaluate() + read-back of C; pinned host buffers; '
                     'bytes are per rank; upload, contraction and read-back are pipelined strip by strip '
                     '(FLAGS.dot_stream_host_operands)'}
      del a_host, b_host, out_host

  # ---------------- fused map+reduce workload (configs[2])
  mr = None
  if not args.skip_mapreduce:
    holder.clear(); eval_cache.clear(); torch.cuda.empty_cache()
    total = 1 << args.mr_log2
    cols = 32768
    rows = total // cols
    trow = max(1, rows // 8)
    X = sp.rand(rows, cols, seed=2, dtype=np.float32, tile_hint=(trow, cols)).evaluate()
    Y = sp.rand(rows, cols, seed=3, dtype=np.float32, tile_hint=(trow, cols)).evaluate()

    general = os.environ.get('SPARTAN_MR_EXPR') == 'general'    # tuning aid: a chain outside the static catalogue

    def mr_eval():
      if general:
        x, y = lazify(X), lazify(Y)
        e = (sp.abs(x - y) * x + sp.maximum(y, 0.5)).sum(axis=0).optimized()
      else:
        e = (lazify(X) * 2 + lazify(Y)).sum(axis=0).optimized()
      return e.evaluate()

    def mr_step_eager():
      holder['S'] = mr_eval()

    # The step is launched the way an iterative driver would launch it: the whole evaluate() -- fused kernel,
    # finalisation, ncclAllReduce -- captured once as a CUDA graph (sp.replayable) and replayed; the eager number
    # (Python host walking the DAG every step) is reported next to it.
    ms_eager = timed(mr_step_eager, max(args.steps, 5), max(args.warmup, 3), sync, maxreduce)
    launch = 'eager'
    ms_mr = ms_eager
    if not args.no_graph:
      holder['rep'] = sp.replayable(mr_eval)

      def mr_step():
        holder['S'] = holder['rep']()
      ms_mr = timed(mr_step, max(args.steps, 5), max(args.warmup, 3), sync, maxreduce)
      launch = 'cuda-graph replay of evaluate() (%d library kernels per step)' % holder['rep'].kernel_launches
    bytes_alg = 2.0 * 4.0 * total
    gbs = bytes_alg / ms_mr / 1e6
    got = holder['S'].glom()
    xs = X.fetch(sp.extent.create((0, 0), (rows, 256), (rows, cols)), dst=None) if world == 1 else None
    mr_par = None
    if xs is not None:
      ys = Y.fetch(sp.extent.create((0, 0), (rows, 256), (rows, cols)))
      xh, yh = xs.cpu().numpy().astype(np.float64), ys.cpu().numpy().astype(np.float64)
      ref = (np.abs(xh - yh) * xh + np.maximum(yh, 0.5)).sum(axis=0) if general else (xh * 2 + yh).sum(axis=0)
      mr_par = float(np.abs(got[:256] - ref).max() / np.abs(ref).max())
    mr = {'metric': 'fused (x*2+y).sum(axis=0) GB/s', 'value': gbs, 'unit': 'GB/s', 'ms_per_step': ms_mr,
          'launch': launch, 'ms_per_step_eager': ms_eager, 'value_eager': bytes_alg / ms_eager / 1e6,
          'elements': total, 'algorithmic_bytes': bytes_alg,
          'roofline': {'bound': 'hbm', 'achieved': gbs / world, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                       'frac': gbs / world / peaks['hbm_gbs'],
                       'traffic': measured_traffic('stream_kernel_mapreduce_2p30') if (world == 1 and args.mr_log2 == 30) else None,
                       'peak_source': peaks['source'] + ' copy bandwidth (read+write); a read-only stream can exceed it'},
          'max_rel_err_vs_fp64': mr_par}
    holder.pop('rep', None)
    if rank == 0 and world == 1 and not args.skip_cpu:
      mr['cpu_baseline'] = cpu_map_reduce_baseline()

  # ---------------- CPU baseline (rank 0, N=1): the oracle's np.dot on a bounded sample
  cpu = None
  if rank == 0 and world == 1 and not args.skip_cpu:
    rows = min(tile, n)
    rng = np.random.default_rng(0)
    a = rng.random((rows, n), dtype=np.float32); b = rng.random((n, min(n, tile)), dtype=np.float32)
    np.dot(a[:256], b)
    t0 = time.perf_counter(); np.dot(a, b); t = time.perf_counter() - t0
    cpu = {'value': 2.0 * rows * n * b.shape[1] / t / 1e9, 'unit': 'GFLOP/s', 'cores': os.cpu_count() or 1,
           'kind': 'port', 'sample': 'np.dot (%d x %d).(%d x %d) fp32 = one C tile-strip, NumPy/BLAS all host threads'
                                     % (rows, n, n, b.shape[1])}

  if rank == 0:
    tf = gflops / 1e3 / world
    out = {'metric': 'spartan.dot fp32 GFLOP/s', 'value': gflops, 'unit': 'GFLOP/s', 'n_gpus': world,
           'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True,
           'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic (device Philox, seeds 0/1)',
           'config': {'workload': 'spartan.dot %dx%d fp32, tile_hint=(%d,%d)' % (n, n, tile, tile),
                      'precision': args.precision, 'l2': 'inputs (%.1f GiB) larger than L2' % (2 * n * n * 4 / 2 ** 30)},
           'roofline': {'bound': 'tensor', 'achieved': tf, 'peak': peaks['bf16_tflops_sustained'] / 1e0,
                        'unit': 'TFLOP/s', 'frac': tf / peaks['bf16_tflops_sustained'],
                        'traffic': measured_traffic('gemm_kernel_bf16x3_dot32768') if (world == 1 and n == 32768 and args.precision == 'bf16x3') else None,
                        'executed_mma_frac': tf * (3 if 'x3' in args.precision else 1) / (peaks['bf16_tflops_sustained'] / (2 if args.precision.startswith('tf32') else 1)),
                        'peak_source': peaks['source'] + ' bf16 sustained (tf32 runs at half the bf16 rate); executed_mma_frac '
                                       'counts the 3 MMA passes of the split modes and can exceed 1: the sustained cuBLAS '
                                       'figure is itself limited by the power cap',
                        'per_gpu': True},
           'clocks': clocks, 'gpu_launches': launches, 'parity': parity, 'e2e': e2e, 'map_reduce': mr,
           'cpu_baseline': cpu}
    print(json.dumps(out), flush=True)
  if world > 1:
    # captured NCCL work must be released before the communicator is torn down (destroy would wait on it forever)
    holder.clear()
    import gc
    gc.collect()
    torch.cuda.synchronize()
    comm.barrier()
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
