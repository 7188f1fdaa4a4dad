"""Host cost of one eager evaluate(): the Python path from expression construction to the kernel launches, timed WITHOUT
a GPU.  The context poses as rank 0 of `--world` ranks on CPU tensors, every `sp_*` entry point that would launch is
replaced by a stub that returns success, and the collective is a no-op -- so what is left is exactly the host work an
un-captured caller pays per evaluate().  Not a bench: no number from here is a device result.

  python tools/host_profile.py [--world 8] [--iters 2000] [--case sum_axis0] [--profile]
"""
import argparse
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--world', type=int, default=8)
  ap.add_argument('--iters', type=int, default=2000)
  ap.add_argument('--case', default='sum_axis0')
  ap.add_argument('--profile', action='store_true')
  ap.add_argument('--repeat', type=int, default=5)
  ap.add_argument('--top', type=int, default=35)
  args = ap.parse_args()

  import spartan_b200 as sp
  from spartan_b200 import blob_ctx, device_ops, comm, _lib
  from spartan_b200.expr.base import lazify, eval_cache

  class _Stub(object):
    """Counts calls; everything that would launch answers 0 (success) / 0 bytes of scratch."""

    def __init__(self, real):
      self.real, self.calls = real, {}

    def __getattr__(self, name):
      fn = getattr(self.real, name)
      if name.startswith('sp_extent') or name in ('sp_good_tile_shape', 'sp_compute_extents', 'sp_gemm_kpad',
                                                   'sp_last_error', 'sp_version'):
        return fn

      def stub(*a, **k):
        self.calls[name] = self.calls.get(name, 0) + 1
        return 0
      return stub

  real = _lib.lib
  stub = _Stub(real)
  for mod in list(sys.modules.values()):
    if getattr(mod, '__name__', '').startswith('spartan_b200') and getattr(mod, 'lib', None) is real:
      mod.lib = stub
  device_ops._require_cuda = lambda *t: None
  comm.allreduce = lambda t, op: t
  blob_ctx.BlobCtx.stream_ptr = lambda self: 0
  blob_ctx.set(blob_ctx.BlobCtx(0, args.world, torch.device('cpu')))

  rows, cols = 32768 // 8, 1024            # host cost does not depend on the extent sizes, only on the tile counts
  trow = rows // 8
  X = sp.rand(rows, cols, seed=2, dtype=np.float32, tile_hint=(trow, cols)).evaluate()
  Y = sp.rand(rows, cols, seed=3, dtype=np.float32, tile_hint=(trow, cols)).evaluate()

  def general(x, y):
    return sp.abs(x - y) * x + sp.maximum(y, 0.5)

  cases = {
    'sum_axis0': lambda: (lazify(X) * 2 + lazify(Y)).sum(axis=0).optimized().evaluate(),
    'general_sum_axis0': lambda: general(lazify(X), lazify(Y)).sum(axis=0).optimized().evaluate(),
    'map': lambda: (lazify(X) * 2 + lazify(Y)).optimized().evaluate(),
    'sum_axis1': lambda: (lazify(X) * 2 + lazify(Y)).sum(axis=1).optimized().evaluate(),
    'sum_all': lambda: (lazify(X) * 2 + lazify(Y)).sum().optimized().evaluate(),
  }
  names = list(cases) if args.case == 'all' else [args.case]
  for name in names:
    fn = cases[name]
    for _ in range(50):
      fn()
    stub.calls.clear()
    if args.profile:
      pr = cProfile.Profile()
      pr.enable()
    dt = float('inf')
    for _ in range(1 if args.profile else args.repeat):      # best of `repeat` blocks: the container's cores are shared
      t0 = time.process_time()
      for _ in range(args.iters):
        fn()
      dt = min(dt, time.process_time() - t0)
    calls = dict((k, v / args.iters / (1 if args.profile else args.repeat)) for k, v in sorted(stub.calls.items()))
    if args.profile:
      pr.disable()
      pstats.Stats(pr).sort_stats(os.environ.get('SORT', 'cumulative')).print_stats(args.top)
    print('%-18s %.1f us per evaluate() on the host; stubbed entry points per call: %s'
          % (name, dt / args.iters * 1e6, calls))
    eval_cache.clear()


if __name__ == '__main__':
  main()
