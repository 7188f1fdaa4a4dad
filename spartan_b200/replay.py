"""Replayable evaluations: a whole ``evaluate()`` -- fused kernels, finalisation, the NCCL combine -- captured once
into a CUDA graph and replayed with a single launch.

The reference re-uses work through its expression cache (``EvalCache``, spartan/expr/operator/base.py:73-114: an
expression evaluated twice is computed once).  Iterative drivers (k-means, PageRank, the benchmark loops in
tests/benchmark_*.py) instead re-evaluate the *same DAG on changed data* every iteration, and on a B200 the Python
host that walks the DAG (~0.3-0.5 ms) is then slower than the kernels it launches for anything below ~2 GiB of
operands.  ``replayable(fn)`` runs ``fn`` (any function that builds and evaluates expressions over arrays that
already live on the device) a couple of times to warm up -- run-time specialisations get compiled, scratch buffers
sized, NCCL communicators created -- then captures one more run; calling the result replays the captured
launches on the current values of the input arrays and returns the same output array object (updated in place).
"""
import torch

from . import blob_ctx
from ._lib import SpartanError


class Replayable(object):
  def __init__(self, fn, warmup=2):
    ctx = blob_ctx.get()
    if ctx.device.type != 'cuda':
      raise SpartanError('replayable() needs a CUDA device (there is no CPU fallback)')
    self.ctx = ctx
    main = torch.cuda.current_stream(ctx.device)
    side = ctx.side_stream('capture')
    side.wait_stream(main)
    with torch.cuda.stream(side):
      for _ in range(max(1, warmup)):
        fn()
    main.wait_stream(side)
    torch.cuda.synchronize(ctx.device)
    launches0 = ctx.kernel_launches
    ctx.graph_captures += 1           # from now on outgrown scratch buffers are retired, not freed (their addresses are in the graph)
    self.graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(self.graph, stream=side):
      self.result = fn()
    self.kernel_launches = ctx.kernel_launches - launches0     # library kernels per replay
    torch.cuda.synchronize(ctx.device)

  def __call__(self):
    self.graph.replay()
    self.ctx.touch()                  # the captured launches rewrote their output arrays in place
    self.ctx.kernel_launches += self.kernel_launches
    return self.result


def replayable(fn, warmup=2):
  """Capture ``fn`` (builds + evaluates expressions; must not read results back to the host) for replay."""
  return Replayable(fn, warmup)
