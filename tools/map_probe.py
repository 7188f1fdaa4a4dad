"""Pure map throughput (x*2+y over 1 GiB operands, contiguous and column-block tilings) next to torch's own
elementwise kernel on the same buffers.  Prints algorithmic GB/s (2 reads + 1 write)."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import spartan_b200 as sp
from spartan_b200.expr.base import lazify

ctx = sp.initialize()
rows, cols = 8192, 32768
nb = rows * cols * 4


def timeit(fn, n=10):
  for _ in range(3): fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(n): fn()
  e1.record(); torch.cuda.synchronize()
  return e0.elapsed_time(e1) / n

print(json.dumps({"map_kernel": os.environ.get("SPARTAN_MAP_KERNEL", "direct")}), flush=True)
for hint in [(rows, cols), (rows, cols // 8)]:
  X = sp.rand(rows, cols, seed=2, dtype=np.float32, tile_hint=hint).evaluate()
  Y = sp.rand(rows, cols, seed=3, dtype=np.float32, tile_hint=hint).evaluate()
  x, y = lazify(X), lazify(Y)
  ms = timeit(lambda: (x * 2 + y).optimized().evaluate())
  print(json.dumps({'case': 'sp x*2+y map', 'tile_hint': hint, 'ms': round(ms, 3), 'GBps': round(3 * nb / ms / 1e6, 1)}), flush=True)
  ms = timeit(lambda: (x * y).optimized().evaluate())
  print(json.dumps({'case': 'sp x*y map', 'tile_hint': hint, 'ms': round(ms, 3), 'GBps': round(3 * nb / ms / 1e6, 1)}), flush=True)
  ms = timeit(lambda: (x * 2).optimized().evaluate())
  print(json.dumps({'case': 'sp x*2 map', 'tile_hint': hint, 'ms': round(ms, 3), 'GBps': round(2 * nb / ms / 1e6, 1)}), flush=True)
  # the same steps replayed from a CUDA graph: what the kernels do without the Python host between launches
  for name, fn, nbytes in [('x*2+y map', lambda: (x * 2 + y).optimized().evaluate(), 3 * nb),
                           ('x*2 map', lambda: (x * 2).optimized().evaluate(), 2 * nb),
                           ('(x*2+y).sum(0)', lambda: (x * 2 + y).sum(axis=0).optimized().evaluate(), 2 * nb),
                           ('(x*2+y).sum()', lambda: (x * 2 + y).sum().optimized().evaluate(), 2 * nb)]:
    rep = sp.replayable(fn)
    ms = timeit(rep, n=20)
    print(json.dumps({'case': 'sp replay ' + name, 'tile_hint': hint, 'ms': round(ms, 3), 'GBps': round(nbytes / ms / 1e6, 1)}), flush=True)
    del rep
a = torch.rand(rows, cols, device='cuda'); b = torch.rand(rows, cols, device='cuda'); c = torch.empty_like(a)
ms = timeit(lambda: torch.add(a, b, alpha=1.0, out=c))
print(json.dumps({'case': 'torch.add(out=)', 'ms': round(ms, 3), 'GBps': round(3 * nb / ms / 1e6, 1)}), flush=True)
ms = timeit(lambda: torch.mul(a, 2.0, out=c))
print(json.dumps({'case': 'torch.mul(scalar, out=)', 'ms': round(ms, 3), 'GBps': round(2 * nb / ms / 1e6, 1)}), flush=True)
ms = timeit(lambda: c.copy_(a))
print(json.dumps({'case': 'torch copy_', 'ms': round(ms, 3), 'GBps': round(2 * nb / ms / 1e6, 1)}), flush=True)
