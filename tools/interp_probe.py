"""Streaming map / map+reduce throughput for expressions inside and outside the static program catalogue.

Prints one line per expression: algorithmic GB/s (bytes of every distinct operand read + bytes written, over the
CUDA-event time of the whole evaluate()).  Results are checked against NumPy on a small slice first."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import spartan_b200 as sp
from spartan_b200.expr.base import lazify

ctx = sp.initialize()
rows, cols = 8192, 32768          # 1 GiB fp32 per operand
X = sp.rand(rows, cols, seed=2, dtype=np.float32, tile_hint=(rows, cols)).evaluate()
Y = sp.rand(rows, cols, seed=3, dtype=np.float32, tile_hint=(rows, cols)).evaluate()
Z = sp.rand(rows, cols, seed=4, dtype=np.float32, tile_hint=(rows, cols)).evaluate()
x, y, z = lazify(X), lazify(Y), lazify(Z)
nb = rows * cols * 4

CASES = [
  ("x*2+y  (catalogue) sum0", lambda: (x * 2 + y).sum(axis=0), 2 * nb),
  ("x*2+y  (catalogue) map", lambda: (x * 2 + y), 3 * nb),
  ("x*y+z  map", lambda: (x * y + z), 4 * nb),
  ("(x-y)*(x-y) sum0", lambda: ((x - y) * (x - y)).sum(axis=0), 2 * nb),
  ("sqrt(x*x+y*y) map", lambda: sp.sqrt(x * x + y * y), 3 * nb),
  ("x*3+y*2-1 sum0", lambda: (x * 3 + y * 2 - 1).sum(axis=0), 2 * nb),
  ("abs(x-y)*z+x sum", lambda: (sp.abs(x - y) * z + x).sum(), 3 * nb),
  ("10-op chain map", lambda: (((x + y) * z - x) * 0.5 + y * y - z * 2 + 1), 4 * nb),
]

from spartan_b200._lib import lib
jit_on = os.environ.get('SPARTAN_JIT', '1') != '0'
out = []
for name, build, nbytes in CASES:
  e = build().optimized()
  r = e.evaluate()
  for _ in range(3): build().optimized().evaluate()
  torch.cuda.synchronize()
  ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  n = 10
  ev0.record()
  for _ in range(n): build().optimized().evaluate()
  ev1.record(); torch.cuda.synchronize()
  ms = ev0.elapsed_time(ev1) / n
  line = {"expr": name, "jit": jit_on, "ms": round(ms, 3), "GBps": round(nbytes / ms / 1e6, 1)}
  print(json.dumps(line), flush=True)
  out.append(line)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "interp_probe_%s.json" % ("jit" if jit_on else "interp")), "w"), indent=1)
