"""dot from pinned host arrays on one GPU (upload || contraction || read-back) for several strip widths
(FLAGS.dot_stream_strip): ms per evaluation incl. the read-back of C."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import spartan_b200 as sp
from spartan_b200.expr.base import eval_cache
ctx = sp.initialize()
n, tile = 32768, 4096
a = torch.empty((n, n), dtype=torch.float32, pin_memory=True); b = torch.empty((n, n), dtype=torch.float32, pin_memory=True)
out = torch.empty((n, n), dtype=torch.float32, pin_memory=True)
A = sp.rand(n, n, seed=0, dtype=np.float32, tile_hint=(tile, tile)).evaluate(); A.read_local_into(a.numpy())
B = sp.rand(n, n, seed=1, dtype=np.float32, tile_hint=(tile, tile)).evaluate(); B.read_local_into(b.numpy())
torch.cuda.synchronize(); del A, B
def step():
  eval_cache.clear()
  c = sp.dot(sp.from_numpy(a.numpy(), tile_hint=(tile, tile)), sp.from_numpy(b.numpy(), tile_hint=(tile, tile)), tile_hint=(tile, tile)).evaluate()
  c.read_local_into(out.numpy()); torch.cuda.current_stream().synchronize()
for strip in [int(s) for s in (sys.argv[1:] or ['4096', '2048', '3072', '1024', '4096'])]:
  sp.FLAGS.dot_stream_strip = strip
  step(); step()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(4): step()
  e1.record(); torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / 4
  print(json.dumps({'strip': strip, 'ms': round(ms, 2), 'tflops': round(2 * n ** 3 / ms / 1e9, 1)}), flush=True)
