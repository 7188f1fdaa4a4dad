"""cProfile of the host side of one fused map+reduce step (GPU work is asynchronous, so wall time here is host time)."""
import cProfile, pstats, io, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import spartan_b200 as sp
from spartan_b200.expr.base import lazify
ctx = sp.initialize()
rows, cols = 4096, 32768
X = sp.rand(rows, cols, seed=2, dtype=np.float32, tile_hint=(512, cols)).evaluate()
Y = sp.rand(rows, cols, seed=3, dtype=np.float32, tile_hint=(512, cols)).evaluate()
def step():
  return (lazify(X) * 2 + lazify(Y)).sum(axis=0).optimized().evaluate()
for _ in range(5): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50): step()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print('host ms/step %.3f   incl. drain %.3f' % ((t1 - t0) * 20, (t2 - t0) * 20))
pr = cProfile.Profile(); pr.enable()
for _ in range(50): step()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(35); print(s.getvalue()[:6000])
