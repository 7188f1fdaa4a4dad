"""Workload for an ncu capture of the streaming kernel: a few launches of the fused map x*2+y and of the fused
map+reduce (x*2+y).sum(axis=0) / .sum(axis=1) over 1 GiB operands (see profiles/)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import spartan_b200 as sp
from spartan_b200.expr.base import lazify
ctx = sp.initialize()
rows, cols = 8192, 32768
X = sp.rand(rows, cols, seed=2, dtype=np.float32, tile_hint=(rows, cols)).evaluate()
Y = sp.rand(rows, cols, seed=3, dtype=np.float32, tile_hint=(rows, cols)).evaluate()
x, y = lazify(X), lazify(Y)
for _ in range(3):
  (x * 2 + y).optimized().evaluate()
  (x * 2 + y).sum(axis=0).optimized().evaluate()
  (x * 2 + y).sum(axis=1).optimized().evaluate()
  (sp.abs(x - y) * x + sp.maximum(y, 0.5)).sum(axis=0).optimized().evaluate()     # outside the catalogue: NVRTC
torch.cuda.synchronize()
