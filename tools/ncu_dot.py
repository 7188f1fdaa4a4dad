"""Workload for an ncu capture of the tensor-core GEMM: spartan.dot 32768^2 fp32 (bf16x3), operands resident."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import spartan_b200 as sp
from spartan_b200.expr.base import lazify
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
ctx = sp.initialize()
A = sp.rand(n, n, seed=0, dtype=np.float32, tile_hint=(4096, 4096)).evaluate()
B = sp.rand(n, n, seed=1, dtype=np.float32, tile_hint=(4096, 4096)).evaluate()
for _ in range(2):
  C = sp.dot(lazify(A), lazify(B), tile_hint=(4096, 4096)).evaluate()
torch.cuda.synchronize()
