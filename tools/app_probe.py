"""One k-means iteration (10M x 256, k = 1024) and one PageRank SpMV (N = 10M) for profiling under ncu:
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/app_probe.py
  ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel|kmeans|spmv' -o gpurun_out/apps python tools/app_probe.py
Prints nothing that counts as a measurement (numbers under a profiler never do)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import spartan_b200 as sp
from spartan_b200.examples import pagerank
from spartan_b200.expr.base import lazify

what = sys.argv[1] if len(sys.argv) > 1 else 'all'
ctx = sp.initialize()
if what in ('all', 'kmeans'):
  n, d, k = 10000000, 256, 1024
  X = sp.rand(n, d, seed=4, dtype=np.float32, tile_hint=(n // 8, d)).evaluate()
  c0 = X.fetch(sp.extent.create((0, 0), (k, d), (n, d))).cpu().numpy()
  km = sp.KMeans(n_clusters=k, n_iter=2)
  km.fit(X, centers=c0)
  torch.cuda.synchronize()
  del X
  torch.cuda.empty_cache()
if what in ('all', 'spmv'):
  N = 10000000
  wts = pagerank.make_weights(N, N // 8, seed=5)
  p = sp.rand(N, 1, seed=6, dtype=np.float32, tile_hint=(N // 8, 1)).evaluate()
  for _ in range(2):
    y = sp.dot(wts, lazify(p)).evaluate()
  torch.cuda.synchronize()
print('probe done')
