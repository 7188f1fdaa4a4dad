"""Times the pieces of the end-to-end path (H2D upload via from_numpy, D2H read-back) on one GPU."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spartan_b200 as sp
ctx = sp.initialize()
n = 16384
host = torch.empty((n, n), dtype=torch.float32).pin_memory(); host.uniform_(0, 1)
arr = host.numpy()
print(json.dumps({'from_numpy_view_is_pinned': torch.from_numpy(arr).is_pinned(), 'host_is_pinned': host.is_pinned()}))
for hint in [(4096, 4096), None]:
  for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    a = sp.from_numpy(arr, tile_hint=hint).evaluate()
    torch.cuda.synchronize(); t1 = time.perf_counter()
  print(json.dumps({'h2d': str(hint), 'ms': (t1 - t0) * 1e3, 'GBps': n * n * 4 / (t1 - t0) / 1e9}))
out = torch.empty((n, n), dtype=torch.float32).pin_memory()
for _ in range(2):
  torch.cuda.synchronize(); t0 = time.perf_counter()
  out.copy_(a.slab, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
print(json.dumps({'d2h_pinned_ms': (t1 - t0) * 1e3, 'GBps': n * n * 4 / (t1 - t0) / 1e9}))
torch.cuda.synchronize(); t0 = time.perf_counter(); x = a.slab.cpu(); t1 = time.perf_counter()
print(json.dumps({'d2h_pageable_ms': (t1 - t0) * 1e3, 'GBps': n * n * 4 / (t1 - t0) / 1e9}))
