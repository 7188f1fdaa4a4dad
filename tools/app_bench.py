"""Measures BASELINE configs 4 (k-means 10M x 256, k=1024, one iteration) and 5 (PageRank SpMV, N=10M, 10
outlinks/page) on the GPUs of this job and the oracle's CPU path on a bounded sample.  One JSON line each.
Run:  python tools/app_bench.py [--small]     (torchrun for several GPUs)"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
  sys.path.insert(0, p)
import torch
import spartan_b200 as sp
from spartan_b200 import comm
from spartan_oracle import apps

ap = argparse.ArgumentParser(); ap.add_argument('--small', action='store_true'); ap.add_argument('--steps', type=int, default=5)
args = ap.parse_args()
ctx = sp.initialize()
W, rank = ctx.num_workers, ctx.worker_id
peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {'hbm_gbs': 6650.0, 'bf16_tflops_sustained': 1400.0}


def timed(fn, steps, warmup=2):
  for _ in range(warmup): fn()
  comm.barrier(); torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(steps): fn()
  e1.record(); comm.barrier(); torch.cuda.synchronize()
  return e0.elapsed_time(e1) / steps


# ---------------- config 4: k-means
n, d, k = (1000000, 256, 1024) if args.small else (10000000, 256, 1024)
rows = n // 8
X = sp.rand(n, d, seed=4, dtype=np.float32, tile_hint=(rows, d)).evaluate()
c0 = X.fetch(sp.extent.create((0, 0), (k, d), (n, d)), dst=0)
c0 = c0.cpu().numpy() if rank == 0 else np.zeros((k, d), np.float32)
if W > 1:
  t = torch.from_numpy(c0).to(ctx.device); comm.broadcast(t, 0); c0 = t.cpu().numpy()
km = sp.KMeans(n_clusters=k, n_iter=1)
ms = timed(lambda: km.fit(X, centers=c0), args.steps)
flops = 2.0 * n * d * k
out = {'config': 'k-means %dx%d k=%d, 1 iteration (assign + accumulate + allreduce)' % (n, d, k), 'n_gpus': W, 'ms_per_iter': ms,
       'distance_tflops': flops / ms / 1e9, 'points_gbs': n * d * 4 / ms / 1e6,
       'frac_of_bf16_sustained_per_gpu': flops / ms / 1e9 / W / peaks['bf16_tflops_sustained']}
if rank == 0:
  m = min(n, 20000)
  xs = X.fetch(sp.extent.create((0, 0), (m, d), (n, d))).cpu().numpy() if W == 1 else None
  if xs is not None:
    centers, labels = km.fit(X, centers=c0)
    lab = labels.fetch(sp.extent.create((0,), (m,), (n,))).cpu().numpy()
    t0 = time.perf_counter(); ref = apps.kmeans_dist_mapper(xs, c0); t1 = time.perf_counter()
    out['label_agreement_sample'] = float((lab == ref).mean())
    out['cpu_baseline'] = {'kind': 'port', 'sample': 'scipy cdist+argmin on %d points' % m, 'cores': 1,
                           'ms_per_iter_extrapolated': (t1 - t0) * 1e3 * n / m}
  print(json.dumps(out), flush=True)
del X
torch.cuda.empty_cache()

# ---------------- config 5: PageRank SpMV
N = 1000000 if args.small else 10000000
outlinks = 10
t0 = time.perf_counter()
Wm = apps.make_weights(N, outlinks, seed=5)
strip = N // 8
wts = sp.sparse.from_scipy(Wm, strip_width=strip)
p = sp.ones((N, 1), tile_hint=(strip, 1)).evaluate()
build_s = time.perf_counter() - t0
from spartan_b200.expr.base import lazify
hold = {}
def step(): hold['y'] = sp.dot(wts, lazify(p)).evaluate()
ms = timed(step, args.steps)
nnz = wts.val.nnz
bytes_alg = 8.0 * nnz + 12.0 * N
out = {'config': 'PageRank SpMV N=%d, %d outlinks/page (nnz=%d), column strips N/8' % (N, outlinks, nnz), 'n_gpus': W,
       'ms_per_spmv': ms, 'algorithmic_gbs': bytes_alg / ms / 1e6, 'frac_of_hbm_per_gpu': bytes_alg / ms / 1e6 / W / peaks['hbm_gbs'],
       'host_build_s': build_s}
if rank == 0:
  if W == 1:
    y = hold['y'].glom().reshape(-1)
    t0 = time.perf_counter(); ref = apps.spmv_strips(Wm, np.ones(N, np.float32), strip); t1 = time.perf_counter()
    out['max_rel_err'] = float(np.abs(y - ref).max() / np.abs(ref).max())
    out['cpu_baseline'] = {'kind': 'port', 'sample': 'scipy csr matvec per strip, whole matrix', 'cores': 1, 'ms': (t1 - t0) * 1e3}
  print(json.dumps(out), flush=True)
