"""K-Means on the device (reference: spartan/examples/sklearn/cluster/k_means_.py:61-160, the 'map2'
implementation: kmeans_map2_dist_mapper -> kmeans_count_mapper -> kmeans_center_mapper per row tile).

One iteration = one `sp_kmeans_assign` per contiguous block of this rank's rows (distance GEMM on the tensor
cores + label/accumulate kernel) followed by ncclAllReduce of the (k x d) sums and the k counts -- the
cross-tile combine the reference's map2 targets were meant to do (they have no reducer and overwrite,
SURVEY.md section 9 Q7).  The centre update and the re-seeding of empty clusters (k_means_.py:148-159) run on the
device too (fused map kernels, counter-based generator), so an iteration never waits for the host."""
import numpy as np
import torch

from .. import blob_ctx, comm, device_ops
from .._lib import lib, check, SP_RED_SUM, SP_F64, SP_FILL_RANDN, SpartanError
from ..array import distarray, extent
from ..expr.base import Expr, evaluate


class KMeans(object):
  def __init__(self, n_clusters=8, n_iter=100):
    self.n_clusters = n_clusters
    self.n_iter = n_iter

  def fit(self, X, centers=None, implementation='map2', seed=0):
    """X: (n_samples, n_features) float32 array tiled by rows.  centers: initial (k, d) ndarray or None.
    Returns (centers ndarray float32, labels DistArray int32)."""
    ctx = blob_ctx.get()
    X = evaluate(X) if isinstance(X, Expr) else X
    if not isinstance(X, distarray.DistArrayImpl) or len(X.shape) != 2 or X.dtype != np.float32:
      raise SpartanError('KMeans.fit needs a 2-D float32 distributed array')
    n, d = X.shape
    k = self.n_clusters
    for ex in X.tiles:
      if ex.ul[1] != 0 or ex.lr[1] != d:
        raise SpartanError('KMeans.fit: X must be tiled by rows (k_means_.py:122)')
    if centers is None:
      centers = np.random.RandomState(seed).rand(k, d)                                   # k_means_.py:132
    centers = np.ascontiguousarray(centers, dtype=np.float32)
    tile_rows = X.tile_shape()[0]
    labels = distarray.create((n,), np.int32, tile_hint=(tile_rows,))
    sums = torch.zeros((k, d), dtype=torch.float32, device=ctx.device)
    counts = torch.zeros((k,), dtype=torch.int64, device=ctx.device)
    # this rank's row blocks, their label views and the points as the tensor cores consume them -- prepared once per
    # array (the points do not change between iterations, nor between fits of an unchanged X)
    blocks = []
    for block in X.local_blocks():
      x = X.fetch(block)
      lab = labels.slab_view(extent.create((block.ul[0],), (block.lr[0],), (n,)))
      if lab is None or not lab.is_contiguous():
        raise SpartanError('KMeans.fit: the label tiles do not line up with the row blocks of X (tile rows %d)' % tile_rows)
      m = x.shape[0]

      def fill(op, x=x, m=m):
        check(lib.sp_kmeans_prepare_points(x.data_ptr(), x.stride(0), m, d, op.buf.data_ptr(), op.buf.numel(),
                                           ctx.stream_ptr()), 'sp_kmeans_prepare_points')
        ctx.kernel_launches += 1
      xp = device_ops.cached_operand(X, ('kmeans_points', block.ul[0], block.lr[0]), m, d, 'bf16x3', fill)
      blocks.append((x, lab, m, xp))
    c_dev = torch.from_numpy(centers).to(ctx.device)
    # centres = sums / counts in float64, stored as float32 (k_means_.py:159); a centre that lost all its points is
    # re-seeded with standard-normal values (:148-157).  All of it on the device -- no host round trip inside the
    # loop: denom = counts + [counts == 0]; c = sums / denom; c += [counts == 0] * randn.  (The reference draws the
    # re-seed values from the master's unseeded np.random; here they come from the counter-based device generator keyed
    # by (seed, iteration), so every rank agrees and a run is reproducible.)
    denom = torch.empty((k, 1), dtype=torch.float64, device=ctx.device)
    rnd = torch.empty((k, d), dtype=torch.float32, device=ctx.device)
    p_denom = device_ops.make_program([('IN', 0), ('IN', 0), ('ISZERO', 0), ('ADD', 0)], SP_F64)
    p_div = device_ops.make_program([('IN', 0), ('IN', 1), ('DIV', 0)], SP_F64)
    p_seed = device_ops.make_program([('IN', 0), ('IN', 1), ('ISZERO', 0), ('IN', 2), ('MUL', 0), ('ADD', 0)], SP_F64)
    counts_col = counts.reshape(k, 1)
    for it in range(self.n_iter):
      sums.zero_(); counts.zero_()
      for x, lab, m, xp in blocks:
        need = lib.sp_kmeans_assign_workspace_bytes(m, d, k)
        ws = ctx.scratch(need, 'kmeans')
        check(lib.sp_kmeans_assign_prepared(xp.buf.data_ptr(), x.data_ptr(), x.stride(0), m, d, c_dev.data_ptr(), k,
                                            lab.data_ptr(), sums.data_ptr(), counts.data_ptr(), ws.data_ptr(), ws.numel(),
                                            ctx.stream_ptr()), 'sp_kmeans_assign_prepared')
        ctx.kernel_launches += 4
      comm.allreduce(sums, SP_RED_SUM)
      comm.allreduce(counts, SP_RED_SUM)
      device_ops.fill(rnd, SP_FILL_RANDN, 0.0, 1.0, seed=(int(seed) * 1000003 + it) & 0xffffffffffff)
      device_ops.run_map(p_denom, [counts_col], denom)
      device_ops.run_map(p_div, [sums, denom], c_dev)
      device_ops.run_map(p_seed, [c_dev, counts_col, rnd], c_dev)
    centers = c_dev.cpu().numpy()
    for tid in labels.tiles.values():
      if ctx.is_local(tid):
        ctx.tile(tid).valid = True
    return centers, labels
