"""diagonal / diagflat / diag (reference: spartan/expr/creation.py:228-330).

The reference runs both through ``map2`` with NumPy tile functions (``_diagonal_mapper``: the diagonal run of every
tile is written to a 1-D target; ``_diagflat_mapper``: full-width rows with one run on the diagonal).  On the device a
diagonal run is a *strided view* of a tile (element stride = row pitch + 1), so both directions are strided rectangle
copies (`sp_copy_rect` / the fused map kernel), no gather kernel:

  * diagonal(a): every rank copies the diagonal runs of its own tiles into a rank-local vector of min(a.shape)
    elements that is zero elsewhere; one ncclAllReduce(sum) merges the disjoint runs (exact: x + 0), and the vector is
    cut into the target's tiles -- the same shape of computation as a reduction.
  * diagflat(v): the target's tiles are zero-filled and the tiles that cross the diagonal get their run from ``v``
    (fetched to the tile's owner).
"""
import numpy as np
import torch

from .. import comm, device_ops
from ..array import distarray, extent
from .._lib import SP_RED_SUM, SP_FILL_CONST, SpartanError
from .base import Expr, lazify


class DiagonalExpr(Expr):
  """``a.diagonal()`` of a 2-D array (creation.py:264-298)."""
  members = ('array',)

  def compute_shape(self):
    return (min(self.array.shape),)

  def _evaluate(self, ctx, deps):
    a = deps['array']
    if len(a.shape) != 2:
      raise ValueError('diag requires an array of two dimensions on the device path')
    n = min(a.shape)
    acc = ctx.empty((n,), a.dtype)
    acc.zero_()
    for ex, tid in sorted(a.tiles.items(), key=lambda kv: (kv[1].worker, kv[1].id)):
      # _diagonal_mapper (creation.py:264-278): the run of the diagonal inside this tile starts at max(ul)
      d0 = max(ex.ul)
      d1 = min(ex.lr[0], ex.lr[1], n)
      if d0 >= d1:
        continue
      region = extent.create((d0, d0), (d1, d1), a.shape)
      block = a.fetch(region, dst=tid.worker)           # zero-copy view on the owner
      if tid.worker == ctx.worker_id:
        device_ops.copy_into(acc[d0:d1], block.diagonal())
    if ctx.num_workers > 1:
      if acc.dtype.is_floating_point or acc.dtype in (torch.int32, torch.int64):
        comm.allreduce(acc, SP_RED_SUM)
      else:
        raise SpartanError('diagonal() of %s arrays across ranks is not supported' % a.dtype)
    out = distarray.create((n,), a.dtype)
    for ex, tid in out.tiles.items():
      if ctx.is_local(tid):
        t = ctx.tile(tid)
        device_ops.copy_rect(t.get(None), acc[ex.to_slice()])
        t.valid = True
    return out


class DiagFlatExpr(Expr):
  """``np.diagflat(v)``: an (n, n) array with the (flattened) data on the diagonal (creation.py:228-261)."""
  members = ('array', 'tile_hint')

  def compute_shape(self):
    n = int(np.prod(self.array.shape))
    return (n, n)

  def _evaluate(self, ctx, deps):
    v = deps['array']
    n = int(np.prod(v.shape))
    if len(v.shape) != 1:
      from ..array.views import Reshape
      v = Reshape(v, (n,))
    out = distarray.create((n, n), v.dtype, tile_hint=self.tile_hint)
    for ex, tid in sorted(out.tiles.items(), key=lambda kv: (kv[1].worker, kv[1].id)):
      d0 = max(ex.ul)
      d1 = min(ex.lr[0], ex.lr[1])
      run = v.fetch(extent.create((d0,), (d1,), (n,)), dst=tid.worker) if d0 < d1 else None
      if not ctx.is_local(tid):
        continue
      t = ctx.tile(tid)
      data = t.get(None)
      device_ops.fill_view(data, SP_FILL_CONST, 0.0)
      if run is not None:
        device_ops.copy_into(data[d0 - ex.ul[0]:d1 - ex.ul[0], d0 - ex.ul[1]:d1 - ex.ul[1]].diagonal(), run)
      t.valid = True
    return out


def diagonal(a):
  """Return the main diagonal of a 2-D array (creation.py:281-298)."""
  a = lazify(a)
  if len(a.shape) < 2:
    raise ValueError('diag requires an array of at least two dimensions')
  return DiagonalExpr(array=a)


def diagflat(array, tile_hint=None):
  """creation.py:252-261."""
  return DiagFlatExpr(array=lazify(array), tile_hint=tile_hint)


def diag(array, offset=0):
  """Extract a diagonal or construct a diagonal array (creation.py:301-330)."""
  if offset != 0:
    raise NotImplementedError
  array = lazify(array)
  if len(array.shape) == 1:
    return diagflat(array)
  if len(array.shape) == 2:
    return diagonal(array)
  raise ValueError('Input must be 1- or 2-d.')
