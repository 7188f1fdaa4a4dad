"""Sparse matrix x dense vector on the device (PageRank's SpMV, BASELINE config 5).

Reference: `expr.dot(wts, p)` with `wts` a sparse (N, M) array tiled in column strips
(tests/benchmark_pagerank.py:27-45) -> map2((a,b),(1,0)) -> per strip `tiles[0].tocsr().dot(tiles[1])`
(spartan/expr/dot.py:213-217; legacy hash-map variant spartan/array/sparse.pyx:103-158) producing a
full-length partial y that is np.add-merged on the owners.

Here a strip is a CSR triple resident in the HBM of its owner (placement = the reference's round-robin over
column strips), a strip's product is one `sp_spmv_csr` launch accumulating into a rank-local y, and the merge
is one ncclAllReduce(sum) of y.  Sparse *tiles* as a general array type stay out of scope: only this
operation is provided."""
import numpy as np
import scipy.sparse as sp_sparse
import torch

from . import blob_ctx, comm, device_ops
from ._lib import lib, check, SP_RED_SUM, SpartanError
from .array import distarray, extent
from .expr.base import Expr, lazify


class SparseStrips(object):
  """Column strips of a scipy sparse matrix; strip i lives on rank i % num_workers."""

  def __init__(self, matrix, strip_width=None):
    ctx = blob_ctx.get()
    matrix = sp_sparse.csc_matrix(matrix)
    self.shape = matrix.shape
    self.dtype = np.dtype(np.float32)
    self.sparse = True
    n_rows, n_cols = self.shape
    if strip_width is None:
      strip_width = -(-n_cols // ctx.num_workers)
    self.strips = []          # (c0, c1, owner): the reference's column strips and their round-robin owners
    self.nnz = int(matrix.nnz)
    for i, c0 in enumerate(range(0, n_cols, strip_width)):
      self.strips.append((c0, min(n_cols, c0 + strip_width), i % ctx.num_workers))
    # Adjacent strips of one owner are stored as ONE CSR block (one launch per contiguous block of the
    # rank's share, like the dense slabs): (c0, c1, owner, device CSR or None, nnz)
    self.blocks = []
    for c0, c1, owner in self.strips:
      if self.blocks and self.blocks[-1][2] == owner and self.blocks[-1][1] == c0:
        self.blocks[-1] = (self.blocks[-1][0], c1, owner)
      else:
        self.blocks.append((c0, c1, owner))
    built = []
    for c0, c1, owner in self.blocks:
      dev, nnz = None, 0
      if owner == ctx.worker_id:
        csr = matrix[:, c0:c1].tocsr()
        csr.sum_duplicates()
        nnz = int(csr.nnz)
        dev = (torch.from_numpy(csr.indptr.astype(np.int64)).to(ctx.device),
               torch.from_numpy(csr.indices.astype(np.int32)).to(ctx.device),
               torch.from_numpy(csr.data.astype(np.float32)).to(ctx.device))
      built.append((c0, c1, owner, dev, nnz))
    self.blocks = built


class SparseStripsExpr(Expr):
  members = ('val',)
  needs_cache = False

  def visit(self, visitor):
    return self

  def dependencies(self):
    return {}

  def compute_shape(self):
    return self.val.shape

  def _evaluate(self, ctx, deps):
    return self.val


def from_scipy(matrix, strip_width=None):
  """A sparse (N, M) matrix as column strips on the GPUs (the layout benchmark_pagerank.py:30-38 builds)."""
  return SparseStripsExpr(val=SparseStrips(matrix, strip_width))


class SpMVExpr(Expr):
  members = ('matrix', 'vector', 'tile_hint')

  def compute_shape(self):
    v = self.vector.shape
    return (self.matrix.shape[0],) if len(v) == 1 else (self.matrix.shape[0], v[1])

  def _evaluate(self, ctx, deps):
    A, xv = deps['matrix'], deps['vector']
    if isinstance(xv, np.ndarray):
      xv = distarray.LocalWrapper(xv)
    n_rows, n_cols = A.shape
    vshape = tuple(xv.shape)
    if vshape not in ((n_cols,), (n_cols, 1)) or np.dtype(xv.dtype) != np.float32:
      raise SpartanError('SpMV needs a float32 vector of length %d (got %s %s)' % (n_cols, vshape, xv.dtype))
    y = torch.zeros((n_rows,), dtype=torch.float32, device=ctx.device)
    for c0, c1, owner, dev, nnz in A.blocks:     # every rank walks every block: fetches may be collective
      region = extent.create((c0,), (c1,), vshape) if len(vshape) == 1 else extent.create((c0, 0), (c1, 1), vshape)
      xs = xv.fetch(region, dst=owner)
      if owner != ctx.worker_id:
        continue
      xs = xs.reshape(-1)
      if not xs.is_contiguous():
        xs = xs.contiguous()
      rowptr, col, val = dev
      check(lib.sp_spmv_csr(rowptr.data_ptr(), col.data_ptr(), val.data_ptr(), n_rows, xs.data_ptr(), y.data_ptr(), 1,
                            max(1, -(-nnz // max(1, n_rows))), ctx.stream_ptr()), 'sp_spmv_csr')
      ctx.kernel_launches += 1
    comm.allreduce(y, SP_RED_SUM)                # the np.add merge of the strips' partial y (tile.pyx:263-268)
    out_shape = self.compute_shape()
    out = distarray.create(out_shape, np.float32, reducer=np.add, tile_hint=self.tile_hint)
    yv = y if len(out_shape) == 1 else y.reshape(n_rows, 1)
    for ex, tid in out.tiles.items():
      if ctx.is_local(tid):
        t = ctx.tile(tid)
        device_ops.copy_rect(t.get(None), yv[ex.to_slice()])
        t.valid = True
    return out


def spmv(matrix, vector, tile_hint=None):
  if not isinstance(vector, np.ndarray):
    vector = lazify(vector)
  return SpMVExpr(matrix=matrix, vector=vector, tile_hint=tile_hint)
