"""Sparse matrix x dense vector on the device (PageRank's SpMV, BASELINE config 5).

Reference: `expr.dot(wts, p)` with `wts` a sparse (N, M) array tiled in column strips
(tests/benchmark_pagerank.py:27-45) -> map2((a,b),(1,0)) -> per strip `tiles[0].tocsr().dot(tiles[1])`
(spartan/expr/dot.py:213-217; legacy hash-map variant spartan/array/sparse.pyx:103-158) producing a
full-length partial y that is np.add-merged on the owners.

Here a strip is a CSR triple resident in the HBM of its owner (placement = the reference's round-robin over
column strips), a strip's product is one `sp_spmv_csr` launch accumulating into a rank-local y, and the merge
is one ncclReduceScatter(sum) of y into the owners' row blocks (ncclAllReduce for irregular result tilings).  Sparse *tiles* as a general array type stay out of scope: only this
operation is provided."""
import numpy as np
import scipy.sparse as sp_sparse
import torch

from . import blob_ctx, comm, device_ops
from ._lib import lib, check, SP_RED_SUM, SpartanError
from .array import distarray, extent
from .expr.base import Expr, lazify


def _strip_layout(n_cols, strip_width, num_workers):
  """The reference's column strips, their round-robin owners, and the per-owner blocks of adjacent strips."""
  strips = [(c0, min(n_cols, c0 + strip_width), i % num_workers) for i, c0 in enumerate(range(0, n_cols, strip_width))]
  blocks = []
  for c0, c1, owner in strips:
    if blocks and blocks[-1][2] == owner and blocks[-1][1] == c0:
      blocks[-1] = (blocks[-1][0], c1, owner)
    else:
      blocks.append((c0, c1, owner))
  return strips, blocks


def _csr_from_coo_device(rows, cols, vals, n_rows, width):
  """CSR (int32 rowptr when it fits, int32 column indices, fp32 values) of COO triples already on the device; entries are
  ordered by (row, column); duplicates stay separate entries (their contributions add up in the product)."""
  key = rows.to(torch.int64) * int(width) + cols.to(torch.int64)
  key, order = torch.sort(key)
  r = torch.div(key, int(width), rounding_mode='floor')
  c = (key - r * int(width)).to(torch.int32)
  counts = torch.bincount(r, minlength=n_rows)
  rowptr = torch.zeros(n_rows + 1, dtype=torch.int64, device=rows.device)
  torch.cumsum(counts, 0, out=rowptr[1:])
  return rowptr, c.contiguous(), vals[order].to(torch.float32).contiguous()


class SparseStrips(object):
  """Column strips of a sparse matrix; strip i lives on rank i % num_workers.  Built from a scipy matrix every rank
  holds (``SparseStrips(matrix)``) or from COO triples generated on the device (``SparseStrips.from_device_coo``)."""

  def __init__(self, matrix=None, strip_width=None, shape=None):
    ctx = blob_ctx.get()
    self.dtype = np.dtype(np.float32)
    self.sparse = True
    if matrix is None:
      self.shape = tuple(shape)
      self.strips, self.blocks, self.nnz = [], [], 0
      return
    matrix = sp_sparse.csc_matrix(matrix)
    self.shape = matrix.shape
    n_rows, n_cols = self.shape
    if strip_width is None:
      strip_width = -(-n_cols // ctx.num_workers)
    self.nnz = int(matrix.nnz)
    # Adjacent strips of one owner are stored as ONE CSR block (one launch per contiguous block of the
    # rank's share, like the dense slabs): (c0, c1, owner, device CSR or None, nnz)
    self.strips, blocks = _strip_layout(n_cols, strip_width, ctx.num_workers)
    built = []
    for c0, c1, owner in blocks:
      dev, nnz = None, 0
      if owner == ctx.worker_id:
        csr = matrix[:, c0:c1].tocsr()
        csr.sum_duplicates()
        nnz = int(csr.nnz)
        dev = _to_device_csr(csr.indptr, csr.indices, csr.data, ctx.device)
      built.append((c0, c1, owner, dev, nnz))
    self.blocks = built

  @classmethod
  def from_device_coo(cls, shape, strip_width, entries):
    """``entries(c0, c1) -> (rows, cols, vals)`` device tensors of the strip's non-zeros (global column indices);
    called only for the strips this rank owns, so every rank builds just its share (the reference's make_weights mapper
    also runs once per tile on the tile's worker, tests/benchmark_pagerank.py:11-24)."""
    ctx = blob_ctx.get()
    self = cls(None, shape=shape)
    n_rows, n_cols = self.shape
    self.strips, blocks = _strip_layout(n_cols, strip_width, ctx.num_workers)
    built = []
    total = 0
    for c0, c1, owner in blocks:
      dev, nnz = None, 0
      if owner == ctx.worker_id:
        parts = [entries(s0, s1) for s0, s1, _ in self.strips if c0 <= s0 and s1 <= c1]
        rows = torch.cat([p[0] for p in parts]); cols = torch.cat([p[1] for p in parts]) - c0
        vals = torch.cat([p[2] for p in parts])
        rowptr, col, val = _csr_from_coo_device(rows, cols, vals, n_rows, c1 - c0)
        nnz = int(val.numel())
        if nnz < 2 ** 31:
          rowptr = rowptr.to(torch.int32)
        dev = (rowptr, col, val)
      total += nnz
      built.append((c0, c1, owner, dev, nnz))
    self.blocks = built
    if ctx.num_workers > 1:
      t = torch.tensor([total], dtype=torch.int64, device=ctx.device)
      comm.allreduce(t, SP_RED_SUM)
      total = int(t.item())
    self.nnz = total
    return self


def _to_device_csr(indptr, indices, data, device):
  """Row pointers as int32 whenever nnz < 2**31 (half the bytes the kernel streams per row)."""
  ptr_dtype = np.int32 if int(indptr[-1]) < 2 ** 31 else np.int64
  return (torch.from_numpy(np.ascontiguousarray(indptr.astype(ptr_dtype))).to(device),
          torch.from_numpy(np.ascontiguousarray(indices.astype(np.int32))).to(device),
          torch.from_numpy(np.ascontiguousarray(data.astype(np.float32))).to(device))


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


def from_device_coo(shape, strip_width, entries):
  """A sparse matrix whose strips are generated on the device of their owner (see SparseStrips.from_device_coo)."""
  return SparseStripsExpr(val=SparseStrips.from_device_coo(shape, strip_width, entries))


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
      check(lib.sp_spmv_csr(rowptr.data_ptr(), 1 if rowptr.dtype == torch.int64 else 0, col.data_ptr(), val.data_ptr(),
                            n_rows, xs.data_ptr(), y.data_ptr(), 1, max(1, -(-nnz // max(1, n_rows))), ctx.stream_ptr()),
            'sp_spmv_csr')
      ctx.kernel_launches += 1
    out_shape = self.compute_shape()
    out = distarray.create(out_shape, np.float32, reducer=np.add, tile_hint=self.tile_hint)
    W = ctx.num_workers
    # The np.add merge of the strips' partial y (tile.pyx:263-268).  When rank r owns exactly the r-th of W equal row
    # blocks of the result (the default tiling), a reduce-scatter straight into its slab moves half the bytes of an
    # all-reduce; otherwise all-reduce and let every rank copy out the tiles it owns.
    rows_per = n_rows // W if W > 0 else n_rows
    regular = (W > 1 and ctx.device.type == 'cuda' and n_rows % W == 0 and out.slab is not None
               and out.slab.is_contiguous() and out.slab.numel() == rows_per)
    if regular:
      for ex, tid in out.tiles.items():
        if not (tid.worker * rows_per <= ex.ul[0] and ex.lr[0] <= (tid.worker + 1) * rows_per):
          regular = False
          break
    if regular:
      import torch.distributed as dist
      dist.reduce_scatter_tensor(out.slab.reshape(-1), y, op=dist.ReduceOp.SUM)
      for tid in out.tiles.values():
        if ctx.is_local(tid):
          ctx.tile(tid).valid = True
      return out
    comm.allreduce(y, SP_RED_SUM)
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
