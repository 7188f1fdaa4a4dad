"""Zero-copy views over distributed arrays: Slice, Transpose, Reshape.

Reference: spartan/expr/operator/slice.py:42-85 (Slice), transpose.py:27-67 (Transpose),
reshape.py:47-193 (Reshape).  Like the reference, a view never copies the base array: ``fetch``
translates the requested extent into base coordinates and asks the base.  On the device the answer
is a *strided* tensor view into the base's HBM (a sub-rectangle, permuted strides, or a re-shaped
contiguous run), which the fused map / map+reduce kernels consume directly through their operand
strides -- so ``x[1:] - x[:-1]``, ``x[:, :, 0].sum()`` or ``transpose(x) * 2`` run as one launch per
output tile with no materialised intermediate.

A view reports ``tiles`` in its OWN coordinates (view extent -> TileId of the base tile that
backs it), which is what the evaluator's per-tile loop needs: the output of a map over a view is
tiled and placed like the view.
"""
import collections

import numpy as np

from . import extent
from .. import blob_ctx, comm, device_ops
from ..core import TileId
from ..util import require_type, require_equal, require_unique
from . import distarray
from .distarray import DistArray, _tile_mapper


class ViewArray(DistArray):
  """Machinery shared by the views: tile table in view coordinates, per-tile dispatch, glom."""
  is_view = True
  slab = None          # a view owns no memory: the blockwise (slab) fast paths do not apply
  sparse = False
  reducer_fn = None

  def __init__(self):
    self.bad_tiles = []
    self.ctx = blob_ctx.get()
    self._tiles = None

  def _view_tiles(self):
    raise NotImplementedError

  @property
  def tiles(self):
    if self._tiles is None:
      self._tiles = self._view_tiles()
      self.blob_to_ex = dict((tid, ex) for ex, tid in self._tiles.items())
    return self._tiles

  def extent_for_blob(self, id):
    self.tiles
    return self.blob_to_ex[id]

  def foreach_tile(self, mapper_fn, kw=None):
    """slice.py:71-75 / transpose.py:51-55 / reshape.py:139-155: run ``mapper_fn`` once per view tile."""
    kw = dict(kw or {})
    kw['array'] = self
    kw['user_fn'] = mapper_fn
    return self.ctx.map(list(self.tiles.values()), mapper_fn=_tile_mapper, kw=kw)

  def local_blocks(self):
    me = self.ctx.worker_id
    return [ex for ex, tid in self.tiles.items() if tid.worker == me]

  def same_layout(self, other):
    return False

  def tile_shape(self):
    return tuple(distarray.good_tile_shape(self.shape, self.ctx.num_workers))

  def glom(self):
    """Every rank ends up with the whole view on the host (one piece per view tile)."""
    ctx = self.ctx
    out = np.empty(self.shape, dtype=self.dtype)
    if len(self.shape) == 0:
      t = self.fetch(extent.create((), (), ()), dst=0 if ctx.num_workers > 1 else None)
      if ctx.num_workers > 1:
        t = t if t is not None else ctx.empty((), self.dtype)
        comm.broadcast(t, 0)
      out[()] = t.cpu().numpy()
      return out
    for ex, tid in sorted(self.tiles.items(), key=lambda kv: (kv[1].worker, kv[1].id)):
      if ctx.num_workers == 1:
        data = self.fetch(ex)
      else:
        data = self.fetch(ex, dst=tid.worker)
        data = data.contiguous() if tid.worker == ctx.worker_id else ctx.empty(ex.shape, self.dtype)
        comm.broadcast(data, tid.worker)
      out[ex.to_slice()] = data.cpu().numpy().reshape(ex.shape)
    return out


class Slice(ViewArray):
  """A NumPy basic slice of a DistArray (slice.py:42-85); unit steps only, like extent.from_slice."""

  def __init__(self, darray, idx):
    ViewArray.__init__(self)
    if not isinstance(idx, extent.TileExtent):
      idx = extent.from_slice(idx, darray.shape)
    require_type(darray, DistArray)
    if idx is None:
      raise ValueError('empty slice')
    self.base = darray
    self.slice = idx
    self.shape = tuple(self.slice.shape)
    self.dtype = darray.dtype

  def _view_tiles(self):
    # _slice_mapper (slice.py:9-39): the view tile of a base tile is its intersection with the slice, re-based
    tiles = collections.OrderedDict()
    for base_ex, tid in self.base.tiles.items():
      inter = extent.intersection(self.slice, base_ex)
      if inter is None:
        continue
      off = extent.offset_from(self.slice, inter)
      tiles[extent.create(off.ul, off.lr, self.shape)] = tid
    return tiles

  def fetch(self, idx, dst=None):
    """slice.py:81-83."""
    offset = extent.compute_slice(self.slice, idx.to_slice())
    return self.base.fetch(offset, dst=dst)


class Transpose(ViewArray):
  """All axes reversed (transpose.py:27-67).  ``fetch`` returns the base region with permuted strides."""

  def __init__(self, base):
    ViewArray.__init__(self)
    require_type(base, DistArray)
    self.base = base
    self.shape = tuple(base.shape[::-1])
    self.dtype = base.dtype

  def tile_shape(self):
    return tuple(self.base.tile_shape()[::-1])

  def _view_tiles(self):
    tiles = collections.OrderedDict()
    for base_ex, tid in self.base.tiles.items():
      tiles[extent.create(base_ex.ul[::-1], base_ex.lr[::-1], self.shape)] = tid
    return tiles

  def base_extent(self, ex):
    return extent.create(ex.ul[::-1], ex.lr[::-1], self.base.shape)

  def fetch(self, ex, dst=None):
    """transpose.py:62-66."""
    t = self.base.fetch(self.base_extent(ex), dst=dst)
    if t is None or t.dim() < 2:
      return t
    return t.permute(*reversed(range(t.dim())))


def _ravelled_ex(ul, lr, shape):
  """reshape.py:20-23."""
  return extent.ravelled_pos(ul, shape), extent.ravelled_pos([l - 1 for l in lr], shape)


class Reshape(ViewArray):
  """Row-major re-shape of a DistArray (reshape.py:47-193).

  ``fetch(ex)`` reads the smallest run of complete leading-dimension rows of the base that covers the
  C-order element range of ``ex``, views it flat, and cuts ``ex`` out of it; unlike the reference
  (reshape.py:170-172 "can't handle column fetch") this also serves extents that do not consist of
  complete rows, so tiles of any shape can be mapped."""

  def __init__(self, base, shape, tile_hint=None):
    ViewArray.__init__(self)
    require_type(base, DistArray)
    shape = tuple(int(s) for s in shape)
    if int(np.prod(shape, dtype=np.int64)) != int(np.prod(base.shape, dtype=np.int64)):
      raise ValueError('total size of new array must be unchanged: %s -> %s' % (base.shape, shape))
    self.base = base
    self.shape = shape
    self.dtype = base.dtype
    self._tile_hint = tile_hint
    self._tile_shape = tuple(distarray.good_tile_shape(shape, self.ctx.num_workers)) if len(shape) else ()
    # reshape.py:92-119: the base tiles stay rectangles only when dimensions are appended
    self._same_tiles = (len(shape) > len(base.shape) and
                        all(base.shape[i] == shape[i] for i in range(len(base.shape))))

  def tile_shape(self):
    return self._tile_shape

  def view_extent(self, ex):
    """reshape.py:124-129."""
    r_ul, r_lr = _ravelled_ex(ex.ul, ex.lr, ex.array_shape)
    ul = extent.unravelled_pos(r_ul, self.shape)
    lr = extent.unravelled_pos(r_lr, self.shape)
    return extent.create(ul, [l + 1 for l in lr], self.shape)

  def _view_tiles(self):
    tiles = collections.OrderedDict()
    W = self.ctx.num_workers
    if self._same_tiles:
      for base_ex, tid in self.base.tiles.items():
        tiles[self.view_extent(base_ex)] = tid
      return tiles
    # reshape.py:146-151 builds an (uninitialised) "shape_array" only for its tile table; the table alone is kept
    if len(self.shape) == 0:
      tiles[extent.create((), (), ())] = TileId(0, -1)
      return tiles
    for i, (ex, worker) in enumerate(distarray.compute_extents(self.shape, self._tile_hint or self._tile_shape,
                                                               W).items()):
      tiles[ex] = TileId(worker % W, -1 - i)
    return tiles

  def fetch(self, ex, dst=None):
    """reshape.py:165-193 (dense)."""
    nd_new, nd_base = len(self.shape), len(self.base.shape)
    if nd_new == 0 or nd_base == 0 or int(np.prod(ex.shape, dtype=np.int64)) == 0:
      t = self.base.fetch(extent.from_shape(self.base.shape) if nd_base else extent.create((), (), ()), dst=dst)
      return None if t is None else device_ops.materialize(t).reshape(ex.shape if nd_new else ())
    # the rows of the NEW shape that contain ex form one C-order run [g0, g1); read the complete
    # leading-dimension rows of the base that cover that run
    nrow = int(np.prod(self.shape[1:], dtype=np.int64))
    row = int(np.prod(self.base.shape[1:], dtype=np.int64))
    g0, g1 = ex.ul[0] * nrow, ex.lr[0] * nrow
    b0, b1 = g0 // row, -(-g1 // row)
    base_ex = extent.create([b0] + [0] * (nd_base - 1), [b1] + list(self.base.shape[1:]), self.base.shape)
    t = self.base.fetch(base_ex, dst=dst)
    if t is None:
      return None
    flat = device_ops.materialize(t).reshape(-1)     # a view when the fetched rows are contiguous, else one gather copy
    f0, f1 = g0 - b0 * row, g1 - b0 * row
    block = flat[f0:f1].reshape((ex.lr[0] - ex.ul[0],) + tuple(self.shape[1:]))
    return block[(slice(None),) + tuple(slice(ex.ul[d], ex.lr[d]) for d in range(1, nd_new))]
