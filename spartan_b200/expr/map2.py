"""Axis joins of tiles on the device: ``map2`` (spartan/expr/operator/map.py:243-375) and ``outer``
(spartan/expr/operator/outer.py:12-120).

The reference joins tiles by re-partitioning: for every tile of the first array ``join_mapper`` turns the tile's
extent into a strip along ``axes[0]``, fetches the matching strips of the other arrays, calls a user function on the
NumPy tiles and scatters the ``(extent, data)`` pairs it yields into a reducer-backed target
(``target.update``, distarray.py:372-422 -> Tile.merge).  ``dot``, k-means and the sparse product are instances.

The same operator here, with two differences forced by where the tiles live:

  * the function is a **device tile function** (``@device_tile_function``): it receives the join extents and the
    joined strips as device tensors and computes through this library's kernels (``spartan_b200.device_ops``) -- a
    Python function over NumPy tiles cannot run on the GPU and there is no CPU fallback, so anything else raises
    ``NotDeviceMappable``;
  * the job is SPMD: every rank walks every tile of the first array in the same order; the strips travel to the
    rank that owns the tile (collective rectangle fetches), the function runs there, and its results are scattered
    to the owners of the target tiles (collective rectangle updates) where ``Tile.merge`` applies the reducer.
    Because the ranks that do not run the function must still know where its results go, a tile function states its
    target extents separately from its arithmetic (``fn.extents(...)`` -- pure host logic).
"""
import numpy as np

from .. import blob_ctx, device_ops
from ..array import distarray, extent
from .. import util
from .._lib import SpartanError
from .base import Expr, TupleExpr, as_array
from .program import NotDeviceMappable


class DeviceTileFunction(object):
  """A tile function of map2 / outer that runs on the device.

  ``extents(*join_extents, **kw)`` -> list of target extents (called on every rank; must not touch data).
  ``compute(*args, **kw)``         -> list of device tensors, one per target extent (called on the owner only).
  For map2 both take ``(extents, tiles)`` like the reference's ``fn(extents, tiles, **kw)`` (map.py:279);
  for outer ``(ex_a, tile_a, ex_b, tile_b)`` (outer.py:40)."""

  def __init__(self, extents_fn, compute_fn, name=None):
    self.extents_fn = extents_fn
    self.compute_fn = compute_fn
    self.__name__ = name or getattr(compute_fn, '__name__', 'device_tile_function')

  def __repr__(self):
    return 'DeviceTileFunction(%s)' % self.__name__


def device_tile_function(extents_fn):
  """Decorator: ``@device_tile_function(target_extents)`` over the compute function."""
  def wrap(compute_fn):
    return DeviceTileFunction(extents_fn, compute_fn)
  return wrap


def _require_device_fn(fn, what):
  if not isinstance(fn, DeviceTileFunction):
    raise NotDeviceMappable(
      '%s() needs a device tile function (spartan_b200.expr.map2.device_tile_function); %r is a Python function over '
      'NumPy tiles, which cannot run on the GPU (there is no CPU fallback).  dot / KMeans / sparse dot are provided '
      'as device operations; element-wise work goes through map() over NumPy ufuncs.' % (what, fn))


def _scatter(target, owner, regions, values):
  """target.update(region, value) for results that live on rank ``owner`` (collective)."""
  for i, region in enumerate(regions):
    target.update(region, values[i] if values is not None else None, src=owner)


def join_extents(ex, shapes, axes):
  """The extents join_mapper fetches for tile ``ex`` of the first array (map.py:248-272): the tile re-partitioned along
  axes[0] (extent.pyx:501-570), and for every other array the full-width strip that covers the same index range along
  its own join axis.  ``axes == ()``: the same extent of every array.  None when the re-partitioned extent is empty."""
  if len(axes) == 0:
    return [ex for _ in shapes]
  first = extent.change_partition_axis(ex, axes[0])
  if first is None:
    return None
  k0, k1 = first.ul[axes[0]], first.lr[axes[0]]
  out = [first]
  for i in range(1, len(shapes)):
    ul = [0] * len(shapes[i])
    lr = list(shapes[i])
    ul[axes[i]], lr[axes[i]] = k0, k1
    e = extent.create(ul, lr, shapes[i])
    if e is None:
      return None
    out.append(e)
  return out


class Map2Expr(Expr):
  """map.py:289-334."""
  members = ('arrays', 'axes', 'fn', 'fn_kw', 'out_shape', 'tile_hint', 'out_dtype', 'reducer')

  def compute_shape(self):
    return tuple(self.out_shape)

  def __str__(self):
    return 'Map2[%s, axes=%s, %s]' % (self.arrays, self.axes, self.fn)

  def _evaluate(self, ctx, deps):
    arrays = list(deps['arrays'])
    axes, fn, kw = deps['axes'], deps['fn'], deps['fn_kw'] or {}
    dtype = deps['out_dtype'] if deps['out_dtype'] is not None else arrays[0].dtype
    target = distarray.create(deps['out_shape'], dtype, reducer=deps['reducer'], tile_hint=deps['tile_hint'])
    first = arrays[0]
    if not isinstance(first, distarray.DistArrayImpl) and not getattr(first, 'is_view', False):
      raise SpartanError('map2: the first array must be distributed (its tiles drive the join)')
    # join_mapper (map.py:243-286) once per tile of the first array, in the canonical (worker, id) order
    for ex, tid in sorted(first.tiles.items(), key=lambda kv: (kv[1].worker, kv[1].id)):
      owner = tid.worker
      join = join_extents(ex, [a.shape for a in arrays], axes)
      if join is None:
        continue
      tiles = [arrays[i].fetch(join[i], dst=owner) for i in range(len(arrays))]      # collective strip fetches
      regions = list(fn.extents_fn(join, **kw))
      values = None
      if owner == ctx.worker_id:
        values = list(fn.compute_fn(join, tiles, **kw))
        if len(values) != len(regions):
          raise SpartanError('%r produced %d results for %d target extents' % (fn, len(values), len(regions)))
      _scatter(target, owner, regions, values)
    return target


def map2(arrays, axes=(), fn=None, fn_kw=None, shape=None, update_region=None, tile_hint=None, dtype=None,
         reducer=None):
  """Join the tiles of ``arrays`` along ``axes`` and call the device tile function ``fn`` on every join
  (map.py:337-375).  ``shape`` is the shape of the result; ``reducer`` merges overlapping results."""
  if not util.is_iterable(arrays):
    arrays = [arrays]
  if not util.is_iterable(axes):
    axes = [axes]
  assert fn is not None and shape is not None
  assert len(axes) == 0 or len(arrays) == len(axes)
  _require_device_fn(fn, 'map2')
  arrays = TupleExpr(vals=tuple(as_array(a) for a in arrays))
  return Map2Expr(arrays=arrays, axes=tuple(axes), fn=fn, fn_kw=fn_kw, out_shape=tuple(shape), tile_hint=tile_hint,
                  out_dtype=dtype, reducer=reducer)


class OuterProductExpr(Expr):
  """outer.py:62-99."""
  members = ('arrays', 'axes', 'fn', 'fn_kw', 'out_shape', 'tile_hint', 'out_dtype', 'reducer')

  def compute_shape(self):
    return tuple(self.out_shape)

  def _evaluate(self, ctx, deps):
    arrays = list(deps['arrays'])
    axes, fn, kw = deps['axes'], deps['fn'], deps['fn_kw'] or {}
    assert len(arrays) == 2
    dtype = deps['out_dtype'] if deps['out_dtype'] is not None else arrays[0].dtype
    target = distarray.create(deps['out_shape'], dtype, reducer=deps['reducer'], tile_hint=deps['tile_hint'])
    a, b = arrays
    for ex, tid in sorted(a.tiles.items(), key=lambda kv: (kv[1].worker, kv[1].id)):
      owner = tid.worker
      # outer_mapper (outer.py:12-59)
      first_extent = extent.change_partition_axis(ex, axes[0])
      if first_extent is None:
        continue
      first_tile = a.fetch(first_extent, dst=owner)
      if axes[1] is None:
        outer_extents = [extent.from_shape(b.shape)]
      else:
        outer_extents, seen = [], set()
        for key in b.tiles.keys():
          oe = extent.change_partition_axis(key, axes[1])
          if oe is None or (oe.ul, oe.lr) in seen:
            continue
          seen.add((oe.ul, oe.lr))
          outer_extents.append(oe)
      for oe in outer_extents:
        outer_tile = b.fetch(oe, dst=owner)
        regions = list(fn.extents_fn(first_extent, oe, **kw))
        values = None
        if owner == ctx.worker_id:
          values = list(fn.compute_fn(first_extent, first_tile, oe, outer_tile, **kw))
        _scatter(target, owner, regions, values)
    return target


def outer(arrays, axes, fn, fn_kw=None, shape=None, tile_hint=None, reducer=None, dtype=None):
  """Cartesian join of the tiles of two arrays (outer.py:102-120): ``fn`` sees every pair (strip of the first array
  along axes[0], strip of the second along axes[1] -- or the whole second array when axes[1] is None)."""
  assert fn is not None and shape is not None
  _require_device_fn(fn, 'outer')
  arrays = TupleExpr(vals=tuple(as_array(a) for a in arrays))
  return OuterProductExpr(arrays=arrays, axes=tuple(axes), fn=fn, fn_kw=fn_kw, out_shape=tuple(shape),
                          tile_hint=tile_hint, out_dtype=dtype, reducer=reducer)


# ------------------------------------------------------------------------------------ the library's own joins as instances
def _as_matrix(t, rows, cols):
  return t.reshape(rows, cols)


def _dot_map2_extents(extents, is_vec=None):
  """Target extent of dot_map2_mapper (dot.py:195-217)."""
  if is_vec:
    return [extent.create((0,), (extents[1].lr[1],), (extents[1].shape[1],))]
  if len(extents[1].array_shape) == 1:
    return [extent.create((0,), (extents[0].lr[0],), (extents[0].shape[0],))]
  return [extent.create((0, 0), (extents[0].lr[0], extents[1].lr[1]), (extents[0].shape[0], extents[1].shape[1]))]


def _tile_product(a, b):
  """tiles[0].dot(tiles[1]) on the device: tcgen05 for float32 (FLAGS.dot_precision), exact CUDA cores otherwise."""
  import torch
  from ..config import FLAGS
  M, K = a.shape
  N = b.shape[1]
  out_dtype = torch.promote_types(a.dtype, b.dtype)
  a, b = a.to(out_dtype), b.to(out_dtype)
  c = torch.empty((M, N), dtype=out_dtype, device=a.device)
  device_ops.gemm_views(a, b, c, accumulate=False, precision=FLAGS.dot_precision)
  return c


@device_tile_function(_dot_map2_extents)
def dot_map2_mapper(extents, tiles, is_vec=None):
  """dot.py:195-217: the rank-k partial product of one K strip pair, shaped like the whole result."""
  a, b = tiles
  if is_vec:
    a = a.reshape(1, -1)
  vec = b.dim() == 1
  a2 = a if a.dim() == 2 else a.reshape(1, -1)
  b2 = b.reshape(-1, 1) if vec else b
  c = _tile_product(a2, b2)
  if is_vec:
    return [c.reshape(-1)]
  return [c.reshape(-1) if vec else c]


def _dot_outer_extents(ex_a, ex_b):
  if len(ex_b.array_shape) == 1:
    return [extent.create((ex_a.ul[0],), (ex_a.lr[0],), (ex_a.array_shape[0],))]
  return [extent.create((ex_a.ul[0], ex_b.ul[1]), (ex_a.lr[0], ex_b.lr[1]), (ex_a.array_shape[0], ex_b.array_shape[1]))]


@device_tile_function(_dot_outer_extents)
def dot_outer_mapper(ex_a, tile_a, ex_b, tile_b):
  """dot.py:222-238: a row strip of A times the whole of B."""
  vec = tile_b.dim() == 1
  c = _tile_product(tile_a, tile_b.reshape(-1, 1) if vec else tile_b)
  return [c.reshape(-1) if vec else c]


def dot_as_join(a, b, tile_hint=None):
  """``dot`` routed exactly like the reference routes it (dot.py:243-299): 2-D x 2-D through ``outer`` when A has more
  rows than columns, otherwise through ``map2`` on (axis 1 of A, axis 0 of B), partials merged with np.add.  The
  shipped ``dot`` computes the same product owner-computes without moving partials (spartan_b200/expr/dot.py); this
  form exists to show the join operator carries it, and as its parity test."""
  a, b = as_array(a), as_array(b)
  if len(a.shape) == 1 or len(b.shape) == 1:
    raise NotDeviceMappable('dot_as_join covers the 2-D x 2-D routes of dot.py:281-294')
  if a.shape[1] != b.shape[0]:
    raise ValueError('objects are not aligned')
  shape = (a.shape[0], b.shape[1])
  if tile_hint is None:
    tile_hint = shape
  dtype = np.result_type(a.evaluate().dtype if isinstance(a, Expr) else a.dtype, b.evaluate().dtype)
  if a.shape[0] > a.shape[1]:
    return outer((a, b), (0, None), dot_outer_mapper, shape=shape, tile_hint=tile_hint, reducer=np.add, dtype=dtype)
  return map2((a, b), (1, 0), dot_map2_mapper, shape=shape, tile_hint=tile_hint, reducer=np.add, dtype=dtype)
