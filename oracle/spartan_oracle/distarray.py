"""ORACLE (test infrastructure only) -- tiles, the combiner, and the distributed array model.

Restates spartan/array/tile.pyx (dense + 0-d paths of ``merge``), spartan/array/distarray.py
(tiling, round-robin placement, fetch stitching, update splitting) and
spartan/expr/operator/broadcast.py in one process.  "Workers" are just integer ids; kernels over
tiles run sequentially in (worker id, creation order), which fixes the combiner order that the
reference leaves to RPC arrival order.
"""
import collections
import itertools

import numpy as np

from . import extent


# --------------------------------------------------------------------------- context
class Ctx(object):
  """Stands in for spartan/blob_ctx.py BlobCtx: a tile store keyed by (worker, id)."""

  def __init__(self, num_workers=3):
    self.num_workers = int(num_workers)     # config.py:131 default num_workers=3
    self.tiles = {}
    self._next = 0
    self._rr = 0

  def create(self, tile, hint=-1):
    # blob_ctx.py:221-254 : hint >= 0 -> worker = hint % num_workers, else round-robin
    if hint is None or hint < 0:
      worker = self._rr % self.num_workers
      self._rr += 1
    else:
      worker = hint % self.num_workers
    tid = (worker, self._next)
    self._next += 1
    self.tiles[tid] = tile
    return tid


_ctx = Ctx(3)


def get_ctx():
  return _ctx


def initialize(num_workers=3):
  global _ctx
  _ctx = Ctx(num_workers)
  return _ctx


# --------------------------------------------------------------------------- tile.pyx
class Tile(object):
  """tile.pyx:24-142, dense mode only (masked-get and sparse modes are out of scope)."""

  def __init__(self, shape, dtype, data=None, mask_all_set=False):
    self.shape = tuple(shape)
    self.dtype = np.dtype(dtype)
    self.data = data
    # mask: False = nothing written yet (MASK_ALL_CLEAR), True = fully written, ndarray = partial
    self.mask = True if mask_all_set else False

  def get(self, subslice=None):
    # tile.pyx:68-112
    if self.data is None:
      return np.ndarray(self.shape, self.dtype)[subslice]     # uninitialised, like the reference (:72-79)
    if len(self.data.shape) == 0:
      return self.data
    return self.data[subslice]

  def _mask_array(self):
    if not isinstance(self.mask, np.ndarray):
      self.mask = np.ones(self.shape, dtype=bool) if self.mask else np.zeros(self.shape, dtype=bool)
    return self.mask

  def update(self, subslice, data, reducer):
    return merge(self, subslice, data, reducer)


def from_data(data):
  data = np.asarray(data)
  return Tile(data.shape, data.dtype, data, mask_all_set=True)     # tile.pyx:145-160


def from_shape(shape, dtype):
  return Tile(shape, dtype, None, mask_all_set=False)              # tile.pyx:163-179


def merge(old_tile, subslice, update, reducer):
  """tile.pyx:200-297 -- the combiner.  First write to a region replaces, later writes
  ``reducer(old, new)``; the full-tile fast path casts to the tile dtype."""
  update = np.asarray(update)
  if len(old_tile.shape) == 0:                                     # :212-217
    if old_tile.data is None or reducer is None:
      old_tile.data = update
    else:
      old_tile.data = reducer(old_tile.data, update)
    return old_tile

  if old_tile.data is None:                                        # :203-206 _initialize
    old_tile.data = np.zeros(old_tile.shape, dtype=old_tile.dtype)
  mask = old_tile._mask_array()

  if old_tile.data.shape == update.shape:                          # :263-268
    if reducer is not None and mask[np.unravel_index(0, old_tile.data.shape)]:
      old_tile.data = np.asarray(reducer(old_tile.data, update))
    else:
      old_tile.data = update.astype(old_tile.data.dtype)
    old_tile.mask = np.ones(old_tile.shape, dtype=bool)
  else:                                                            # :270-283
    replaced = ~mask[subslice]
    updated = mask[subslice]
    old_region = old_tile.data[subslice]
    if np.any(replaced):
      old_region[replaced] = update[replaced]
    if np.any(updated):
      if reducer is not None:
        old_region[updated] = reducer(old_region[updated], update[updated])
      else:
        old_region[updated] = update[updated]
    mask[subslice] = True
  return old_tile


# --------------------------------------------------------------------------- distarray.py tiling
DEFAULT_TILE_SIZE = 100000


def good_tile_shape(shape, num_shards=-1):
  # distarray.py:26-48 (Python-2 integer division)
  if num_shards != -1:
    tile_size = int(np.prod(shape, dtype=np.int64)) // num_shards
  else:
    tile_size = DEFAULT_TILE_SIZE
  tile_shape = [1] * len(shape)
  idx = len(shape) - 1
  while tile_size > 1:
    tile_shape[idx] = min(shape[idx], tile_size)
    tile_size //= shape[idx]
    idx -= 1
  return tile_shape


def compute_splits(shape, tile_hint):
  # distarray.py:51-71
  splits = [None] * len(shape)
  for dim in range(len(shape)):
    step = tile_hint[dim]
    splits[dim] = [(i, min(shape[dim], i + step)) for i in range(0, shape[dim], step)]
  return splits


def compute_extents(shape, tile_hint=None, num_shards=-1):
  # distarray.py:73-110 : ordered dict extent -> shard index, itertools.product (row-major) order
  if len(shape) == 0:
    return collections.OrderedDict([(extent.create([], [], ()), 0)])
  if tile_hint is None:
    tile_hint = good_tile_shape(shape, num_shards)
  else:
    assert len(tile_hint) == len(shape), '#dimensions in tile hint does not match shape'
  splits = compute_splits(shape, tile_hint)
  result = collections.OrderedDict()
  idx = 0
  for slc in itertools.product(*splits):
    if num_shards != -1:
      idx = idx % num_shards
    ul, lr = zip(*slc)
    result[extent.create(ul, lr, shape)] = idx
    idx += 1
  return result


# --------------------------------------------------------------------------- DistArray
class DistArray(object):
  def real_size(self):
    return int(np.prod(self.shape, dtype=np.int64))

  @property
  def ndim(self):
    return len(self.shape)

  def select(self, idx):
    if isinstance(idx, extent.TileExtent):
      return self.fetch(idx)
    return self.fetch(extent.from_slice(idx, self.shape))

  def glom(self):
    return self.select(np.index_exp[:])

  def map_to_array(self, mapper_fn, kw=None):
    # distarray.py:202-208
    results = self.foreach_tile(mapper_fn=mapper_fn, kw=kw)
    extents = collections.OrderedDict()
    for d in results:
      for ex, tid in d:
        extents[ex] = tid
    return from_table(extents)


def _tile_order(tiles):
  # Worker._run_kernel pops its local tiles (worker.py:248-263); workers run concurrently.  The oracle
  # serialises them as (worker id, creation order).
  return sorted(tiles.items(), key=lambda kv: (kv[1][0], kv[1][1]))


class DistArrayImpl(DistArray):
  def __init__(self, shape, dtype, tiles, reducer_fn, sparse=False):
    self.shape = tuple(shape)
    self.dtype = np.dtype(dtype)
    self.tiles = tiles              # extent -> tile id
    self.reducer_fn = reducer_fn
    self.sparse = sparse
    self.ctx = get_ctx()

  def tile_shape(self):
    # distarray.py:275-281
    scounts = collections.defaultdict(int)
    for ex in self.tiles:
      scounts[ex.shape] += 1
    return sorted(scounts.items(), key=lambda kv: (kv[1], kv[0]))[-1][0]

  def foreach_tile(self, mapper_fn, kw=None):
    # distarray.py:283-292 -> BlobCtx.map -> Worker._run_kernel
    kw = dict(kw or {})
    return [mapper_fn(ex, **kw) for ex, _ in _tile_order(self.tiles)]

  def fetch(self, region):
    # distarray.py:294-367 (dense)
    assert region.array_shape == self.shape, (region.array_shape, self.shape)
    assert all(l <= s for l, s in zip(region.lr, self.shape)), 'Requested region is out of bounds'
    if region in self.tiles:
      return self.ctx.tiles[self.tiles[region]].get(extent.offset_slice(region, region))
    splits = list(extent.find_overlapping(self.tiles.keys(), region))
    results = [self.ctx.tiles[self.tiles[ex]].get(extent.offset_slice(ex, inter)) for ex, inter in splits]
    if len(splits) == 1:
      return results[0]
    tgt = np.ndarray(region.shape, dtype=self.dtype)
    for (ex, inter), result in zip(splits, results):
      dst_slice = extent.offset_slice(region, inter)
      if extent.all_nonzero_shape(result.shape):
        tgt[dst_slice] = result
    return tgt

  def update(self, region, data, wait=True):
    # distarray.py:372-422
    data = np.asarray(data)
    assert region.shape == data.shape, 'Size of extent does not match size of data %s %s' % (region.shape, data.shape)
    if region in self.tiles:
      tile = self.ctx.tiles[self.tiles[region]]
      tile.update(extent.offset_slice(region, region), data, self.reducer_fn)
      return
    slices = []
    if region.shape == self.shape:
      for ex, tid in self.tiles.items():
        slices.append((tid, ex.to_slice(), extent.offset_slice(ex, ex)))
    else:
      for dst_extent, inter in extent.find_overlapping(self.tiles, region):
        src_slice = extent.offset_slice(region, inter)
        dst_slice = extent.offset_slice(dst_extent, inter)
        if extent.all_nonzero_shape([s.stop - s.start for s in dst_slice]):
          slices.append((self.tiles[dst_extent], src_slice, dst_slice))
    slices.sort(key=lambda x: x[1][0].start if len(x[1]) else 0)
    for tid, src_slice, dst_slice in slices:                      # sparse.pyx:281-301 multiple_slice (dense)
      self.ctx.tiles[tid].update(dst_slice, data[src_slice], self.reducer_fn)



def create(shape, dtype=np.float64, sharder=None, reducer=None, tile_hint=None, sparse=False):
  # distarray.py:425-487, tile_assignment_strategy == 'round_robin' (:441-445)
  ctx = get_ctx()
  dtype = np.dtype(dtype)
  shape = tuple(shape)
  extents = compute_extents(shape, tile_hint, ctx.num_workers)
  tiles = collections.OrderedDict()
  for ex, i in extents.items():
    tiles[ex] = ctx.create(from_shape(ex.shape, dtype), hint=i)
  return DistArrayImpl(shape=shape, dtype=dtype, tiles=tiles, reducer_fn=reducer, sparse=sparse)


def from_table(extents):
  # distarray.py:519-550
  ctx = get_ctx()
  if not extents:
    return DistArrayImpl(shape=(), dtype=np.float64, tiles=extents, reducer_fn=None)
  shape = extent.find_shape(list(extents.keys()))
  tid = next(iter(extents.values()))
  dtype = ctx.tiles[tid].dtype
  return DistArrayImpl(shape=shape, dtype=dtype, tiles=extents, reducer_fn=None)


class LocalWrapper(DistArray):
  # distarray.py:553-602
  def __init__(self, data):
    self._data = np.asarray(data)
    self.sparse = False
    self._ex = extent.from_slice(np.index_exp[:], self.shape)

  @property
  def dtype(self):
    return self._data.dtype

  @property
  def shape(self):
    return self._data.shape

  @property
  def tiles(self):
    return {self._ex: (-1, 0)}

  def fetch(self, ex):
    return self._data[ex.to_slice()]

  def foreach_tile(self, mapper_fn, kw=None):
    # distarray.py:587-602 : result of the single mapper call, re-wrapped
    result = mapper_fn(self._ex, **(kw or {}))
    assert len(result) == 1
    _, tid = result[0]
    return as_array(get_ctx().tiles[tid].get(slice(None, None, None)))

  def map_to_array(self, mapper_fn, kw=None):
    return self.foreach_tile(mapper_fn=mapper_fn, kw=kw)


def as_array(data):
  return data if isinstance(data, DistArray) else LocalWrapper(data)


def largest_value(vals):
  return max(vals, key=lambda v: v.real_size())                  # distarray.py:636-642


# --------------------------------------------------------------------------- broadcast.py
class Broadcast(DistArray):
  # broadcast.py:28-109
  def __init__(self, base, shape):
    self.base = base.base if isinstance(base, Broadcast) else base
    self.shape = tuple(shape)
    self.tiles = self.base.tiles
    self.dtype = base.dtype
    self.sparse = False
    self.prepend_dim = len(shape) - len(base.shape)

  def real_size(self):
    return int(np.prod(self.base.shape, dtype=np.int64)) - 1      # :56-61

  def foreach_tile(self, mapper_fn, kw=None):
    # broadcast.py:8-26,63-72 : iterate base tiles, expand extent along broadcast dims
    out = []
    kw = dict(kw or {})
    for base_ex, _ in _tile_order(self.base.tiles):
      ul = [0] * len(self.shape)
      lr = list(self.shape)
      for i in range(len(base_ex.ul) - 1, -1, -1):
        bi = i + self.prepend_dim
        if self.base.shape[i] == self.shape[bi]:
          ul[bi] = base_ex.ul[i]
          lr[bi] = base_ex.lr[i]
      out.append(mapper_fn(extent.create(ul, lr, self.shape), **kw))
    return out

  def _base_ex(self, ex):
    # broadcast.py:74-92
    while len(ex.shape) > len(self.base.shape):
      ex = extent.drop_axis(ex, 0)
    ul, lr = [], []
    for i in range(len(self.base.shape)):
      if self.base.shape[i] == 1:
        ul.append(0); lr.append(1)
      else:
        ul.append(ex.ul[i]); lr.append(ex.lr[i])
    return extent.create(ul, lr, self.base.shape)

  def fetch(self, ex):
    template = np.ndarray(ex.shape, dtype=self.base.dtype)
    fetched = self.base.fetch(self._base_ex(ex))
    _, bcast = np.broadcast_arrays(template, fetched)
    return bcast

  def fetch_base_tile(self, ex):
    return self.base.fetch(self._base_ex(ex))


def broadcast(args):
  # broadcast.py:111-158
  if len(args) == 1:
    return args
  orig_shapes = [list(x.shape) for x in args]
  max_dim = max(len(s) for s in orig_shapes)
  new_shapes = [[1] * (max_dim - len(s)) + s for s in orig_shapes]
  for axis in range(max_dim):
    axis_shape = set(shp[axis] for shp in new_shapes)
    assert len(axis_shape) <= 2, 'Mismatched shapes for broadcast: %s' % orig_shapes
    if len(axis_shape) == 2:
      assert 1 in axis_shape, 'Mismatched shapes for broadcast: %s' % orig_shapes
    max_size = max(shp[axis] for shp in new_shapes)
    for shp in new_shapes:
      shp[axis] = max_size
  results = []
  for i in range(len(args)):
    if new_shapes[i] == orig_shapes[i]:
      results.append(args[i])
    else:
      results.append(Broadcast(args[i], tuple(new_shapes[i])))
  return results
