"""Distributed arrays whose tiles are resident in GPU HBM.

Keeps the interface of spartan/array/distarray.py (DistArray.fetch / update / foreach_tile /
map_to_array / glom / tiles / tile_shape, create, from_table, LocalWrapper, as_array,
largest_value, good_tile_shape, compute_extents) and of spartan/expr/operator/broadcast.py
(Broadcast, broadcast).  Differences that come from being B200-native:

  * ``fetch`` returns a *device* tensor (a zero-copy view whenever the region lies inside one local
    tile or one local slab); ``glom`` is the only call that copies to the host;
  * the local tiles of an array are carved out of ONE HBM allocation per rank (a "slab") whenever the
    rank's tiles form a cartesian product of per-axis intervals (always true for one rank and for
    round-robin placement of row / column / evenly divisible grid tilings).  A slab lets a tiled
    ``dot`` run as one launch over the rank's whole share instead of one launch per tile pair;
  * tiling (``compute_extents``) and placement (``idx % num_workers``, distarray.py:99-108,441-445)
    are bit-identical to the reference and computed in the native shim.
"""
import collections
import math
import ctypes
import itertools

import numpy as np
import torch

from . import extent, tile
from .. import blob_ctx, comm, device_ops
from .._lib import lib, check, i64arr, SpartanError
from ..core import TileId
from ..util import require_type, require_equal, require_unique

DEFAULT_TILE_SIZE = 100000   # distarray.py:20
_serials = itertools.count(1)


def good_tile_shape(shape, num_shards=-1):
  """distarray.py:26-48 (native: sp_good_tile_shape)."""
  n = len(shape)
  out = (ctypes.c_int64 * max(1, n))()
  check(lib.sp_good_tile_shape(n, i64arr(shape), int(num_shards), out), 'good_tile_shape')
  return [out[i] for i in range(n)]


_EXTENTS_MEMO = collections.OrderedDict()      # (shape, hint, shards) -> ((extent, shard), ...); extents are immutable
_EXTENTS_MEMO_ENTRIES = 256
_EXTENTS_MEMO_MAX_TILES = 4096


def compute_extents(shape, tile_hint=None, num_shards=-1):
  """distarray.py:73-110: ordered {extent: shard index}, itertools.product order (native).  The tiling of a shape is a
  pure function of (shape, hint, shards), and iterative programs ask for the same few over and over: recent answers
  are kept (the extent objects are shared, the dict is the caller's own)."""
  shape = tuple(int(s) for s in shape)
  n = len(shape)
  if n == 0:
    return collections.OrderedDict([(extent.create([], [], ()), 0)])
  if tile_hint is not None:
    require_equal(len(tile_hint), n, '#dimensions in tile hint does not match shape %s vs %s' % (tile_hint, shape))
  key = (shape, None if tile_hint is None else tuple(int(h) for h in tile_hint), int(num_shards))
  hit = _EXTENTS_MEMO.get(key)
  if hit is not None:
    _EXTENTS_MEMO.move_to_end(key)
    return collections.OrderedDict(hit)
  result = _compute_extents(shape, n, tile_hint, num_shards)
  if len(result) <= _EXTENTS_MEMO_MAX_TILES:
    _EXTENTS_MEMO[key] = tuple(result.items())
    if len(_EXTENTS_MEMO) > _EXTENTS_MEMO_ENTRIES:
      _EXTENTS_MEMO.popitem(last=False)
  return result


def _compute_extents(shape, n, tile_hint, num_shards):
  hint = None if tile_hint is None else i64arr(tile_hint)
  total = check(lib.sp_compute_extents(n, i64arr(shape), hint, int(num_shards), None, None, None), 'compute_extents')
  ul = (ctypes.c_int64 * max(1, total * n))(); lr = (ctypes.c_int64 * max(1, total * n))()
  wk = (ctypes.c_int64 * max(1, total))()
  check(lib.sp_compute_extents(n, i64arr(shape), hint, int(num_shards), ul, lr, wk), 'compute_extents')
  result = collections.OrderedDict()
  for t in range(total):
    ex = extent.TileExtent([ul[t * n + i] for i in range(n)], [lr[t * n + i] for i in range(n)], shape)
    result[ex] = int(wk[t])
  return result


def _tile_mapper(tile_id, blob, array=None, user_fn=None, **kw):
  """distarray.py:113-116."""
  ex = array.extent_for_blob(tile_id)
  return user_fn(ex, **kw)


class DistArray(object):
  """distarray.py:119-215 -- the interface every array-like exposes to the evaluator."""

  def fetch(self, ex, dst=None):
    raise NotImplementedError

  def update(self, ex, data, wait=True):
    raise NotImplementedError

  def foreach_tile(self, mapper_fn, kw=None):
    raise NotImplementedError

  def extent_for_blob(self, id):
    raise NotImplementedError

  def real_size(self):
    return math.prod(self.shape)

  def __len__(self):
    return self.shape[0]

  def __repr__(self):
    return '%s(id=%s, shape=%s, dtype=%s)' % (self.__class__.__name__, id(self), self.shape, self.dtype)

  def select(self, idx):
    if isinstance(idx, extent.TileExtent):
      return self.fetch(idx)
    return self.fetch(extent.from_slice(idx, self.shape))

  def __getitem__(self, idx):
    return self.select(idx)

  def glom(self):
    raise NotImplementedError

  def map_to_array(self, mapper_fn, kw=None):
    """distarray.py:202-208."""
    results = self.foreach_tile(mapper_fn=mapper_fn, kw=kw)
    extents = collections.OrderedDict()
    for tile_id in sorted(results, key=lambda t: (t.worker, t.id)):
      for ex, tid in results[tile_id].result:
        extents[ex] = tid
    return from_table(extents)

  def __hash__(self):
    return id(self)

  @property
  def ndim(self):
    return len(self.shape)


def _product_layout(local_extents, ndim):
  """If the extents form a cartesian product of per-axis intervals, returns per-axis sorted interval
  lists; otherwise None."""
  if not local_extents or ndim == 0:
    return None
  axes = []
  count = 1
  for d in range(ndim):
    ivs = sorted(set((ex.ul[d], ex.lr[d]) for ex in local_extents))
    for (a0, a1), (b0, b1) in zip(ivs, ivs[1:]):
      if b0 < a1:
        return None          # overlapping intervals on one axis: not a product of disjoint intervals
    axes.append(ivs)
    count *= len(ivs)
  if count != len(local_extents):
    return None
  return axes


class DistArrayImpl(DistArray):
  def __init__(self, shape, dtype, tiles, reducer_fn, sparse=False):
    if sparse:
      raise SpartanError('sparse arrays are out of scope of the device evaluator')
    self.shape = tuple(int(s) for s in shape)
    self.dtype = np.dtype(dtype)
    self.reducer_fn = reducer_fn
    self.sparse = False
    self.bad_tiles = []
    self.ctx = blob_ctx.get()
    require_type(tiles, dict)
    self.tiles = tiles                      # extent -> TileId, ALL tiles of the array (every rank knows them)
    self.blob_to_ex = dict((v, k) for k, v in tiles.items())
    self.slab = None                        # this rank's tiles as one HBM allocation (or None)
    self.slab_axes = None                   # per-axis sorted (lo, hi) intervals covered by the slab
    self.block_events = None                # [(region, CUDA event)] left by a producer that finished block by block
    self.serial = next(_serials)            # identity of this array for caches of derived data (never reused)
    self._layout_key = None

  def __del__(self):
    try:
      device_ops.prepared_cache.drop(self.serial)
      self.ctx.destroy_all(list(self.tiles.values()))
    except Exception:
      pass

  def extent_for_blob(self, id):
    return self.blob_to_ex[id]

  def tile_shape(self):
    """distarray.py:275-281: the most common tile shape."""
    scounts = collections.defaultdict(int)
    for ex in self.tiles:
      scounts[ex.shape] += 1
    return sorted(scounts.items(), key=lambda kv: (kv[1], kv[0]))[-1][0]

  def local_extents(self):
    me = self.ctx.worker_id
    return [ex for ex, tid in self.tiles.items() if tid.worker == me]

  def foreach_tile(self, mapper_fn, kw=None):
    """distarray.py:283-292 -> BlobCtx.map."""
    kw = dict(kw or {})
    kw['array'] = self
    kw['user_fn'] = mapper_fn
    return self.ctx.map(list(self.tiles.values()), mapper_fn=_tile_mapper, kw=kw)

  # ------------------------------------------------------------------ slab helpers
  def slab_view(self, region):
    """Zero-copy view of ``region`` inside this rank's slab, or None if the region is not covered by
    one contiguous run of slab intervals on every axis."""
    if self.slab is None:
      return None
    slices = []
    for d, ivs in enumerate(self.slab_axes):
      off = 0
      lo = hi = None
      pos = region.ul[d]
      for a, b in ivs:
        if lo is None:
          if a <= pos < b:
            lo = off + (pos - a)
            if region.lr[d] <= b:
              hi = off + (region.lr[d] - a)
              break
            pos = b
        else:
          if a != pos:
            return None      # gap: the run is not contiguous in the slab
          if region.lr[d] <= b:
            hi = off + (region.lr[d] - a)
            break
          pos = b
        off += b - a
      if lo is None or hi is None:
        return None
      slices.append(slice(lo, hi))
    return self.slab[tuple(slices)]

  def local_blocks(self):
    """This rank's share as few rectangular regions as possible: the product of the contiguous runs of
    its slab intervals (each block is addressable as ONE zero-copy slab view), or its tiles one by one
    when there is no slab.  One kernel launch per block instead of one per tile."""
    me = self.ctx.worker_id
    if self.slab is None or not self.shape:
      return [ex for ex, tid in self.tiles.items() if tid.worker == me]
    import itertools
    runs = []
    for ivs in self.slab_axes:
      axis_runs = []
      for a, b in ivs:
        if axis_runs and axis_runs[-1][1] == a:
          axis_runs[-1] = (axis_runs[-1][0], b)
        else:
          axis_runs.append((a, b))
      runs.append(axis_runs)
    return [extent.create([a for a, _ in blk], [b for _, b in blk], self.shape) for blk in itertools.product(*runs)]

  def same_layout(self, other):
    """True when ``other`` has exactly this array's tiles on the same owners (then every block of one is a
    zero-copy block of the other)."""
    if not isinstance(other, DistArrayImpl) or other.shape != self.shape or len(other.tiles) != len(self.tiles):
      return False
    if other.slab is None or self.slab is None:
      return False
    return other is self or self.layout_key() == other.layout_key()

  def layout_key(self):
    """The set of (tile bounds, owner): the tile table of an array never changes after construction, so it is built
    once."""
    if self._layout_key is None:
      self._layout_key = frozenset((ex.ul, ex.lr, tid.worker) for ex, tid in self.tiles.items())
    return self._layout_key

  # ------------------------------------------------------------------ fetch / update / glom
  def fetch(self, region, dst=None):
    """Device tensor holding ``region`` (distarray.py:294-367).

    ``dst`` = rank that needs the data (default: the caller, which then must own every overlapping
    tile).  With dst given the call is collective: every rank calls it with the same arguments, owners
    of overlapping tiles send their rectangles over NCCL, and only rank ``dst`` gets a tensor back."""
    require_type(region, extent.TileExtent)
    require_equal(region.array_shape, self.shape)
    ctx = self.ctx
    me = ctx.worker_id
    want = me if dst is None else dst
    tid = self.tiles.get(region)
    if tid is not None and tid.worker == want:
      return ctx.get(tid, None) if want == me else None
    if want == me:
      view = self.slab_view(region)
      if view is not None:
        return view
    splits = list(extent.find_overlapping(self.tiles.keys(), region))
    if dst is None and ctx.num_workers > 1:
      remote = [ex for ex, _ in splits if self.tiles[ex].worker != me]
      if remote:
        raise SpartanError('fetch(%s) without dst= touches %d tile(s) owned by other ranks; remote pieces travel only in '
                           'the collective form fetch(region, dst=rank) that every rank calls' % (region, len(remote)))
    if len(splits) == 1 and self.tiles[splits[0][0]].worker == want:
      ex, inter = splits[0]
      return ctx.get(self.tiles[ex], extent.offset_slice(ex, inter)) if want == me else None
    return self._fetch_pieces(region, splits, want)

  def _fetch_pieces(self, region, splits, want):
    ctx = self.ctx
    me = ctx.worker_id
    tgt = ctx.empty(region.shape if len(self.shape) else (), self.dtype) if want == me else None
    for ex, inter in splits:
      tid = self.tiles[ex]
      src_slice = extent.offset_slice(ex, inter)
      dst_slice = extent.offset_slice(region, inter)
      if tid.worker == want:
        if want == me:
          piece = ctx.get(tid, src_slice)
          if tgt.dim() == 0:
            tgt.copy_(piece)
          else:
            device_ops.copy_rect(tgt[dst_slice], piece)
      elif tid.worker == me:            # I own a piece somebody else needs
        piece = ctx.get(tid, src_slice).contiguous()
        comm.batch_p2p([('send', piece, want)])
      elif want == me:
        tmp = ctx.empty(inter.shape, self.dtype)
        comm.batch_p2p([('recv', tmp, tid.worker)])
        device_ops.copy_rect(tgt[dst_slice], tmp)
    return tgt

  def update_slice(self, slc, data):
    return self.update(extent.from_slice(slc, self.shape), data)

  def _update_from(self, region, data, src):
    """Collective form of ``update``: ``data`` (a device tensor covering ``region``) lives on rank ``src``; every rank
    calls this with the same region, pieces that land in tiles of other ranks travel point-to-point (cast to this
    array's dtype) and are merged by the tile's owner -- the RPC ``update`` of the reference (blob_ctx.py:163-179)."""
    ctx = self.ctx
    me = ctx.worker_id
    ctx.touch()
    self.block_events = None
    if len(self.shape) == 0:
      pieces = [(next(iter(self.tiles.keys())), region)]
    else:
      pieces = list(extent.find_overlapping(self.tiles.keys(), region))
    for dst_extent, inter in pieces:
      tid = self.tiles[dst_extent]
      src_slice = extent.offset_slice(region, inter)
      dst_slice = extent.offset_slice(dst_extent, inter)
      if len(self.shape) and not extent.all_nonzero_shape([s.stop - s.start for s in dst_slice]):
        continue
      full = len(self.shape) == 0 or tuple(inter.shape) == tuple(dst_extent.shape)
      if tid.worker == src:
        if me == src:
          piece = data[src_slice] if data.dim() else data
          ctx.update(tid, None if full else dst_slice, piece, self.reducer_fn)
      elif me == src:
        piece = (data[src_slice] if data.dim() else data).to(blob_ctx.torch_dtype(self.dtype)).contiguous()
        comm.batch_p2p([('send', piece, tid.worker)])
      elif me == tid.worker:
        tmp = ctx.empty(inter.shape if len(self.shape) else (), self.dtype)
        comm.batch_p2p([('recv', tmp, src)])
        ctx.update(tid, None if full else dst_slice, tmp, self.reducer_fn)
    return None

  def update(self, region, data, wait=True, src=None):
    """distarray.py:372-422.  ``data`` is either a host ndarray that every rank holds (each rank
    uploads the parts that land in its own tiles) or a device tensor, in which case every tile it
    overlaps must be local to the caller -- unless ``src`` names the rank that holds it: then the call is
    collective and remote pieces are sent to their owners (see _update_from)."""
    require_type(region, extent.TileExtent)
    if src is not None:
      return self._update_from(region, data, src)
    host = isinstance(data, np.ndarray)
    if not host and not torch.is_tensor(data):
      data = np.asarray(data); host = True
    require_equal(tuple(region.shape), tuple(data.shape), 'Size of extent does not match size of data')
    ctx = self.ctx
    me = ctx.worker_id
    self.block_events = None
    ctx.touch()
    if (host and self.reducer_fn is None and self.slab is not None and len(self.shape) > 0
        and tuple(region.shape) == self.shape and data.dtype == self.dtype):
      # whole-array upload: one H2D copy per contiguous block of this rank's slab instead of one per tile
      import itertools
      runs = []
      for ivs in self.slab_axes:
        axis_runs, off = [], 0
        for a, b in ivs:
          if axis_runs and axis_runs[-1][1] == a:
            axis_runs[-1] = (axis_runs[-1][0], b, axis_runs[-1][2])
          else:
            axis_runs.append((a, b, off))
          off += b - a
        runs.append(axis_runs)
      for block in itertools.product(*runs):
        g = tuple(slice(a, b) for a, b, _ in block)
        l = tuple(slice(o, o + (b - a)) for a, b, o in block)
        device_ops.upload_rect(self.slab[l], data[g])      # pitched DMA straight from the (pinned) host array
      for tid in self.tiles.values():
        if tid.worker == me:
          ctx.tile(tid).valid = True
      return None
    if len(self.shape) == 0:
      pieces = [(next(iter(self.tiles.keys())), region)]
    else:
      pieces = list(extent.find_overlapping(self.tiles.keys(), region))
    for dst_extent, inter in pieces:
      tid = self.tiles[dst_extent]
      if tid.worker != me:
        if not host:
          raise SpartanError('update() with a device tensor touches a tile owned by rank %d' % tid.worker)
        continue
      src_slice = extent.offset_slice(region, inter)
      dst_slice = extent.offset_slice(dst_extent, inter)
      if not extent.all_nonzero_shape([s.stop - s.start for s in dst_slice]):
        continue
      if host:
        piece = torch.from_numpy(np.ascontiguousarray(data[src_slice] if data.ndim else data))
        piece = piece.to(ctx.device, non_blocking=False)
      else:
        piece = data[src_slice] if data.dim() else data
      full = tuple(piece.shape) == tuple(dst_extent.shape) or len(self.shape) == 0
      ctx.update(tid, None if full else dst_slice, piece, self.reducer_fn)
    return None

  def read_local_into(self, out):
    """Copies THIS rank's share into the matching region of the full-size host array ``out`` (pitched D2H per
    contiguous slab block; asynchronous when ``out`` is pinned: synchronise before reading).  Returns bytes."""
    ctx = self.ctx
    total = 0
    if self.block_events and self.slab is not None and ctx.device.type == 'cuda':
      # the producer (a streamed dot) finished the array block by block: copy each block out on a side stream as soon
      # as its event fires, while later blocks are still being computed on the main stream
      main = torch.cuda.current_stream(ctx.device)
      d2h = ctx.side_stream('d2h')
      events, self.block_events = self.block_events, None
      with torch.cuda.stream(d2h):
        for region, ev in events:
          d2h.wait_event(ev)
          view = self.slab_view(region)
          device_ops.download_rect(out[region.to_slice()], view)
          total += view.numel() * view.element_size()
      main.wait_stream(d2h)
      return total
    if self.slab is not None and self.shape:
      for block in self.local_blocks():
        view = self.slab_view(block)
        device_ops.download_rect(out[block.to_slice()], view)
        total += view.numel() * view.element_size()
      return total
    for ex, tid in self.tiles.items():
      if ctx.is_local(tid):
        t = ctx.get(tid, None)
        device_ops.download_rect(out[ex.to_slice()] if self.shape else out, t)
        total += t.numel() * t.element_size()
    return total

  def glom(self):
    """Gathers the whole array to host memory on every rank (distarray.py:198-200)."""
    ctx = self.ctx
    me = ctx.worker_id
    out = np.empty(self.shape, dtype=self.dtype)
    if ctx.num_workers == 1 and self.slab is not None and self.slab.shape == tuple(self.shape):
      return self.slab.cpu().numpy()
    for ex, tid in sorted(self.tiles.items(), key=lambda kv: (kv[1].worker, kv[1].id)):
      if ctx.num_workers == 1:
        data = ctx.get(tid, None)
      else:
        data = ctx.get(tid, None).contiguous() if tid.worker == me else ctx.empty(ex.shape if self.shape else (), self.dtype)
        comm.broadcast(data, tid.worker)
      host = data.cpu().numpy()
      if len(self.shape) == 0:
        out[()] = host
      else:
        out[ex.to_slice()] = host.reshape(ex.shape)
    return out


def create(shape, dtype=np.float64, sharder=None, reducer=None, tile_hint=None, sparse=False):
  """Make a new, empty DistArray (distarray.py:425-487, round_robin placement :441-445)."""
  ctx = blob_ctx.get()
  dtype = np.dtype(dtype)
  shape = tuple(int(s) for s in shape)
  extents = compute_extents(shape, tile_hint, ctx.num_workers)
  me = ctx.worker_id
  local = [ex for ex, i in extents.items() if i % ctx.num_workers == me]
  axes = _product_layout(local, len(shape))
  slab = None
  if axes is not None:
    slab = ctx.empty([sum(b - a for a, b in ivs) for ivs in axes], dtype)
  tiles = collections.OrderedDict()
  for ex, i in extents.items():
    def factory(ex=ex):
      data = None
      if slab is not None:
        slices = []
        for d, ivs in enumerate(axes):
          off = 0
          for a, b in ivs:
            if (a, b) == (ex.ul[d], ex.lr[d]):
              slices.append(slice(off, off + (b - a)))
              break
            off += b - a
        data = slab[tuple(slices)]
      return tile.DeviceTile(ex.shape if shape else (), dtype, data, valid=False)
    tiles[ex] = ctx.create(factory, hint=i)
  array = DistArrayImpl(shape=shape, dtype=dtype, tiles=tiles, reducer_fn=reducer, sparse=sparse)
  array.slab = slab
  array.slab_axes = axes
  return array


def create_like(src, dtype, reducer=None):
  """A new, empty array with the tiling *and placement* of ``src`` (every output tile of a map lives on
  the rank of the input tile it was computed from, like tile_mapper's ctx.create in map.py:84-86), with
  this rank's share carved out of one slab."""
  ctx = blob_ctx.get()
  dtype = np.dtype(dtype)
  me = ctx.worker_id
  local = [ex for ex, tid in src.tiles.items() if tid.worker == me]
  axes = _product_layout(local, len(src.shape))
  slab = ctx.empty([sum(b - a for a, b in ivs) for ivs in axes], dtype) if axes is not None else None
  tiles = collections.OrderedDict()
  for ex, tid in sorted(src.tiles.items(), key=lambda kv: (kv[1].worker, kv[1].id)):
    def factory(ex=ex):
      data = None
      if slab is not None:
        slices = []
        for d, ivs in enumerate(axes):
          off = 0
          for a, b in ivs:
            if (a, b) == (ex.ul[d], ex.lr[d]):
              slices.append(slice(off, off + (b - a)))
              break
            off += b - a
        data = slab[tuple(slices)]
      return tile.DeviceTile(ex.shape if src.shape else (), dtype, data, valid=False)
    tiles[ex] = ctx.create(factory, hint=tid.worker)
  array = DistArrayImpl(shape=src.shape, dtype=dtype, tiles=tiles, reducer_fn=reducer)
  array.slab = slab
  array.slab_axes = axes
  return array


def from_table(extents):
  """distarray.py:519-550: shape = max extent corner, dtype from the (local) tiles."""
  ctx = blob_ctx.get()
  require_unique(extents)
  if not extents:
    return DistArrayImpl(shape=(), dtype=np.float64, tiles=extents, reducer_fn=None)
  shape = extent.find_shape(list(extents.keys()))
  dtype = getattr(extents, 'dtype', None)
  if dtype is None:
    for tid in extents.values():
      if ctx.is_local(tid):
        dtype = ctx.tile(tid).dtype
        break
  if dtype is None:
    raise SpartanError('from_table: dtype unknown on a rank without local tiles; pass a TileTable with dtype')
  return DistArrayImpl(shape=shape, dtype=dtype, tiles=extents, reducer_fn=None)


class TileTable(collections.OrderedDict):
  """extent -> TileId table that also carries the dtype, so ranks that own none of the tiles agree on it."""
  dtype = None


class LocalWrapper(DistArray):
  """DistArray interface over host data every rank holds (scalars, NumPy operands);
  distarray.py:553-602.  ``device_data`` uploads it once per rank."""

  def __init__(self, data):
    self._data = np.asarray(data)
    self.sparse = False
    self.bad_tiles = []
    self._ex = extent.from_slice(np.index_exp[:], self.shape) if self._data.ndim else extent.create((), (), ())
    self._dev = None

  @property
  def dtype(self):
    return self._data.dtype

  @property
  def shape(self):
    return self._data.shape

  @property
  def tiles(self):
    return {self._ex: TileId(-1, 0)}

  def extent_for_blob(self, tile_id):
    return self._ex

  def host_data(self):
    return self._data

  def device_data(self):
    if self._dev is None:
      self._dev = torch.from_numpy(np.ascontiguousarray(self._data)).to(blob_ctx.get().device)
    return self._dev

  def fetch(self, ex, dst=None):
    d = self.device_data()
    return d[ex.to_slice()] if d.dim() else d

  def glom(self):
    return self._data

  def foreach_tile(self, mapper_fn, kw=None):
    raise SpartanError('mapping over a purely local value is not supported on the device path')


def as_array(data):
  """distarray.py:605-617."""
  if isinstance(data, DistArray):
    return data
  return LocalWrapper(data)


def largest_value(vals):
  """distarray.py:636-642: the input with the most elements drives the map.  Distributed arrays are preferred
  over values every rank holds, so an empty distributed array still drives (and yields an empty result)."""
  dist = [v for v in vals if isinstance(v, DistArrayImpl) or getattr(v, 'is_view', False)]
  return max(dist or vals, key=lambda v: v.real_size())


# ------------------------------------------------------------------------------------ broadcast.py
class Broadcast(DistArray):
  """NumPy broadcasting as a view (spartan/expr/operator/broadcast.py:28-109)."""

  def __init__(self, base, shape):
    require_type(base, DistArray)
    self.base = base.base if isinstance(base, Broadcast) else base
    self.shape = tuple(shape)
    self.tiles = self.base.tiles
    self.dtype = base.dtype
    self.sparse = False
    self.bad_tiles = []
    self.prepend_dim = len(shape) - len(self.base.shape)

  def real_size(self):
    return int(np.prod(self.base.shape, dtype=np.int64)) - 1       # broadcast.py:56-61

  def extent_for_blob(self, tile_id):
    return self.base.extent_for_blob(tile_id)

  def _base_ex(self, ex):
    """broadcast.py:74-92."""
    while len(ex.shape) > len(self.base.shape):
      ex = extent.drop_axis(ex, 0)
    if len(self.base.shape) == 0:
      return extent.create((), (), ())
    ul, lr = [], []
    for i in range(len(self.base.shape)):
      if self.base.shape[i] == 1:
        ul.append(0); lr.append(1)
      else:
        ul.append(ex.ul[i]); lr.append(ex.lr[i])
    return extent.create(ul, lr, self.base.shape)

  def fetch_base_tile(self, ex, dst=None):
    """The un-expanded base region that broadcasts to ``ex`` (broadcast.py:106-109); the kernels
    broadcast through zero strides instead of materialising."""
    return self.base.fetch(self._base_ex(ex), dst=dst)

  def fetch(self, ex, dst=None):
    t = self.fetch_base_tile(ex, dst=dst)
    return None if t is None else t.expand(ex.shape)

  def foreach_tile(self, mapper_fn, kw=None):
    raise SpartanError('a broadcast operand cannot be the largest input of a map')


def broadcast(args):
  """broadcast.py:111-158."""
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
    # NumPy rule: sizes are equal or 1; the result takes the size that is not 1 (which may be 0)
    size = next((v for v in (shp[axis] for shp in new_shapes) if v != 1), 1)
    for shp in new_shapes:
      shp[axis] = size
  results = []
  for i in range(len(args)):
    if new_shapes[i] == orig_shapes[i]:
      results.append(args[i])
    else:
      results.append(Broadcast(args[i], tuple(new_shapes[i])))
  return results
