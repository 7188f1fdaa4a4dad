"""DeviceTile: one tile of a distributed array, resident in the HBM of its owning GPU.

Counterpart of spartan/array/tile.pyx (Tile, from_data, from_shape, merge).  Only the dense mode
exists on the device; instead of a per-element mask the tile tracks ``valid`` (nothing written yet /
written), and a first partial write under a reducer initialises the rest of the tile with the
reducer's identity -- observably the same as the reference's "first write replaces, later writes
reduce" rule (tile.pyx:250-283) for every combiner it uses (add, multiply, minimum, maximum,
logical_and, logical_or) on every element that has been written.  The full-tile fast path keeps the
reference's test (tile.pyx:263-268): a full-tile update reduces only if the tile's FIRST element has been
written before, otherwise it replaces the whole tile -- ``origin_written`` tracks that one mask bit.
"""
import numpy as np

from .. import device_ops
from .._lib import SpartanError, SP_RED_SUM, SP_RED_MIN, SP_RED_MAX, SP_RED_PROD, SP_RED_ALL, SP_RED_ANY

TYPE_EMPTY, TYPE_DENSE, TYPE_MASKED, TYPE_SPARSE = 0, 1, 2, 3

_REDUCERS = {np.add: SP_RED_SUM, np.minimum: SP_RED_MIN, np.maximum: SP_RED_MAX, np.multiply: SP_RED_PROD,
             np.logical_and: SP_RED_ALL, np.logical_or: SP_RED_ANY}


def reducer_op(reducer):
  """Maps the NumPy combiner functions the reference passes as ``reducer`` / ``accumulate_fn`` to a
  device reduce op.  Arbitrary Python callables cannot run on the GPU."""
  if reducer is None:
    return None
  op = _REDUCERS.get(reducer)
  if op is None:
    op = getattr(reducer, 'device_reduce', None)
  if op is None:
    raise SpartanError('reducer %r is not GPU-mappable (supported: np.add, np.multiply, np.minimum, np.maximum, '
                       'np.logical_and, np.logical_or)' % (reducer,))
  return op


def identity_of(red_op, dtype):
  if red_op == SP_RED_SUM or red_op == SP_RED_ANY: return 0
  if red_op == SP_RED_PROD or red_op == SP_RED_ALL: return 1
  dtype = np.dtype(dtype)
  if dtype.kind == 'f':
    return float('inf') if red_op == SP_RED_MIN else float('-inf')
  if dtype.kind == 'b':
    return 1 if red_op == SP_RED_MIN else 0
  info = np.iinfo(dtype)
  return info.max if red_op == SP_RED_MIN else info.min


class DeviceTile(object):
  def __init__(self, shape, dtype, data=None, valid=False):
    self.shape = tuple(int(s) for s in shape)
    self.dtype = np.dtype(dtype)
    self.data = data            # torch tensor on the owning device (may be a view into an array slab)
    self._valid = bool(valid)
    self.origin_written = bool(valid)     # mask[0, 0, ...] of the reference's Tile (tile.pyx:264)
    self.type = TYPE_DENSE

  @property
  def valid(self):
    return self._valid

  @valid.setter
  def valid(self, v):
    """Set by producers that write whole tiles (kernels, uploads): every element is written, the first included."""
    self._valid = bool(v)
    self.origin_written = bool(v)

  def _alloc(self):
    if self.data is None:
      from .. import blob_ctx
      self.data = blob_ctx.get().empty(self.shape, self.dtype)
    return self.data

  def get(self, subslice=None):
    """tile.pyx:68-112: an unwritten tile yields uninitialised memory, like the reference (:72-79)."""
    data = self._alloc()
    if subslice is None or data.dim() == 0:
      return data
    return data[subslice]

  def update(self, subslice, update, reducer):
    return merge(self, subslice, update, reducer)


def from_data(data):
  """tile.pyx:145-160 -- wrap an existing device tensor."""
  from ..blob_ctx import _TORCH_DTYPES
  np_dtype = [k for k, v in _TORCH_DTYPES.items() if v == data.dtype][0]
  return DeviceTile(tuple(data.shape), np_dtype, data, valid=True)


def from_shape(shape, dtype, tile_type=TYPE_DENSE):
  """tile.pyx:163-179."""
  if tile_type != TYPE_DENSE:
    raise SpartanError('only dense tiles exist on the device')
  return DeviceTile(shape, dtype, None, valid=False)


def _covers_origin(subslice):
  for sl in subslice:
    if isinstance(sl, slice):
      if (sl.start or 0) != 0:
        return False
    elif int(sl) != 0:
      return False
  return True


def merge(old_tile, subslice, update, reducer):
  """tile.pyx:200-297, dense and 0-d paths, on the device."""
  red = reducer_op(reducer)
  data = old_tile._alloc()
  full = subslice is None or data.dim() == 0 or tuple(update.shape) == tuple(data.shape)
  region = data if full else data[subslice]
  valid = old_tile.valid
  if full:
    # tile.pyx:263-268: reduce when there is a reducer and the first element was written, else replace (with the cast
    # to the tile dtype of :267); 0-d tiles: :212-217
    if red is not None and valid and (old_tile.origin_written or data.dim() == 0):
      device_ops.combine_into(region, update, red)
    else:
      device_ops.copy_into(region, update)
    old_tile.valid = True
    old_tile.origin_written = True
    return old_tile
  if not valid:
    data.fill_(identity_of(red, old_tile.dtype) if red is not None else 0)
    device_ops.copy_into(region, update)          # first write replaces (:272-273)
    old_tile._valid = True
  elif red is None:
    device_ops.copy_into(region, update)
  else:
    device_ops.combine_into(region, update, red)
  if _covers_origin(subslice):
    old_tile.origin_written = True
  return old_tile
