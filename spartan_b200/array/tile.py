"""DeviceTile: one tile of a distributed array, resident in the HBM of its owning GPU.

Counterpart of spartan/array/tile.pyx (Tile, from_data, from_shape, merge), dense mode.  The reference keeps a
per-element mask so that the first write to an element replaces and later writes reduce (tile.pyx:250-283).  The
same rule here: the mask is ``False`` (nothing written), ``True`` (every element written -- the state a kernel that
produces the whole tile leaves) or, only once a tile has been written partially, a bool tensor next to the data; a
tile that is first touched by a partial update is zero-filled like ``np.zeros`` in Tile._initialize (:115-127).  The
full-tile fast path keeps the reference's test (:263-268): it reduces -- every element, the never-written zeros
included -- if the tile's FIRST element has been written, otherwise it replaces the whole tile.
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
    self.mask = bool(valid)     # False: nothing written; True: all written; bool tensor: partially written
    self.origin_written = bool(valid)     # mask[0, 0, ...] (tile.pyx:264), tracked on the host
    self.type = TYPE_DENSE

  @property
  def valid(self):
    """Something has been written to the tile."""
    return self.mask is not False

  @valid.setter
  def valid(self, v):
    """Set by producers that write whole tiles (kernels, uploads): every element is written, the first included."""
    self.mask = bool(v)
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
  import torch
  red = reducer_op(reducer)
  data = old_tile._alloc()
  if data.dim() == 0:                                     # :212-217
    if red is not None and old_tile.valid:
      device_ops.combine_into(data, update, red)
    else:
      device_ops.copy_into(data, update)
    old_tile.valid = True
    return old_tile
  if subslice is None or tuple(update.shape) == tuple(data.shape):
    # :263-268 -- reduce when there is a reducer and the first element was written, else replace (casting to the tile
    # dtype, :267).  A partially written tile holds zeros where nothing was written; the reduce includes them.
    if red is not None and old_tile.origin_written:
      device_ops.combine_into(data, update, red)
    else:
      device_ops.copy_into(data, update)
    old_tile.valid = True
    return old_tile
  # :270-283 partial region
  if old_tile.mask is True:
    region = data[subslice]
    if red is None:
      device_ops.copy_into(region, update)
    else:
      device_ops.combine_into(region, update, red)
    return old_tile
  if old_tile.mask is False:
    data.zero_()                                          # Tile._initialize: np.zeros (:125-126)
    old_tile.mask = torch.zeros(tuple(data.shape), dtype=torch.bool, device=data.device)
  device_ops.merge_masked(data[subslice], update, old_tile.mask[subslice], red)
  if _covers_origin(subslice):
    old_tile.origin_written = True
  return old_tile
