"""DeviceTile: one tile of a distributed array, resident in the HBM of its owning GPU.

Counterpart of spartan/array/tile.pyx (Tile, from_data, from_shape, merge).  Only the dense mode
exists on the device; instead of a per-element mask the tile tracks ``valid`` (nothing written yet /
written), and a first partial write under a reducer initialises the rest of the tile with the
reducer's identity -- observably the same as the reference's "first write replaces, later writes
reduce" rule (tile.pyx:250-283) for every combiner it uses (add, multiply, minimum, maximum,
logical_and, logical_or).
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
    self.valid = bool(valid)
    self.type = TYPE_DENSE

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


def merge(old_tile, subslice, update, reducer):
  """tile.pyx:200-297, dense and 0-d paths, on the device."""
  red = reducer_op(reducer)
  data = old_tile._alloc()
  full = subslice is None or data.dim() == 0 or tuple(update.shape) == tuple(data.shape)
  region = data if full else data[subslice]
  if not old_tile.valid:
    if not full:
      data.fill_(identity_of(red, old_tile.dtype) if red is not None else 0)
    device_ops.copy_into(region, update)          # first write replaces (and casts to the tile dtype, :267)
    old_tile.valid = True
  elif red is None:
    device_ops.copy_into(region, update)
  else:
    device_ops.combine_into(region, update, red)
  return old_tile
