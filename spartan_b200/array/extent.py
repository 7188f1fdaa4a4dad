"""TileExtent and extent algebra for the B200 host.

Same names and semantics as the reference's Cython module (spartan/array/extent.pyx); the integer
algebra itself runs in the native shim (spartan_b200/csrc/extent.cu) through the C ABI, so the
Python layer only boxes/unboxes tuples.  An extent is [ul, lr) inside an array of ``array_shape``.
"""
import ctypes

import numpy as np

from .._lib import lib, check, i64arr, SP_AXIS_NONE
from ..util import divup  # noqa: F401  (re-exported like the reference's util.divup users expect)

MAX_DIM = 32   # extent.pyx:20-21


class TileExtent(object):
  """extent.pyx:23-136."""
  __slots__ = ('ul', 'lr', 'array_shape', '_hash')

  def __init__(self, ul, lr, array_shape):
    self.ul = tuple(int(x) for x in ul)
    self.lr = tuple(int(x) for x in lr)
    self.array_shape = None if array_shape is None else tuple(int(x) for x in array_shape)
    self._hash = hash(self.ul)                             # extent.pyx:93-94; extents are immutable, hashed once

  @property
  def shape(self):
    # zero-length dims report 1 (extent.pyx:66-72)
    return tuple((l - u) if (l - u) != 0 else 1 for u, l in zip(self.ul, self.lr))

  @property
  def size(self):
    return int(np.prod(self.shape, dtype=np.int64))

  @property
  def ndim(self):
    return len(self.ul)

  def to_slice(self):
    return tuple(slice(u, l) for u, l in zip(self.ul, self.lr))

  def to_tuple(self):
    return (self.ul, self.lr, self.array_shape)

  def __reduce__(self):
    return create, (self.ul, self.lr, self.array_shape)

  def __repr__(self):
    return 'extent(' + ','.join('%s:%s' % (a, b) for a, b in zip(self.ul, self.lr)) + ')'

  def __getitem__(self, idx):
    return create((self.ul[idx],), (self.lr[idx],), (self.array_shape[idx],))

  def __hash__(self):
    return self._hash

  def __eq__(self, other):                                 # extent.pyx:107-110
    return isinstance(other, TileExtent) and self.ul == other.ul and self.lr == other.lr

  def __ne__(self, other):
    return not self.__eq__(other)

  def __lt__(self, other):                                 # extent.pyx:96-106
    for a, b in zip(self.ul, other.ul):
      if a < b: return True
      if a > b: return False
    return True

  def __gt__(self, other):
    return not self.__lt__(other)

  def ravelled_pos(self):
    return ravelled_pos(self.ul, self.array_shape)

  def to_global(self, idx, axis):
    """extent.pyx:121-127."""
    n = self.ndim
    return int(lib.sp_extent_to_global(n, i64arr(self.ul), i64arr(self.lr), i64arr(self.array_shape), int(idx),
                                       SP_AXIS_NONE if axis is None else int(axis)))

  def add_dim(self):
    return create(self.ul + (0,), self.lr + (1,), self.array_shape + (1,))

  def clone(self):
    return create(self.ul, self.lr, self.array_shape)


def create(ul, lr, array_shape):
  """extent.pyx:161-182: None when any ul >= lr; the 0-d extent create((), (), ()) is valid."""
  ul = tuple(ul); lr = tuple(lr)
  if len(ul) > MAX_DIM:
    raise ValueError('extents support at most %d dimensions' % MAX_DIM)
  for u, l in zip(ul, lr):
    if u >= l:
      return None
  return TileExtent(ul, lr, array_shape)


def from_shape(shp):
  return create([0] * len(shp), list(shp), tuple(shp))


def from_tuple(tup):
  return create(tup[0], tup[1], tup[2])


def unravelled_pos(idx, array_shape):
  n = len(array_shape)
  out = (ctypes.c_int64 * max(1, n))()
  check(lib.sp_extent_unravelled_pos(int(idx), n, i64arr(array_shape), out), 'unravelled_pos')
  return tuple(out[i] for i in range(n))


def ravelled_pos(idx, array_shape):
  return int(lib.sp_extent_ravelled_pos(len(array_shape), i64arr(idx), i64arr(array_shape)))


def all_nonzero_shape(shape):
  return all(i != 0 for i in shape)


def intersection(a, b):
  """extent.pyx:367-387."""
  if a is None:
    return None
  assert a.array_shape == b.array_shape, 'Tiles must have compatible shapes!'
  n = a.ndim
  oul = (ctypes.c_int64 * max(1, n))(); olr = (ctypes.c_int64 * max(1, n))()
  rc = check(lib.sp_extent_intersection(n, i64arr(a.ul), i64arr(a.lr), i64arr(b.ul), i64arr(b.lr), oul, olr),
             'intersection')
  if rc == 0:
    return None
  return TileExtent([oul[i] for i in range(n)], [olr[i] for i in range(n)], a.array_shape)


def find_overlapping(extents, region):
  for ex in extents:
    overlap = intersection(ex, region)
    if overlap is not None:
      yield (ex, overlap)


def compute_slice(base, idx):
  """extent.pyx:266-296."""
  if np.isscalar(idx):
    idx = slice(idx, idx + 1)
  if not isinstance(idx, tuple):
    idx = (idx,)
  ul, lr = [], []
  for i in range(base.ndim):
    if i >= len(idx):
      ul.append(base.ul[i]); lr.append(base.lr[i])
    else:
      axis_idx = idx[i]
      if np.isscalar(axis_idx):
        axis_idx = slice(axis_idx, axis_idx + 1)
      start, stop, _ = axis_idx.indices(base.shape[i])
      ul.append(base.ul[i] + start)
      lr.append(base.ul[i] + stop)
  return create(ul, lr, base.array_shape)


def offset_from(base, other):
  """extent.pyx:298-314."""
  ul, lr = [], []
  for i in range(base.ndim):
    assert not (other.ul[i] < base.ul[i] or other.lr[i] > base.lr[i])
    ul.append(other.ul[i] - base.ul[i])
    lr.append(other.lr[i] - base.ul[i])
  return create(ul, lr, other.array_shape)


def offset_slice(base, other):
  """extent.pyx:316-324."""
  return tuple(slice(other.ul[i] - base.ul[i], other.lr[i] - base.ul[i], None) for i in range(base.ndim))


def from_slice(idx, shape):
  """extent.pyx:326-363."""
  if not isinstance(idx, tuple):
    idx = (idx,)
  if len(idx) < len(shape):
    idx = tuple(list(idx) + [slice(None, None, None)] * (len(shape) - len(idx)))
  ul, lr = [], []
  for i in range(len(shape)):
    slc = idx[i]
    if np.isscalar(slc):
      slc = int(slc)
      slc = slice(slc, slc + 1, None)
    start, stop, _ = slc.indices(shape[i])
    ul.append(start); lr.append(stop)
  return create(ul, lr, shape)


def shape_for_reduction(input_shape, axis):
  """extent.pyx:390-400: () for axis=None, a list otherwise."""
  if axis is None:
    return ()
  input_shape = list(input_shape)
  del input_shape[axis]
  return input_shape


def shapes_match(offset, data):
  return tuple(offset.shape) == tuple(data.shape)


def drop_axis(ex, axis):
  """extent.pyx:411-432."""
  n = ex.ndim
  if axis is None:
    return create((), (), ())
  oul = (ctypes.c_int64 * max(1, n))(); olr = (ctypes.c_int64 * max(1, n))(); osh = (ctypes.c_int64 * max(1, n))()
  ond = ctypes.c_int(0)
  rc = check(lib.sp_extent_drop_axis(n, i64arr(ex.ul), i64arr(ex.lr), i64arr(ex.array_shape), int(axis), oul, olr, osh,
                                     ctypes.byref(ond)), 'drop_axis')
  if rc == 0:
    return None
  m = ond.value
  return TileExtent([oul[i] for i in range(m)], [olr[i] for i in range(m)], [osh[i] for i in range(m)])


index_for_reduction = drop_axis


def find_shape(extents):
  """extent.pyx:434-443."""
  shape = np.max([ex.lr for ex in extents], axis=0)
  shape[shape == 0] = 1
  return tuple(int(s) for s in shape)


def is_complete(shape, slices):
  if len(shape) != len(slices):
    return False
  for dim, slc in zip(shape, slices):
    if slc.start > 0: return False
    if slc.stop < dim: return False
  return True


def partition_axes(ex):
  return [i for i in range(len(ex.shape)) if ex.shape[i] != ex.array_shape[i]]


def change_partition_axis(ex, axis):
  """extent.pyx:501-570, one-dimensional target.  Grid-tiled extents raise SpartanError: the
  reference's handling of them (extent.pyx:545-552) is a defect and callers route around it."""
  if isinstance(axis, (list, tuple)):
    raise NotImplementedError('grid re-partition is out of scope')
  n = ex.ndim
  oul = (ctypes.c_int64 * n)(); olr = (ctypes.c_int64 * n)()
  rc = check(lib.sp_extent_change_partition_axis(n, i64arr(ex.ul), i64arr(ex.lr), i64arr(ex.array_shape), int(axis),
                                                 oul, olr), 'change_partition_axis')
  if rc == 0:
    return None
  return TileExtent([oul[i] for i in range(n)], [olr[i] for i in range(n)], ex.array_shape)
