"""ORACLE (test infrastructure only) -- extent algebra.

Restates spartan/array/extent.pyx (reference) in plain Python 3.  All arithmetic is on Python ints
(the reference uses int64 ``coordinate_t``, extent.pyx:18); Python-2 ``/`` on ints is written ``//``.
"""
import math

import numpy as np


def divup(a, b):
  # spartan/util.py:404-408 : int(ceil(float(a) / b))
  return int(math.ceil(float(a) / b))


class TileExtent(object):
  """extent.pyx:23-136.  [ul, lr) inside an array of shape ``array_shape``."""
  __slots__ = ('ul', 'lr', 'array_shape')

  def __init__(self, ul, lr, array_shape):
    self.ul = tuple(int(x) for x in ul)
    self.lr = tuple(int(x) for x in lr)
    self.array_shape = None if array_shape is None else tuple(int(x) for x in array_shape)

  @property
  def shape(self):
    # extent.pyx:66-72 : zero-length dims report 1
    return tuple((l - u) if (l - u) != 0 else 1 for u, l in zip(self.ul, self.lr))

  @property
  def size(self):
    return int(np.prod(self.shape))

  @property
  def ndim(self):
    return len(self.ul)

  def to_slice(self):
    return tuple(slice(u, l) for u, l in zip(self.ul, self.lr))

  def to_tuple(self):
    return (self.ul, self.lr, self.array_shape)

  def __repr__(self):
    return 'extent(' + ','.join('%s:%s' % (a, b) for a, b in zip(self.ul, self.lr)) + ')'

  def __getitem__(self, idx):
    return create((self.ul[idx],), (self.lr[idx],), (self.array_shape[idx],))

  def __hash__(self):
    return hash(self.ul)          # extent.pyx:93-94

  def __eq__(self, other):        # extent.pyx:107-110
    return isinstance(other, TileExtent) and self.ul == other.ul and self.lr == other.lr

  def __ne__(self, other):
    return not self.__eq__(other)

  def __lt__(self, other):        # extent.pyx:96-106 (lexicographic on ul; equal -> "smaller" stays True)
    for a, b in zip(self.ul, other.ul):
      if a < b: return True
      if a > b: return False
    return True

  def __gt__(self, other):
    return not self.__lt__(other)

  def ravelled_pos(self):
    return ravelled_pos(self.ul, self.array_shape)

  def to_global(self, idx, axis):
    # extent.pyx:121-127
    if axis is not None:
      return idx + self.ul[axis]
    local_idx = unravelled_pos(idx, self.shape)
    return ravelled_pos(tuple(u + l for u, l in zip(self.ul, local_idx)), self.array_shape)

  def add_dim(self):
    return create(self.ul + (0,), self.lr + (1,), self.array_shape + (1,))

  def clone(self):
    return create(self.ul, self.lr, self.array_shape)


def create(ul, lr, array_shape):
  """extent.pyx:141-182 : returns None when any ul >= lr (0-d extents are valid)."""
  ul = tuple(ul); lr = tuple(lr)
  for u, l in zip(ul, lr):
    if u >= l:
      return None
  return TileExtent(ul, lr, array_shape)


def from_shape(shp):
  return create([0] * len(shp), list(shp), tuple(shp))


def unravelled_pos(idx, array_shape):
  # extent.pyx:196-205
  out = []
  for dim in reversed(array_shape):
    out.append(idx % dim)
    idx //= dim
  return tuple(reversed(out))


def ravelled_pos(idx, array_shape):
  # extent.pyx:207-219
  rpos = 0
  mul = 1
  for i in range(len(array_shape) - 1, -1, -1):
    rpos += mul * idx[i]
    mul *= array_shape[i]
  return rpos


def all_nonzero_shape(shape):
  return all(i != 0 for i in shape)


def find_overlapping(extents, region):
  # extent.pyx:254-264
  for ex in extents:
    overlap = intersection(ex, region)
    if overlap is not None:
      yield (ex, overlap)


def compute_slice(base, idx):
  # extent.pyx:266-296
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
  # extent.pyx:298-314
  ul, lr = [], []
  for i in range(base.ndim):
    assert not (other.ul[i] < base.ul[i] or other.lr[i] > base.lr[i])
    ul.append(other.ul[i] - base.ul[i])
    lr.append(other.lr[i] - base.ul[i])
  return create(ul, lr, other.array_shape)


def offset_slice(base, other):
  # extent.pyx:316-324
  return tuple(slice(other.ul[i] - base.ul[i], other.lr[i] - base.ul[i], None) for i in range(base.ndim))


def from_slice(idx, shape):
  # extent.pyx:326-363
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


def intersection(a, b):
  """extent.pyx:367-387 (note the strict ``<`` tests; emptiness is caught by create())."""
  if a is None:
    return None
  assert a.array_shape == b.array_shape, 'Tiles must have compatible shapes!'
  ul, lr = [], []
  for i in range(a.ndim):
    if b.lr[i] < a.ul[i]: return None
    if a.lr[i] < b.ul[i]: return None
    ul.append(a.ul[i] if a.ul[i] >= b.ul[i] else b.ul[i])
    lr.append(a.lr[i] if a.lr[i] < b.lr[i] else b.lr[i])
  return create(ul, lr, a.array_shape)


def shape_for_reduction(input_shape, axis):
  # extent.pyx:390-400 : () for axis=None, a *list* otherwise
  if axis is None:
    return ()
  input_shape = list(input_shape)
  del input_shape[axis]
  return input_shape


def drop_axis(ex, axis):
  # extent.pyx:411-432
  if axis is None:
    return create((), (), ())
  if axis < 0:
    axis = ex.ndim + axis
  shape = list(ex.array_shape)
  del shape[axis]
  ul = list(ex.ul[:axis]) + list(ex.ul[axis + 1:])
  lr = list(ex.lr[:axis]) + list(ex.lr[axis + 1:])
  return create(ul, lr, shape)


index_for_reduction = drop_axis


def find_shape(extents):
  # extent.pyx:434-443
  shape = np.max([ex.lr for ex in extents], axis=0)
  shape[shape == 0] = 1
  return tuple(int(s) for s in shape)


def is_complete(shape, slices):
  # extent.pyx:446-466
  if len(shape) != len(slices):
    return False
  for dim, slc in zip(shape, slices):
    if slc.start > 0: return False
    if slc.stop < dim: return False
  return True


def partition_axes(ex):
  # extent.pyx:493-499
  return [i for i in range(len(ex.shape)) if ex.shape[i] != ex.array_shape[i]]


def change_partition_axis(ex, axis):
  """extent.pyx:501-570, one-dimensional target only.

  The grid->strip branch (extent.pyx:545-552) maps a grid tile to a 1-wide strip, which makes the
  reference's dot() wrong on grid-tiled operands.  It is not restated; callers
  that reach it get an exception so no test silently depends on it.
  """
  assert not isinstance(axis, (list, tuple)), 'grid re-partition is out of scope'
  if axis < 0:
    axis += len(ex.array_shape)
  if len(ex.shape) == 1:                                  # :533-539
    if axis == 1:
      return create((0,), ex.array_shape, ex.array_shape)
    return ex
  old_axes = partition_axes(ex)
  if len(old_axes) > 1:
    raise NotImplementedError('grid-tiled change_partition_axis is a reference defect (extent.pyx:545-552)')
  if len(old_axes) == 0 or old_axes[0] == axis:           # :554-555
    return ex
  old_axis = old_axes[0]                                  # :557-570
  new_ul = list(ex.ul)
  new_lr = list(ex.lr)
  new_ul[axis] = divup(new_ul[old_axis] * ex.array_shape[axis], ex.array_shape[old_axis])
  new_ul[old_axis] = 0
  new_lr[axis] = divup(new_lr[old_axis] * ex.array_shape[axis], ex.array_shape[old_axis])
  new_lr[old_axis] = ex.array_shape[old_axis]
  return create(new_ul, new_lr, ex.array_shape)
