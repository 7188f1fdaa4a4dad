"""ORACLE (test infrastructure only) -- builders the reference's creation / statistics tests pin that are not part of
expr.py: eye, identity, diagonal, diagflat, diag, std.

Restates spartan/expr/creation.py: _eye_mapper :51-53, eye :56-60, identity :63-64, _diagflat_mapper :228-249,
diagflat :252-261, _diagonal_mapper :264-278, diagonal :281-298, diag :301-330; spartan/expr/statistics.py: std :86-102.
Pinned by tests/test_creation.py:10-16,65-92 and tests/test_statistics.py:32-70 (tests/test_oracle_reference_vectors.py).
"""
import builtins as _b

import numpy as np

from . import extent
from . import expr as _e
from .expr import map_with_location, map2, ndarray, lazify


def _eye_mapper(tile, ex, k=None, dtype=None):
  # creation.py:51-53
  return np.eye(ex[1][0] - ex[0][0], M=(ex[1][1] - ex[0][1]), k=(ex[0][0] + k), dtype=dtype)


def eye(N, M=None, k=0, dtype=np.float32, tile_hint=None):
  # creation.py:56-60
  if M is None:
    M = N
  return map_with_location(ndarray((N, M), dtype, tile_hint), _eye_mapper, fn_kw={'k': k, 'dtype': dtype})


def identity(n, dtype=np.float32, tile_hint=None):
  return eye(n, dtype=dtype, tile_hint=tile_hint)


def _diagflat_mapper(extents, tiles, shape=None):
  # creation.py:228-249
  ex = extents[0]
  tile = tiles[0]
  head = extent.ravelled_pos(ex.ul, ex.array_shape)
  tail = extent.ravelled_pos([l - 1 for l in ex.lr], ex.array_shape)
  result = np.diagflat(tile)
  if head != 0:
    result = np.hstack((np.zeros(((tail - head + 1), head)), result))
  if tail + 1 != shape[0]:
    result = np.hstack((result, np.zeros((tail - head + 1, shape[0] - (tail + 1)))))
  target_ex = extent.create((head, 0), (tail + 1, shape[1]), shape)
  yield target_ex, result


def diagflat(array):
  # creation.py:252-261
  array = lazify(array)
  shape = (int(np.prod(array.shape)), int(np.prod(array.shape)))
  return map2(array, 0, fn=_diagflat_mapper, fn_kw={'shape': shape}, shape=shape)


def _diagonal_mapper(ex, tiles, shape=None):
  # creation.py:264-278
  tile = tiles[0]
  max_dim = _b.max(*ex.ul)
  first_point = [max_dim for i in range(len(ex.ul))]
  slices = []
  for i in range(len(ex.ul)):
    if first_point[i] >= ex.lr[i]:
      return
    slices.append(slice(first_point[i] - ex.ul[i], ex.shape[i]))
  result = tile[tuple(slices)].diagonal()
  target_ex = extent.create((first_point[0],), (first_point[0] + result.shape[0],), shape)
  yield target_ex, result


def diagonal(a):
  # creation.py:281-298
  a = lazify(a)
  if len(a.shape) < 2:
    raise ValueError('diag requires an array of at least two dimensions')
  shape = (_b.min(a.shape),)
  return map2(a, fn=_diagonal_mapper, fn_kw={'shape': shape}, shape=shape)


def diag(array, offset=0):
  # creation.py:301-330
  if offset != 0:
    raise NotImplementedError
  array = lazify(array)
  if len(array.shape) == 1:
    return diagflat(array)
  elif len(array.shape) == 2:
    return diagonal(array)
  raise ValueError('Input must be 1- or 2-d.')


def std(a, axis=None):
  # statistics.py:86-102
  a_casted = _e.astype(a, np.float64)
  return _e.sqrt(_e.mean(a_casted ** 2, axis) - _e.mean(a_casted, axis) ** 2)


for _name in ('eye', 'identity', 'diagflat', 'diagonal', 'diag', 'std'):
  setattr(_e, _name, globals()[_name])
