"""ORACLE (test infrastructure only) -- Slice / Transpose / Reshape views and their Expr nodes.

Restates (NumPy on the host, one function per reference function):
  spartan/expr/operator/slice.py      _slice_mapper :9-39, Slice :42-85, SliceExpr :88-137
  spartan/expr/operator/transpose.py  _tile_mapper :19-24, Transpose :27-67, TransposeExpr :70-84, transpose :86-100
  spartan/expr/operator/reshape.py    _ravelled_ex :20-23, _unravelled_ex :26-29, _tile_mapper :32-44,
                                      Reshape :47-193 (dense fetch), ReshapeExpr :196-209, reshape :212-239
  spartan/array/extent.pyx            find_rect :234-252
  spartan/expr/operator/base.py       Expr.__getitem__ :401-448
  spartan/expr/manipulation.py        ravel :13-22
Pinned by the assertions of the reference's tests/test_slice.py, tests/test_transpose.py and
tests/test_reshape.py re-run against this module (tests/test_oracle_reference_vectors.py).
"""
import itertools

import numpy as np

from . import distarray, extent
from .distarray import DistArray, _tile_order
from . import expr as _e
from .expr import Expr, NotShapeable, lazify


# ----------------------------------------------------------------------------------- slice.py
def _slice_mapper(ex, **kw):
  # slice.py:9-39
  mapper_fn = kw['_slice_fn']
  slice_extent = kw['_slice_extent']
  fn_kw = kw['fn_kw']
  if fn_kw is None:
    fn_kw = {}
  intersection = extent.intersection(slice_extent, ex)
  if intersection is None:
    return []
  offset = extent.offset_from(slice_extent, intersection)
  offset = extent.create(offset.ul, offset.lr, slice_extent.shape)     # offset.array_shape = slice_extent.shape
  return mapper_fn(offset, **fn_kw)


class Slice(DistArray):
  # slice.py:42-85
  def __init__(self, darray, idx):
    if not isinstance(idx, extent.TileExtent):
      idx = extent.from_slice(idx, darray.shape)
    assert isinstance(darray, DistArray)
    self.base = darray
    self.slice = idx
    self.shape = self.slice.shape
    self.tiles = self.base.tiles
    self.dtype = darray.dtype
    self.sparse = self.base.sparse
    self._tile_shape = distarray.good_tile_shape(self.shape, distarray.get_ctx().num_workers)

  def tile_shape(self):
    return self._tile_shape

  def foreach_tile(self, mapper_fn, kw):
    return self.base.foreach_tile(mapper_fn=_slice_mapper,
                                  kw={'fn_kw': kw, '_slice_extent': self.slice, '_slice_fn': mapper_fn})

  def fetch(self, idx):
    offset = extent.compute_slice(self.slice, idx.to_slice())
    return self.base.fetch(offset)


class SliceExpr(Expr):
  # slice.py:88-137
  members = ('src', 'idx', 'broadcast_to')

  def dependencies(self):
    return {'src': self.src, 'idx': self.idx}

  def visit(self, visitor):
    return SliceExpr(src=visitor.visit(self.src), idx=self.idx, broadcast_to=self.broadcast_to, expr_id=self.expr_id,
                     shape_cache=self.shape_cache)

  def compute_shape(self):
    if isinstance(self.idx, (int, slice, tuple)):
      src_shape = self.src.shape
      ex = extent.from_shape(src_shape)
      return extent.compute_slice(ex, self.idx).shape
    raise NotShapeable

  def _evaluate(self, ctx, deps):
    src, idx = deps['src'], deps['idx']
    if self.broadcast_to is not None and src.shape != self.broadcast_to:
      src = distarray.Broadcast(src, self.broadcast_to)
    return Slice(src, idx)


# ----------------------------------------------------------------------------------- transpose.py
def _transpose_tile_mapper(ex, **kw):
  # transpose.py:19-24
  user_fn, fn_kw, base = kw['_fn'], kw['_fn_kw'], kw['_base']
  base_ex = extent.create(ex.ul[::-1], ex.lr[::-1], base.shape)
  return user_fn(base_ex, **(fn_kw or {}))


class Transpose(DistArray):
  # transpose.py:27-67
  def __init__(self, base):
    assert isinstance(base, DistArray)
    self.base = base
    self.shape = self.base.shape[::-1]
    self.dtype = base.dtype
    self.sparse = self.base.sparse
    self.tiles = base.tiles

  def tile_shape(self):
    return self.base.tile_shape()[::-1]

  def foreach_tile(self, mapper_fn, kw=None):
    return self.base.foreach_tile(mapper_fn=_transpose_tile_mapper, kw={'_fn_kw': kw, '_base': self, '_fn': mapper_fn})

  def fetch(self, ex):
    base_ex = extent.create(ex.ul[::-1], ex.lr[::-1], self.base.shape)
    return self.base.fetch(base_ex).transpose()


class TransposeExpr(Expr):
  # transpose.py:70-84
  members = ('array', 'tile_hint')

  def dependencies(self):
    return {'array': self.array}

  def visit(self, visitor):
    return TransposeExpr(array=visitor.visit(self.array), tile_hint=self.tile_hint, expr_id=self.expr_id,
                         shape_cache=self.shape_cache)

  def _evaluate(self, ctx, deps):
    return Transpose(deps['array'])

  def compute_shape(self):
    return self.array.shape[::-1]


def transpose(array, tile_hint=None):
  return TransposeExpr(array=lazify(array), tile_hint=tile_hint)


# ----------------------------------------------------------------------------------- reshape.py
def find_rect(ravelled_ul, ravelled_lr, shape):
  # extent.pyx:234-252 (Python-2 integer division)
  if shape[-1] == 1 or ravelled_ul // shape[-1] == ravelled_lr // shape[-1]:
    return ravelled_ul, ravelled_lr
  div = 1
  for i in shape[1:]:
    div = div * i
  return ravelled_ul - (ravelled_ul % div), ravelled_lr + (div - ravelled_lr % div) % div - 1


def _ravelled_ex(ul, lr, shape):
  return extent.ravelled_pos(ul, shape), extent.ravelled_pos([l - 1 for l in lr], shape)


def _unravelled_ex(ravelled_ul, ravelled_lr, shape):
  return extent.unravelled_pos(ravelled_ul, shape), extent.unravelled_pos(ravelled_lr, shape)


class Reshape(DistArray):
  # reshape.py:47-193, dense arrays
  def __init__(self, base, shape, tile_hint=None):
    assert isinstance(base, DistArray)
    self.base = base
    self.shape = tuple(shape)
    self.dtype = base.dtype
    self.sparse = self.base.sparse
    self.tiles = self.base.tiles
    self._tile_shape = distarray.good_tile_shape(self.shape, distarray.get_ctx().num_workers)
    self.shape_array = None
    self.is_add_dimension = False
    shape = self.shape
    if len(shape) == len(self.base.shape) + 1:                   # :69-88
      self.is_add_dimension = True
      extra = 0
      for i in range(len(self.base.shape)):
        if shape[i + extra] != self.base.shape[i]:
          if extra == 0 and shape[i] == 1:
            self.new_dimension_idx = i
            extra = 1
          else:
            self.is_add_dimension = False
            break
      if extra == 0:
        self.new_dimension_idx = len(shape) - 1
    self._check_extents()

  def _check_extents(self):
    # reshape.py:92-119.  `if rect_ul or ul or ...` tests the truthiness of the tuple `ul`, so any non-0-d
    # split clears _same_tiles; only the appended-dimension case keeps the base tiles.
    self._same_tiles = True
    if len(self.shape) > len(self.base.shape):
      for i in range(len(self.base.shape)):
        if self.base.shape[i] != self.shape[i]:
          self._same_tiles = False
          break
      if self._same_tiles:
        return
    splits = distarray.compute_splits(self.shape, self._tile_shape)
    for slc in itertools.product(*splits):
      ul, lr = zip(*slc)
      ravelled_ul, ravelled_lr = _ravelled_ex(ul, lr, self.shape)
      rect_ul, rect_lr = find_rect(ravelled_ul, ravelled_lr, self.base.shape)
      if rect_ul or ul or rect_lr != lr:
        self._same_tiles = False
        break

  def tile_shape(self):
    return self._tile_shape

  def _view_extent_of_base(self, ex):
    ravelled_ul, ravelled_lr = _ravelled_ex(ex.ul, ex.lr, self.base.shape)
    ul, lr = _unravelled_ex(ravelled_ul, ravelled_lr, self.shape)
    return extent.create(ul, tuple(np.array(lr) + 1), self.shape)

  def foreach_tile(self, mapper_fn, kw=None):
    # reshape.py:131-149 + _tile_mapper :32-44
    kw = dict(kw or {})
    if self._same_tiles:
      return [mapper_fn(self._view_extent_of_base(ex), **kw) for ex, _ in _tile_order(self.base.tiles)]
    if self.shape_array is None:
      self.shape_array = distarray.create(self.shape, self.base.dtype, tile_hint=self._tile_shape)
    return [mapper_fn(ex, **kw) for ex, _ in _tile_order(self.shape_array.tiles)]

  def fetch(self, ex):
    # reshape.py:159-179 (the "complete rows" limitation of :166-168 is the reference's)
    if self.is_add_dimension:
      ul = ex.ul[0:self.new_dimension_idx] + ex.ul[self.new_dimension_idx + 1:]
      lr = ex.lr[0:self.new_dimension_idx] + ex.lr[self.new_dimension_idx + 1:]
      base_ex = extent.create(ul, lr, self.base.shape)
      return self.base.fetch(base_ex).reshape(ex.shape)
    ravelled_ul, ravelled_lr = _ravelled_ex(ex.ul, ex.lr, self.shape)
    base_ravelled_ul, base_ravelled_lr = find_rect(ravelled_ul, ravelled_lr, self.base.shape)
    base_ul, base_lr = _unravelled_ex(base_ravelled_ul, base_ravelled_lr, self.base.shape)
    base_ex = extent.create(base_ul, tuple(np.array(base_lr) + 1), self.base.shape)
    tile = np.ravel(self.base.fetch(base_ex))
    tile = tile[(ravelled_ul - base_ravelled_ul):(ravelled_lr - base_ravelled_ul) + 1]
    assert np.prod(tile.shape) == np.prod(ex.shape), (tile.shape, ex.shape)
    return tile.reshape(ex.shape)


class ReshapeExpr(Expr):
  # reshape.py:196-209
  members = ('array', 'new_shape', 'tile_hint')

  def dependencies(self):
    return {'array': self.array}

  def visit(self, visitor):
    return ReshapeExpr(array=visitor.visit(self.array), new_shape=self.new_shape, tile_hint=self.tile_hint,
                       expr_id=self.expr_id, shape_cache=self.shape_cache)

  def _evaluate(self, ctx, deps):
    return Reshape(deps['array'], self.new_shape, self.tile_hint)

  def compute_shape(self):
    return tuple(self.new_shape)


def reshape(array, *args, **kargs):
  # reshape.py:212-239
  if len(args) == 1 and isinstance(args[0], (tuple, list)):
    new_shape = tuple(args[0])
  else:
    new_shape = tuple(args)
  return ReshapeExpr(array=lazify(array), new_shape=new_shape, tile_hint=kargs.get('tile_hint'))


def ravel(v):
  # manipulation.py:13-22
  return reshape(v, (int(np.prod(v.shape)),))


# ----------------------------------------------------------------------------------- base.py:401-448
def _getitem(self, idx):
  newaxis = _e.newaxis
  if isinstance(idx, (int, tuple, slice)):
    is_del_dim = False
    del_dim = list()
    if isinstance(idx, tuple):
      for x in range(len(idx)):
        if isinstance(idx[x], int):
          is_del_dim = True
          del_dim.append(x)
    if isinstance(idx, int) or is_del_dim or (isinstance(idx, tuple) and (newaxis in idx)):
      if isinstance(idx, tuple):
        new_shape = tuple([slice(x, None, None) if (isinstance(x, int) and x == -1) else x
                           for x in idx if not x == newaxis])
      else:
        new_shape = idx
      ret = SliceExpr(src=self, idx=new_shape)
      new_shape = []
      if isinstance(idx, tuple):
        shape_ptr = idx_ptr = 0
        while shape_ptr < len(ret.shape) or idx_ptr < len(idx):
          if idx_ptr < len(idx) and idx[idx_ptr] == newaxis:
            new_shape.append(1)
          else:
            new_shape.append(ret.shape[shape_ptr])
            shape_ptr += 1
          idx_ptr += 1
      else:
        new_shape = list(ret.shape)
        del_dim.append(0)
      if is_del_dim:
        for i in del_dim:
          new_shape.pop(i)
      return ReshapeExpr(array=ret, new_shape=tuple(new_shape))
    return SliceExpr(src=self, idx=idx)
  raise NotImplementedError('FilterExpr (filter.py) is outside the hot path')


Expr.__getitem__ = _getitem
Expr.reshape = reshape
Expr.ravel = ravel
Expr.transpose = transpose
Expr.T = property(transpose)
_e.transpose = transpose
_e.reshape = reshape
_e.ravel = ravel
_e.SliceExpr = SliceExpr
_e.TransposeExpr = TransposeExpr
_e.ReshapeExpr = ReshapeExpr
