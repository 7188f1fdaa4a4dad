"""Reductions: sum / min / max / prod / all / any / counts (reference:
spartan/expr/operator/reduce.py:21-167).

Per tile the reference calls the local reducer (``data.sum(axis)``...) and ships the partial to the
owner of the output tile, where ``Tile.merge`` folds it in with ``accumulate_fn`` in RPC-arrival
order (reduce.py:59-69, tile.pyx:263-283).  Here each rank folds the partials of its own tiles on the
GPU -- one fused map+reduce launch per tile, the mapped value never touches HBM -- into a
rank-local accumulator of the output's shape, and the cross-rank combine is one NCCL all-reduce with
the matching op.  The order is fixed, so results are reproducible.
"""
import collections

import numpy as np

from .. import blob_ctx, comm, device_ops
from ..array import distarray, extent, tile
from ..array.distarray import broadcast
from ..core import LocalKernelResult
from .._lib import SP_F64, SP_I64
from . import program
from .base import Expr, ListExpr, as_array
from .local import make_var, LocalReduceExpr, LocalInput
from .map import bind_operands, get_local_values, blockwise_ok, slabwise_operands


class _DtypeOf(object):
  """What ``dtype_fn`` sees: the reference passes children[0] (reduce.py:110), which after fusion is
  not the reduced value; the fused expression's own dtype is used instead."""

  def __init__(self, dtype):
    self.dtype = np.dtype(dtype)


def _value_tree(op):
  """The mapped value of a (possibly fused) reduce op: deps = [extent, value] (reduce.py:156-160,
  optimize.py:205-219)."""
  vals = [d for d in op.deps if not (isinstance(d, LocalInput) and d.idx == 'extent')]
  if len(vals) != 1:
    raise program.NotDeviceMappable('reduce over %d value operands' % len(vals))
  return vals[0]


def _reduce_mapper(ex, children, child_to_var, op, axis, output, compiled=None, red_op=None, acc=None, owner=None):
  """Local reduction of one tile (or one contiguous block of tiles) into the rank's accumulator
  (reduce.py:21-70)."""
  ctx = blob_ctx.get()
  largest = children[0]
  if owner is None:
    owner = largest.tiles[ex].worker
  values = get_local_values(ex, children, child_to_var, compiled.used_vars, owner)
  if owner == ctx.worker_id:
    dst_extent = extent.index_for_reduction(ex, axis)
    inputs = [values[v] for v in compiled.used_vars]
    target = acc[dst_extent.to_slice()] if acc.dim() else acc
    device_ops.run_map_reduce(compiled.program, inputs, ex.shape, axis, red_op, target, accumulate=True)
  return LocalKernelResult(result=[])


class ReduceExpr(Expr):
  members = ('children', 'child_to_var', 'axis', 'dtype_fn', 'op', 'accumulate_fn', 'tile_hint')

  def __init__(self, *args, **kw):
    super(ReduceExpr, self).__init__(*args, **kw)
    assert self.dtype_fn is not None
    assert isinstance(self.children, ListExpr)

  def compute_shape(self):
    # reduce.py:87-94
    shapes = [i.shape for i in self.children]
    child_shape = collections.defaultdict(int)
    for s in shapes:
      for i, v in enumerate(s):
        child_shape[i] = max(child_shape[i], v)
    input_shape = tuple(child_shape[i] for i in range(len(child_shape)))
    return tuple(extent.shape_for_reduction(input_shape, self.axis))

  def pretty_str(self):
    return 'Reduce(%s, axis=%s, %s, hint=%s)' % (getattr(self.op.fn, '__name__', '?'), self.axis, self.children,
                                                 self.tile_hint)

  def _evaluate(self, ctx, deps):
    # reduce.py:100-127
    children = list(deps['children'])
    child_to_var = list(deps['child_to_var'])
    axis = deps['axis']
    op = deps['op']
    tile_accum = deps['accumulate_fn']

    children = broadcast(children)
    largest = distarray.largest_value(children)
    i = children.index(largest)
    children[0], children[i] = children[i], children[0]
    child_to_var[0], child_to_var[i] = child_to_var[i], child_to_var[0]
    if not (isinstance(largest, distarray.DistArrayImpl) or getattr(largest, 'is_view', False)):
      raise program.NotDeviceMappable('a reduction needs a distributed (non-broadcast) input')

    spec = getattr(op.fn, 'device_reduce', None)
    if spec is None:
      raise program.NotDeviceMappable('local reducer %r is not GPU-mappable' % (op.fn,))
    red_op, post_ops, force_wide = spec
    value = _value_tree(op)
    operands = bind_operands(children, child_to_var)
    compiled = program.compile_tree(value, operands, post_ops=post_ops)
    dtype = np.dtype(deps['dtype_fn'](_DtypeOf(compiled.out_dtype)))
    if force_wide:
      # counts: 0/1 values must add exactly; float32 would saturate at 2^24
      if compiled.compute_dtype not in (SP_F64, SP_I64):
        compiled = program.compile_tree(value, operands, force_compute=SP_F64, post_ops=post_ops)
    combiner = tile.reducer_op(tile_accum)

    shape = tuple(extent.shape_for_reduction(largest.shape, axis))
    acc = ctx.empty(shape, dtype)
    acc.fill_(tile.identity_of(red_op, dtype))     # exact for int64 extremes (a double immediate is not)
    nd = len(largest.shape)
    free = None if axis is None else [axis + nd if axis < 0 else axis]
    slabs = slabwise_operands(largest, children, child_to_var, compiled.used_vars, free) \
      if blockwise_ok(largest, children) else None
    if slabs is not None:
      # the rank's whole share in one fused launch: a selection of rows / columns along the reduced axis folds into
      # the same accumulator whatever global positions it stands for
      if largest.slab.numel():
        device_ops.run_map_reduce(compiled.program, slabs, tuple(largest.slab.shape), axis, red_op, acc, accumulate=True)
    elif blockwise_ok(largest, children):
      for block in largest.local_blocks():       # one fused launch per contiguous block of this rank's slab
        _reduce_mapper(block, children, child_to_var, op, axis, None, compiled=compiled, red_op=red_op, acc=acc,
                       owner=ctx.worker_id)
    else:
      largest.foreach_tile(_reduce_mapper, kw={'children': children, 'child_to_var': child_to_var, 'op': op,
                                               'axis': axis, 'output': None, 'compiled': compiled, 'red_op': red_op,
                                               'acc': acc})
    # cross-tile combiner across GPUs: one all-reduce instead of N update RPCs into the owner tile
    comm.allreduce(acc, combiner)

    output_array = distarray.create(shape, dtype, reducer=tile_accum, tile_hint=self.tile_hint)
    for ex, tid in output_array.tiles.items():
      if ctx.is_local(tid):
        t = ctx.tile(tid)
        src = acc[ex.to_slice()] if acc.dim() else acc
        if len(output_array.tiles) == 1:
          t.data = src                      # single output tile: adopt the accumulator, no copy
        else:
          device_ops.copy_rect(t.get(None), src)
        t.valid = True
    if len(output_array.tiles) == 1 and output_array.slab is not None:
      output_array.slab = acc if acc.dim() else None
    return output_array


def reduce(v, axis, dtype_fn, local_reduce_fn, accumulate_fn, fn_kw=None, tile_hint=None):
  """Reduce ``v`` over ``axis`` (reduce.py:130-167).  ``local_reduce_fn`` must be one of the library's
  local reducers (it carries the ``device_reduce`` spec); ``accumulate_fn`` one of the NumPy combiners
  np.add / np.multiply / np.minimum / np.maximum / np.logical_and / np.logical_or."""
  fn_kw = dict(fn_kw or {})
  varname = make_var()
  assert 'axis' not in fn_kw, '"axis" argument is reserved.'
  fn_kw['axis'] = axis
  reduce_op = LocalReduceExpr(fn=local_reduce_fn, deps=[LocalInput(idx='extent'), LocalInput(idx=varname)],
                              kw=fn_kw)
  return ReduceExpr(children=ListExpr(vals=[as_array(v)]), child_to_var=[varname], axis=axis, dtype_fn=dtype_fn,
                    op=reduce_op, accumulate_fn=accumulate_fn, tile_hint=tile_hint)


class ArgReduceExpr(Expr):
  """argmin / argmax (reference: spartan/expr/sorting.py:67-124).

  The reference composes ``min -> map_with_location(_arg_mapper) -> min``: the extreme value, then for
  every element its global position if it equals the extreme (else a sentinel = array size), then the
  smallest position.  Same algorithm here in two fused passes per block of the rank's slab -- pass 2 is one
  map+reduce whose program reads the element position from the bytecode's INDEX leaf -- with the two
  cross-tile combines done by ncclAllReduce(min/max) and ncclAllReduce(min)."""
  members = ('array', 'axis', 'which')

  def compute_shape(self):
    return tuple(extent.shape_for_reduction(self.array.shape, self.axis))

  def _evaluate(self, ctx, deps):
    from .._lib import SP_RED_MIN, SP_RED_MAX, SP_F64, SP_I64, SP_F32
    arr, axis, which = deps['array'], deps['axis'], deps['which']
    if not isinstance(arr, distarray.DistArrayImpl):
      raise program.NotDeviceMappable('argmin/argmax need a distributed array')
    nd = len(arr.shape)
    if axis is not None and axis < 0:
      axis += nd
    red = SP_RED_MIN if which == 'min' else SP_RED_MAX
    out_shape = tuple(extent.shape_for_reduction(arr.shape, axis))
    is_float = arr.dtype.kind == 'f'
    blocks = arr.local_blocks()

    # pass 1: the extreme value of every output cell, replicated on every rank
    compute1 = SP_F64 if arr.dtype == np.float64 else (SP_F32 if is_float else SP_I64)
    prog1 = device_ops.make_program([('IN', 0)], compute1)
    m = ctx.empty(out_shape, arr.dtype)
    m.fill_(tile.identity_of(red, arr.dtype))
    for block in blocks:
      x = arr.fetch(block)
      dst = extent.index_for_reduction(block, axis)
      device_ops.run_map_reduce(prog1, [x], block.shape, axis, red, m[dst.to_slice()] if m.dim() else m, True)
    comm.allreduce(m, red)

    # pass 2: smallest global position whose value equals the extreme; sentinel = number of elements
    big = int(np.prod(arr.shape, dtype=np.int64))
    compute2 = SP_F64 if is_float else SP_I64
    prog2 = device_ops.make_program([('IN', 0), ('IN', 1), ('EQ', 0), ('INDEX', 0), ('CONST', 0), ('SUB', 0),
                                     ('MUL', 0), ('CONST', 0), ('ADD', 0)], compute2, [big])
    pos = ctx.empty(out_shape, np.int64)
    pos.fill_(np.iinfo(np.int64).max)
    for block in blocks:
      x = arr.fetch(block)
      dst = extent.index_for_reduction(block, axis)
      mb = m[dst.to_slice()] if m.dim() else m
      if axis is not None:
        mb = mb.unsqueeze(axis)
        index = (block.ul[axis], [1 if d == axis else 0 for d in range(nd)])
      else:
        coefs = [int(np.prod(arr.shape[d + 1:], dtype=np.int64)) for d in range(nd)]
        index = (extent.ravelled_pos(block.ul, arr.shape), coefs)
      device_ops.run_map_reduce(prog2, [x, mb], block.shape, axis, SP_RED_MIN, pos[dst.to_slice()] if pos.dim() else pos,
                                True, index=index)
    comm.allreduce(pos, SP_RED_MIN)

    output_array = distarray.create(out_shape, np.int64, reducer=np.minimum)
    for ex, tid in output_array.tiles.items():
      if ctx.is_local(tid):
        t = ctx.tile(tid)
        src = pos[ex.to_slice()] if pos.dim() else pos
        if len(output_array.tiles) == 1:
          t.data = src
        else:
          device_ops.copy_rect(t.get(None), src)
        t.valid = True
    if len(output_array.tiles) == 1 and output_array.slab is not None:
      output_array.slab = pos if pos.dim() else None
    return output_array
