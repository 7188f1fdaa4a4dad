"""NumPy-style builders: thin ``map`` / ``reduce`` wrappers (reference:
spartan/expr/{creation,mathematics,statistics,logic,sorting,arrays,srandom}.py).

The functions placed into the expression tree here (``_make_ones``, ``_sum_local``...) keep the
reference's names; they are never *called* on the host -- each carries the spec the device compiler
needs (a bytecode lowering for map functions, a (reduce-op, post-ops) pair for local reducers, a fill
kernel for location functions).
"""
import builtins as _py
import numpy as np

from .. import blob_ctx, device_ops
from ..array import extent
from .._lib import (SP_RED_SUM, SP_RED_MIN, SP_RED_MAX, SP_RED_PROD, SP_RED_ALL, SP_RED_ANY, SP_FILL_CONST,
                    SP_FILL_IOTA, SP_FILL_RAND, SP_FILL_RANDN, SP_I64, SpartanError)
from . import program
from .base import Expr
from .map import map, map_with_location
from .ndarray import ndarray
from .optimize import not_idempotent
from .reduce import reduce, ArgReduceExpr
from .base import as_array as _as_array


# ------------------------------------------------------------------------------------ creation.py
def _constant_fill(value_of):
  """Lowering for _make_zeros/_make_ones/_full_mapper: the input tile is never read (creation.py:67-93):
  the node becomes an immediate of the input's dtype."""
  def handler(node, operands, analyse):
    src = analyse(node.deps[0], operands)
    dtype = node.kw.get('dtype') if node.kw and node.kw.get('dtype') is not None else src.dtype
    return program._Typed('const', [], np.dtype(dtype), weak=False, value=value_of(node))
  return handler


def _make_zeros(input): raise SpartanError('host evaluation is not available')   # creation.py:67-68
def _make_ones(input): raise SpartanError('host evaluation is not available')    # creation.py:92-93
def _full_mapper(tile, fill_value=None, dtype=None): raise SpartanError('host evaluation is not available')


program.register_special(_make_zeros, _constant_fill(lambda n: 0))
program.register_special(_make_ones, _constant_fill(lambda n: 1))
program.register_special(_full_mapper, _constant_fill(lambda n: n.kw['fill_value']))


def empty(shape, dtype=np.float32, tile_hint=None):
  """creation.py:35-42 (the reference creates a *sparse* array here -- SURVEY.md section 9 Q6; dense is the intent)."""
  return ndarray(shape, dtype=dtype, tile_hint=tile_hint)


def empty_like(array, dtype=None, tile_hint=None):
  return ndarray(array.shape, dtype=array.dtype if dtype is None else dtype, tile_hint=tile_hint)


def zeros(shape, dtype=np.float32, tile_hint=None):
  """creation.py:71-81."""
  return map(ndarray(shape, dtype=dtype, tile_hint=tile_hint), fn=_make_zeros)


def ones(shape, dtype=np.float32, tile_hint=None):
  """creation.py:96-106."""
  return map(ndarray(shape, dtype=dtype, tile_hint=tile_hint), fn=_make_ones)


def _hint_of(array):
  return getattr(array, 'tile_hint', None) if isinstance(array, Expr) else array.tile_shape()


def zeros_like(array, dtype=None, tile_hint=None):
  return zeros(array.shape, dtype=array.dtype if dtype is None else dtype,
               tile_hint=_hint_of(array) if tile_hint is None else tile_hint)


def ones_like(array, dtype=None, tile_hint=None):
  return ones(array.shape, dtype=array.dtype if dtype is None else dtype,
              tile_hint=_hint_of(array) if tile_hint is None else tile_hint)


def full(shape, fill_value, dtype=np.float32, tile_hint=None):
  """creation.py:121-123 (the reference forgets to pass fill_value -- Q6; the intent is implemented)."""
  return map(ndarray(shape, dtype=dtype, tile_hint=tile_hint), fn=_full_mapper,
             fn_kw={'fill_value': fill_value, 'dtype': dtype})


def full_like(array, fill_value, dtype=None, tile_hint=None):
  return full(array.shape, fill_value, array.dtype if dtype is None else dtype,
              _hint_of(array) if tile_hint is None else tile_hint)


def _eye_mapper(tile, ex, k=None, dtype=None):
  """creation.py:51-53 (host form, never called)."""
  raise SpartanError('host evaluation is not available')


def _eye_kernel(out, ex, k=None, dtype=None):
  """Tile of np.eye: 1 where (global column - global row) == k.  The reference offsets the tile's k by its first ROW only
  (creation.py:52), which is right for the row tilings it produces by default; the column offset is honoured here too."""
  prog = device_ops.make_program([('INDEX', 0), ('CONST', 0), ('EQ', 0)], SP_I64, [0])
  device_ops.run_map(prog, [], out, index=(int(ex.ul[1]) - int(ex.ul[0]) - int(k or 0), [-1, 1]))


_eye_mapper.device_location_kernel = _eye_kernel
_eye_mapper.result_dtype = lambda in_dtype, kw: np.dtype(kw.get('dtype') or in_dtype)


def eye(N, M=None, k=0, dtype=np.float32, tile_hint=None):
  """creation.py:56-60."""
  if M is None:
    M = N
  return map_with_location(ndarray((N, M), dtype, tile_hint), _eye_mapper, fn_kw={'k': k, 'dtype': dtype})


def identity(n, dtype=np.float32, tile_hint=None):
  """creation.py:63-64."""
  return eye(n, dtype=dtype, tile_hint=tile_hint)


def _arange_mapper(tile, ex, start, stop, step, dtype=None):
  """creation.py:135-141 (host form, never called)."""
  raise SpartanError('host evaluation is not available')


def _fill_by_global_index(out, ex, kind, a=0.0, b=0.0, seed=0):
  """Fills tile ``out`` (extent ``ex``) so that element = f(global C-order index): one launch when the tile
  spans the full trailing dimensions, one 2-D launch per leading index otherwise."""
  shape = ex.array_shape
  nd = len(shape)
  # (`all` is this module's reduction builder -- the Python builtin is _py.all)
  if out.is_contiguous() and (nd <= 1 or _py.all(ex.shape[d] == shape[d] for d in range(1, nd))):
    device_ops.fill(out.reshape(-1), kind, a, b, seed=seed, offset=extent.ravelled_pos(ex.ul, shape))
    return
  pitch = int(np.prod(shape[nd - 1:]))          # elements per step of the second-to-last index
  for idx in np.ndindex(*out.shape[:-2]):
    pos = tuple(u + i for u, i in zip(ex.ul[:-2], idx)) + tuple(ex.ul[-2:])
    view = out[idx] if idx else out
    device_ops.fill2d(view, kind, a, b, seed=seed, offset=extent.ravelled_pos(pos, shape), pitch=pitch)


def _arange_kernel(out, ex, start, stop, step, dtype=None):
  """Extent-aware iota: element = start + step * (global C-order index)."""
  _fill_by_global_index(out, ex, SP_FILL_IOTA, start, step)


_arange_mapper.device_location_kernel = _arange_kernel
_arange_mapper.result_dtype = lambda in_dtype, kw: np.dtype(kw.get('dtype') or in_dtype)


def arange(start=None, stop=None, step=1, dtype=np.float64, tile_hint=None):
  """An extended np.arange (creation.py:144-206)."""
  if start is None and stop is None:
    raise ValueError('No valid parameters')
  shape = None
  if isinstance(start, (tuple, list)):
    shape = start
    start = 0
    if stop is not None:
      start = stop
      stop = None
  elif start is None:
    start = 0
  elif stop is None:
    stop = start
    start = 0
  if shape is None and stop is None:
    raise ValueError('Shape or stop expected, none supplied.')
  if shape is not None and stop is not None:
    raise ValueError('Only shape OR stop can be supplied, not both.')
  if shape is None:
    length = int(np.ceil((stop - start) / float(step)))
    shape = (length,)
  return map_with_location(ndarray(shape, dtype, tile_hint), _arange_mapper,
                           fn_kw={'start': start, 'stop': stop, 'step': step, 'dtype': dtype})


# ------------------------------------------------------------------------------------ srandom.py
def _rand_kernel_for(kind):
  def kernel(out, ex, seed=0, dtype=None):
    # one Philox stream per array; a tile regenerates exactly its elements (global C-order offsets)
    _fill_by_global_index(out, ex, kind, seed=seed)
  return kernel


def _make_rand(input, ex, seed=0, dtype=None): raise SpartanError('host evaluation is not available')
def _make_randn(input, ex, seed=0, dtype=None): raise SpartanError('host evaluation is not available')


_make_rand.device_location_kernel = _rand_kernel_for(SP_FILL_RAND)
_make_randn.device_location_kernel = _rand_kernel_for(SP_FILL_RANDN)
_make_rand.result_dtype = _make_randn.result_dtype = lambda in_dtype, kw: np.dtype(kw.get('dtype') or in_dtype)

_seed_counter = [0]


def _random(fn, shape, kw):
  tile_hint = kw.pop('tile_hint', None)
  dtype = kw.pop('dtype', np.float64)       # srandom.py:83: np.float
  seed = kw.pop('seed', None)
  assert len(kw) == 0, 'Unknown keywords %s' % kw
  if seed is None:                           # the reference seeds from time/pid (srandom.py:24-30); a counter keeps
    _seed_counter[0] += 1                    # every rank in agreement
    seed = 0x5bd1e995 + _seed_counter[0]
  for s in shape:
    assert isinstance(s, (int, np.integer))
  return map_with_location(ndarray(shape, dtype=dtype, tile_hint=tile_hint), fn,
                           fn_kw={'seed': int(seed), 'dtype': dtype})


@not_idempotent
def rand(*shape, **kw):
  """Uniform [0, 1) (srandom.py:69-85).  Device RNG: Philox4x32-10 keyed by ``seed=``; ``dtype=`` selects
  float32/float64 (reference: float64)."""
  return _random(_make_rand, shape, kw)


@not_idempotent
def randn(*shape, **kw):
  """Standard normal (srandom.py:88-100)."""
  return _random(_make_randn, shape, kw)


# ------------------------------------------------------------------------------------ arrays.py
def _astype_mapper(t, dtype): raise SpartanError('host evaluation is not available')   # arrays.py:26-27


def _astype_handler(node, operands, analyse):
  src = analyse(node.deps[0], operands)
  return program._Typed('cast', [src], np.dtype(node.kw['dtype']))


program.register_special(_astype_mapper, _astype_handler)


def astype(x, dtype):
  """arrays.py:30-42."""
  assert x is not None
  return map(x, _astype_mapper, fn_kw={'dtype': np.dtype(dtype).str})


def size(x, axis=None):
  if axis is None:
    return int(np.prod(x.shape))
  return x.shape[axis]


# ------------------------------------------------------------------------------------ mathematics.py
def add(a, b): return map((a, b), fn=np.add)
def reciprocal(a): return map(a, fn=np.reciprocal)
def negative(a): return map(a, fn=np.negative)
def sub(a, b): return map((a, b), fn=np.subtract)
def multiply(a, b): return map((a, b), fn=np.multiply)
def divide(a, b): return map((a, b), fn=np.divide)
def floor_divide(a, b): return map((a, b), fn=np.floor_divide)
def fmod(a, b): return map((a, b), fn=np.fmod)
def mod(a, b): return map((a, b), fn=np.mod)
def remainder(a, b): return map((a, b), fn=np.remainder)      # mathematics.py:86-87 recurses forever (Q6)
def power(a, b): return map((a, b), fn=np.power)
def maximum(a, b): return map((a, b), np.maximum)
def minimum(a, b): return map((a, b), np.minimum)
def ln(v): return map(v, fn=np.log)
def log(v): return map(v, fn=np.log)
def exp(v): return map(v, fn=np.exp)
def square(v): return map(v, fn=np.square)
def sqrt(v): return map(v, fn=np.sqrt)
def abs(v): return map(v, fn=np.abs)


def _true_divide(a, b): raise SpartanError('host evaluation is not available')


def _true_divide_handler(node, operands, analyse):
  args = [analyse(d, operands) for d in node.deps]
  in_dt = program.legacy_result_type([(a.dtype, a.weak, a.value) for a in args])
  if in_dt.kind in 'biu':
    in_dt = np.dtype(np.float64)
  return program._Typed('DIV', args, in_dt, weak=_py.all(a.weak for a in args), in_dtype=in_dt)


program.register_special(_true_divide, _true_divide_handler)


def true_divide(a, b): return map((a, b), fn=_true_divide)


def _sum_local(ex, data, axis): raise SpartanError('host evaluation is not available')    # mathematics.py:126-127
def _prod_local(ex, data, axis): raise SpartanError('host evaluation is not available')   # mathematics.py:146-147


# device_reduce = (reduce op, post-ops appended to the value program, force >=53-bit accumulation)
_sum_local.device_reduce = (SP_RED_SUM, (), False)
_prod_local.device_reduce = (SP_RED_PROD, (), False)


def sum(x, axis=None, tile_hint=None):
  """Sum ``x`` over ``axis`` (mathematics.py:130-143)."""
  return reduce(x, axis=axis, dtype_fn=lambda input: input.dtype, local_reduce_fn=_sum_local,
                accumulate_fn=np.add, tile_hint=tile_hint)


def _prod_dtype_fn(input):
  """mathematics.py:150-154."""
  return np.dtype(np.int64) if input.dtype == np.int32 else input.dtype


def prod(x, axis=None, tile_hint=None):
  """mathematics.py:157-170."""
  return reduce(x, axis=axis, dtype_fn=_prod_dtype_fn, local_reduce_fn=_prod_local, accumulate_fn=np.multiply,
                tile_hint=tile_hint)


# ------------------------------------------------------------------------------------ statistics.py
def _max_local(ex, data, axis): raise SpartanError('host evaluation is not available')    # statistics.py:40
def _min_local(ex, data, axis): raise SpartanError('host evaluation is not available')    # statistics.py:59


_max_local.device_reduce = (SP_RED_MAX, (), False)
_min_local.device_reduce = (SP_RED_MIN, (), False)


def max(x, axis=None, tile_hint=None):
  """statistics.py:26-42."""
  return reduce(x, axis=axis, dtype_fn=lambda input: input.dtype, local_reduce_fn=_max_local,
                accumulate_fn=np.maximum, tile_hint=tile_hint)


def min(x, axis=None, tile_hint=None):
  """statistics.py:45-61."""
  return reduce(x, axis=axis, dtype_fn=lambda input: input.dtype, local_reduce_fn=_min_local,
                accumulate_fn=np.minimum, tile_hint=tile_hint)


def mean(x, axis=None):
  """statistics.py:64-76."""
  if axis is None:
    return true_divide(sum(x, axis), float(np.prod(x.shape)))
  return true_divide(sum(x, axis), float(x.shape[axis]))


def std(a, axis=None):
  """Standard deviation (statistics.py:86-102): sqrt(mean(a**2) - mean(a)**2) in float64, like the reference."""
  a_casted = astype(a, np.float64)
  return sqrt(sub(mean(power(a_casted, 2), axis), power(mean(a_casted, axis), 2)))


# ------------------------------------------------------------------------------------ logic.py
def _all_reducer(ex, tile, axis=None): raise SpartanError('host evaluation is not available')   # logic.py:25
def _any_reducer(ex, tile, axis=None): raise SpartanError('host evaluation is not available')   # logic.py:37


_all_reducer.device_reduce = (SP_RED_ALL, (), False)
_any_reducer.device_reduce = (SP_RED_ANY, (), False)


def all(array, axis=None):
  return reduce(array, axis=axis, dtype_fn=lambda input: np.bool_, local_reduce_fn=_all_reducer,
                accumulate_fn=np.logical_and)


def any(array, axis=None):
  return reduce(array, axis=axis, dtype_fn=lambda input: np.bool_, local_reduce_fn=_any_reducer,
                accumulate_fn=np.logical_or)


def equal(a, b): return map((a, b), fn=np.equal)
def not_equal(a, b): return map((a, b), fn=np.not_equal)
def greater(a, b): return map((a, b), fn=np.greater)
def greater_equal(a, b): return map((a, b), fn=np.greater_equal)
def less(a, b): return map((a, b), fn=np.less)
def less_equal(a, b): return map((a, b), fn=np.less_equal)
def logical_and(a, b): return map((a, b), fn=np.logical_and)
def logical_or(a, b): return map((a, b), fn=np.logical_or)
def logical_xor(a, b): return map((a, b), fn=np.logical_xor)


# ------------------------------------------------------------------------------------ sorting.py (counts)
def _countnonzero_local(ex, data, axis): raise SpartanError('host evaluation is not available')  # sorting.py:126-133
def _countzero_local(ex, data, axis): raise SpartanError('host evaluation is not available')     # sorting.py:153-157


# count_nonzero counts x != 0 (np.count_nonzero, axis=None) -- the per-axis form in the reference counts
# x > 0 (sorting.py:133), identical for the non-negative inputs its tests use; != 0 is used for both.
_countnonzero_local.device_reduce = (SP_RED_SUM, ('NONZERO',), True)
_countzero_local.device_reduce = (SP_RED_SUM, ('ISZERO',), True)


def count_nonzero(array, axis=None, tile_hint=None):
  """sorting.py:136-150."""
  return reduce(array, axis, dtype_fn=lambda input: np.int64, local_reduce_fn=_countnonzero_local,
                accumulate_fn=np.add, tile_hint=tile_hint)


def count_zero(array, axis=None):
  """sorting.py:160-172."""
  return reduce(array, axis, dtype_fn=lambda input: np.int64, local_reduce_fn=_countzero_local,
                accumulate_fn=np.add)


def argmin(x, axis=None):
  """Compute argmin over ``axis`` (sorting.py:88-105)."""
  return ArgReduceExpr(array=_as_array(x), axis=axis, which='min')


def argmax(x, axis=None):
  """Compute argmax over ``axis`` (sorting.py:108-124)."""
  return ArgReduceExpr(array=_as_array(x), axis=axis, which='max')
