"""Lowers a LocalExpr tree to the postfix bytecode the CUDA evaluator runs (sp_program).

This replaces ``FnCallExpr.evaluate`` (spartan/expr/operator/local.py:115-127), which calls one
NumPy ufunc per node and materialises a temporary per call.  The compiler reproduces the dtype
semantics NumPy applied in the reference's era:

  * per-ufunc result dtypes (comparisons -> bool, sqrt/exp/log of ints -> float64, ...);
  * value-based casting of scalars: a 0-d operand never widens an array operand of the same or a
    higher kind (``x * 2`` stays float32) -- under NumPy 2 the same call
    would produce float64, so the rule is encoded here explicitly;
  * ``np.divide`` on two integer operands floors (Python 2 ``/`` -> np.divide, base.py:346-347).

The whole tree runs in ONE register type (float32, float64 or int64 = the widest dtype any node
needs).  Nodes whose NumPy dtype is narrower get an explicit CAST so rounding happens where NumPy
would round (float32 results inside a float64 program, int32 wrap inside an int64 program).

Python callables that are not in the table below cannot run on the GPU: compilation raises
``NotDeviceMappable`` (there is no CPU fallback).
"""
import numpy as np

from . import local
from .. import device_ops
from .._lib import SpartanError, SP_F32, SP_F64, SP_I64, SP_MAX_STACK, SP_MAX_OPERANDS, SP_MAX_CONSTS


class NotDeviceMappable(SpartanError):
  pass


# ufunc -> (opcode, class)
#   'arith'  : result dtype = promoted input dtype
#   'cmp'    : bool result, evaluated in the promoted input dtype
#   'logic'  : bool result
#   'div'    : np.divide (ints floor, reference era), 'tdiv' : true_divide (ints -> float64)
#   'float'  : float result (ints promote to float64)
_UFUNCS = {
  np.add: ('ADD', 'arith'), np.subtract: ('SUB', 'arith'), np.multiply: ('MUL', 'arith'),
  np.divide: ('DIV', 'div'), np.true_divide: ('DIV', 'div'), np.floor_divide: ('FLOORDIV', 'arith'),
  np.mod: ('MOD', 'arith'), np.remainder: ('MOD', 'arith'), np.fmod: ('FMOD', 'arith'),
  np.power: ('POW', 'arith'), np.maximum: ('MAX', 'arith'), np.minimum: ('MIN', 'arith'),
  np.equal: ('EQ', 'cmp'), np.not_equal: ('NE', 'cmp'), np.less: ('LT', 'cmp'), np.less_equal: ('LE', 'cmp'),
  np.greater: ('GT', 'cmp'), np.greater_equal: ('GE', 'cmp'),
  np.logical_and: ('AND', 'logic'), np.logical_or: ('OR', 'logic'), np.logical_xor: ('XOR', 'logic'),
  np.logical_not: ('NOT', 'logic'),
  np.negative: ('NEG', 'arith'), np.abs: ('ABS', 'arith'), np.absolute: ('ABS', 'arith'),
  np.square: ('SQUARE', 'arith'), np.sqrt: ('SQRT', 'float'), np.exp: ('EXP', 'float'), np.log: ('LOG', 'float'),
  np.reciprocal: ('RECIP', 'arith'),
}
# np.true_divide differs from np.divide only for integer operands
_TRUE_DIVIDE = np.true_divide if np.true_divide is not np.divide else None

_KIND_RANK = {'b': 0, 'u': 1, 'i': 1, 'f': 2}
_CAST_OP = {np.dtype(np.float32): 'CAST_F32', np.dtype(np.int64): 'CAST_I64', np.dtype(np.int32): 'CAST_I32',
            np.dtype(np.bool_): 'CAST_BOOL', np.dtype(np.uint8): 'CAST_U8'}


def legacy_result_type(items):
  """items: [(dtype, weak, value-or-None)].  NumPy 1.x ufunc type resolution with value-based scalars."""
  arrays = [d for d, weak, _ in items if not weak]
  scalars = [(d, v) for d, weak, v in items if weak]
  if not arrays or not scalars:
    return np.result_type(*[d for d, _, _ in items])
  max_arr = max(_KIND_RANK[d.kind] for d in arrays)
  max_sc = max(_KIND_RANK[d.kind] for d, _ in scalars)
  if max_sc <= max_arr:
    dts = list(arrays) + [np.min_scalar_type(v) if v is not None else d for d, v in scalars]
    return np.result_type(*dts)
  return np.result_type(*[d for d, _, _ in items])


class Operand(object):
  """What a LocalInput name is bound to when a tile kernel is launched."""
  __slots__ = ('kind', 'dtype', 'value', 'index', 'alias')

  def __init__(self, kind, dtype, value=None, index=None, alias=None):
    self.kind = kind          # 'array' (device tensor operand) | 'scalar' (host 0-d value)
    self.dtype = np.dtype(dtype)
    self.value = value
    self.index = index
    self.alias = alias        # name of an earlier input bound to the very same array: loaded once


class _Typed(object):
  """Typed tree node produced by the analysis pass."""
  __slots__ = ('op', 'args', 'dtype', 'weak', 'value', 'in_dtype', 'leaf')

  def __init__(self, op, args, dtype, weak=False, value=None, in_dtype=None, leaf=None):
    self.op, self.args, self.dtype, self.weak, self.value = op, args, np.dtype(dtype), weak, value
    self.in_dtype = None if in_dtype is None else np.dtype(in_dtype)
    self.leaf = leaf


class CompiledProgram(object):
  def __init__(self, program, out_dtype, used_vars, compute_dtype):
    self.program = program          # sp_program
    self.out_dtype = out_dtype      # numpy dtype of the expression value
    self.used_vars = used_vars      # LocalInput names in operand order (arrays only)
    self.compute_dtype = compute_dtype


# functions the reference's builders put into the tree that are not ufuncs; registered by the modules that
# define them:  fn -> handler(node, typed_args, kw) -> _Typed
_SPECIAL = {}


def register_special(fn, handler):
  _SPECIAL[fn] = handler


def is_mappable_fn(fn):
  return fn in _UFUNCS or fn in _SPECIAL


def tree_is_mappable(op):
  if isinstance(op, local.LocalInput):
    return True
  if isinstance(op, local.LocalMapLocationExpr):
    return False
  if isinstance(op, local.FnCallExpr):
    return is_mappable_fn(op.fn) and all(tree_is_mappable(d) for d in op.deps)
  return False


def tree_cost(op):
  """(evaluation-stack depth with left-to-right operand order, set of input names) of a LocalExpr tree."""
  if isinstance(op, local.LocalInput):
    return 1, {op.idx}
  depth, names = 1, set()
  for i, d in enumerate(op.deps):
    dd, nn = tree_cost(d)
    depth = max(depth, i + dd)
    names |= nn
  return depth, names


def fits_one_kernel(op, children=None, child_to_var=None):
  """True when the tree can run as ONE fused kernel: stack depth, distinct array operands and scalar constants within
  the evaluator's budget.  Fusion passes stop fusing at this boundary; the un-fused child is then evaluated (once,
  cached by expression id) into a temporary array.  With ``children`` / ``child_to_var`` given, an array that is used
  several times counts once (it is bound to one kernel operand, map.bind_operands) and Python scalars count as
  immediates, not operands."""
  depth, names = tree_cost(op)
  if depth > SP_MAX_STACK:
    return False
  if children is None:
    return len(names) <= SP_MAX_OPERANDS
  arrays, consts = set(), 0
  for child, var in zip(children, child_to_var):
    if var not in names:
      continue
    val = getattr(child, 'val', None)
    if type(child).__name__ == 'AsArray' and (np.isscalar(val) or getattr(val, 'shape', None) == ()):
      consts += 1
    else:
      arrays.add(child.expr_id)
  return len(arrays) <= SP_MAX_OPERANDS and consts <= SP_MAX_CONSTS


def _analyse(node, operands):
  if isinstance(node, local.LocalInput):
    if node.idx not in operands:
      raise NotDeviceMappable('input %s is not bound' % node.idx)
    o = operands[node.idx]
    if o.kind == 'scalar':
      return _Typed('const', [], o.dtype, weak=True, value=o.value)
    return _Typed('in', [], o.dtype, leaf=o.alias or node.idx)
  if not isinstance(node, local.FnCallExpr):
    raise NotDeviceMappable('cannot lower %r' % (node,))
  if node.fn in _SPECIAL:
    return _SPECIAL[node.fn](node, operands, _analyse)
  if node.fn not in _UFUNCS:
    raise NotDeviceMappable('function %s is not GPU-mappable: only NumPy ufuncs and the built-in tile functions '
                            'run on the device (no CPU fallback)' % getattr(node.fn, '__name__', node.fn))
  opcode, klass = _UFUNCS[node.fn]
  args = [_analyse(d, operands) for d in node.deps]
  in_dt = legacy_result_type([(a.dtype, a.weak, a.value) for a in args])
  weak = all(a.weak for a in args)
  if klass == 'arith':
    out = in_dt
    if in_dt == np.bool_:
      # NumPy maps add/multiply/maximum/minimum on bools to logical ops; the rest are not defined
      remap = {'ADD': 'OR', 'MUL': 'AND', 'MAX': 'OR', 'MIN': 'AND'}
      if opcode not in remap:
        raise NotDeviceMappable('%s on boolean operands' % opcode)
      opcode = remap[opcode]
  elif klass == 'div':
    if in_dt.kind in 'biu':
      if node.fn is _TRUE_DIVIDE:
        in_dt = np.dtype(np.float64); out = in_dt
      else:
        opcode = 'FLOORDIV'; out = in_dt if in_dt.kind != 'b' else np.dtype(np.int8)
        if in_dt.kind == 'b':
          raise NotDeviceMappable('divide on boolean operands')
    else:
      out = in_dt
  elif klass == 'float':
    if in_dt.kind in 'biu':
      in_dt = np.dtype(np.float64)
    out = in_dt
  elif klass == 'cmp':
    out = np.dtype(np.bool_)
  else:  # logic
    out = np.dtype(np.bool_)
    in_dt = in_dt
  return _Typed(opcode, args, out, weak=weak, in_dtype=in_dt)


def _collect_dtypes(t, acc):
  acc.append(t.dtype)
  if t.in_dtype is not None:
    acc.append(t.in_dtype)
  for a in t.args:
    _collect_dtypes(a, acc)


def _compute_dtype(dtypes, force=None):
  if force is not None:
    return force
  if any(d == np.float64 for d in dtypes):
    return SP_F64
  if any(d.kind == 'f' for d in dtypes):
    if any(d.kind in 'iu' and d.itemsize >= 4 for d in dtypes):
      return SP_F64          # int32/int64 meet float32 -> NumPy computes in float64
    return SP_F32
  return SP_I64


def _needs_cast(dtype, compute):
  dtype = np.dtype(dtype)
  if compute == SP_F64:
    return dtype == np.float32 or dtype.kind in 'iu'
  if compute == SP_F32:
    return dtype.kind in 'iu'
  if compute == SP_I64:
    return dtype in (np.dtype(np.int32), np.dtype(np.uint8), np.dtype(np.int8), np.dtype(np.int16))
  return False


def _cast_opcode(dtype):
  dtype = np.dtype(dtype)
  if dtype in _CAST_OP:
    return _CAST_OP[dtype]
  if dtype.kind in 'iu':
    return 'CAST_I64' if dtype.itemsize == 8 else 'CAST_I32'
  raise NotDeviceMappable('cannot cast to %s on the device' % dtype)


_COMMUTATIVE = ('ADD', 'MUL', 'MAX', 'MIN', 'EQ', 'NE', 'AND', 'OR', 'XOR')


class _Emitter(object):
  def __init__(self, compute):
    self.compute = compute
    self.ops = []
    self.consts = []
    self.vars = []
    self.depth = 0
    self.max_depth = 0

  def _push(self):
    self.depth += 1
    self.max_depth = max(self.max_depth, self.depth)

  def const(self, value):
    value = float(value) if self.compute != SP_I64 else int(value)
    if value not in self.consts:
      if len(self.consts) >= SP_MAX_CONSTS:
        raise NotDeviceMappable('more than %d distinct scalar constants in one fused expression' % SP_MAX_CONSTS)
      self.consts.append(value)
    self.ops.append(('CONST', self.consts.index(value)))
    self._push()

  def emit(self, t):
    if t.op == 'const':
      v = np.asarray(t.value).astype(t.dtype)[()]      # the value as NumPy would have cast it
      self.const(v.item() if hasattr(v, 'item') else v)
      return
    if t.op == 'in':
      if t.leaf not in self.vars:
        if len(self.vars) >= SP_MAX_OPERANDS:
          raise NotDeviceMappable('more than %d array operands in one fused expression' % SP_MAX_OPERANDS)
        self.vars.append(t.leaf)
      self.ops.append(('IN', self.vars.index(t.leaf)))
      self._push()
      return
    if t.op == 'cast':
      self.emit(t.args[0])
      if t.dtype.kind == 'b':
        self.ops.append(('CAST_BOOL', 0))
      elif _needs_cast(t.dtype, self.compute) or (t.dtype.kind in 'iu' and self.compute != SP_I64):
        self.ops.append((_cast_opcode(t.dtype), 0))
      elif t.dtype.kind in 'iu' and t.args[0].dtype.kind == 'f':
        self.ops.append((_cast_opcode(t.dtype), 0))
      return
    # evaluate deeper operands first (Sethi-Ullman order is not needed: commutativity is not assumed,
    # operands are simply evaluated left to right)
    args = t.args
    if len(args) == 2 and t.op in _COMMUTATIVE and args[0].op == 'const' and args[1].op != 'const':
      args = [args[1], args[0]]       # canonical operand order (array first): IEEE add/mul/min/max commute exactly
    for a in args:
      self.emit(a)
    self.ops.append((t.op, 0))
    self.depth -= (len(t.args) - 1)
    if t.dtype.kind != 'b' and _needs_cast(t.dtype, self.compute):
      self.ops.append((_cast_opcode(t.dtype), 0))


def _fix_weak_consts(t):
  """A weak (0-d) operand is converted to the dtype the ufunc runs in before the loop executes."""
  if t.in_dtype is not None:
    for a in t.args:
      if a.op == 'const':
        a.dtype = t.in_dtype
  for a in t.args:
    _fix_weak_consts(a)


def compile_tree(op, operands, force_compute=None, post_ops=()):
  """op: LocalExpr tree; operands: {var name: Operand}.  Returns CompiledProgram."""
  typed = _analyse(op, operands)
  _fix_weak_consts(typed)
  dts = []
  _collect_dtypes(typed, dts)
  compute = _compute_dtype(dts, force_compute)
  em = _Emitter(compute)
  em.emit(typed)
  for p in post_ops:
    em.ops.append((p, 0))
  if em.max_depth > SP_MAX_STACK:
    raise NotDeviceMappable('fused expression needs an evaluation stack of %d (limit %d)' % (em.max_depth, SP_MAX_STACK))
  prog = device_ops.make_program(em.ops, compute, em.consts)
  return CompiledProgram(prog, typed.dtype, em.vars, compute)
