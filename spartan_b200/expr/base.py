"""Expression base classes (reference: spartan/expr/operator/base.py).

`Expr` objects capture user operations lazily; ``evaluate`` walks dependencies, calls
``_evaluate(ctx, deps)`` (the node contract, base.py:315-323) and caches the resulting DistArray by
expression id (EvalCache, base.py:73-114).  As in the reference, ``evaluate()`` does not optimise;
fusion happens on ``optimized()`` (base.py:477-492).
"""
import collections
import itertools
import sys

import numpy as np

from .. import blob_ctx
from ..array import distarray
from ..config import FLAGS
from ..util import require_type, require_equal, require_unique


class newaxis(object):
  pass


class NotShapeable(Exception):
  """Raised when a shape cannot be known without evaluating (base.py:30-34)."""


unique_id = itertools.count()


class EvalCache(object):
  """base.py:73-114: results keyed by expression id, reference counted."""

  def __init__(self):
    self.refs = collections.defaultdict(int)
    self.cache = {}

  def set(self, exprid, value):
    self.cache[exprid] = value

  def get(self, exprid):
    return self.cache.get(exprid, None)

  def register(self, exprid):
    self.refs[exprid] += 1

  def deregister(self, expr_id):
    self.refs[expr_id] -= 1
    if self.refs[expr_id] == 0:
      self.cache.pop(expr_id, None)
      del self.refs[expr_id]

  def clear(self):
    self.refs.clear()
    self.cache.clear()


eval_cache = EvalCache()


def _map(*args, **kw):
  """Indirection for the operator overloads (base.py:39-47)."""
  from .map import map
  return map(args, kw['fn'])


class Expr(object):
  """Base class of all expressions (base.py:163-505).  Subclasses list their dependency
  attributes in ``members`` (the reference uses traits for this, spartan/node.py)."""
  members = ()
  needs_cache = True

  def __init__(self, expr_id=None, shape_cache=None, **kw):
    for k in self.members:
      setattr(self, k, kw.pop(k, None))
    assert not kw, 'unknown attributes %s for %s' % (list(kw), type(self).__name__)
    self.expr_id = next(unique_id) if expr_id is None else expr_id
    self.shape_cache = shape_cache
    self.optimized_expr = None
    self.is_optimized = False
    eval_cache.register(self.expr_id)
    self.needs_cache = self.needs_cache and FLAGS.opt_expression_cache

  def __del__(self):
    try:
      eval_cache.deregister(self.expr_id)
    except Exception:
      pass

  def cache(self):
    return eval_cache.get(self.expr_id)

  def dependencies(self):
    return dict((k, getattr(self, k)) for k in self.members)

  def compute_shape(self):
    raise NotShapeable

  def visit(self, visitor):
    deps, same = {}, True
    for k in self.members:
      old = getattr(self, k)
      new = deps[k] = visitor.visit(old)
      same = same and new is old
    if same:
      return self            # a pass that changed nothing below this node leaves the node itself in place
    return expr_like(self, **deps)

  def typename(self):
    return self.__class__.__name__

  def pretty_str(self):
    return '%s[%d]' % (self.typename(), self.expr_id)

  def __repr__(self):
    return self.pretty_str()

  def evaluate(self):
    """base.py:272-313."""
    cache = self.cache()
    if cache is not None:
      return cache
    ctx = blob_ctx.get()
    deps = {}
    for k, vs in self.dependencies().items():
      deps[k] = vs.evaluate() if isinstance(vs, Expr) else vs
    try:
      value = self._evaluate(ctx, deps)
    except Exception:
      sys.stderr.write('Error executing expression %s\n' % self.pretty_str())
      raise
    if self.needs_cache:
      eval_cache.set(self.expr_id, value)
    return value

  def _evaluate(self, ctx, deps):
    raise NotImplementedError

  def __hash__(self):
    return self.expr_id

  # operator overloads (base.py:331-388)
  def __add__(self, other): return _map(self, other, fn=np.add)
  def __sub__(self, other): return _map(self, other, fn=np.subtract)
  def __mul__(self, other): return _map(self, other, fn=np.multiply)
  def __mod__(self, other): return _map(self, other, fn=np.mod)
  def __div__(self, other): return _map(self, other, fn=np.divide)
  __truediv__ = __div__
  def __eq__(self, other): return _map(self, other, fn=np.equal)
  def __ne__(self, other): return _map(self, other, fn=np.not_equal)
  def __lt__(self, other): return _map(self, other, fn=np.less)
  def __gt__(self, other): return _map(self, other, fn=np.greater)
  def __and__(self, other): return _map(self, other, fn=np.logical_and)
  def __or__(self, other): return _map(self, other, fn=np.logical_or)
  def __xor__(self, other): return _map(self, other, fn=np.logical_xor)
  def __pow__(self, other): return _map(self, other, fn=np.power)
  def __neg__(self): return _map(self, fn=np.negative)
  def __rsub__(self, other): return _map(other, self, fn=np.subtract)
  def __radd__(self, other): return _map(other, self, fn=np.add)
  def __rmul__(self, other): return _map(other, self, fn=np.multiply)
  def __rdiv__(self, other): return _map(other, self, fn=np.divide)
  __rtruediv__ = __rdiv__

  def __getitem__(self, idx):
    """base.py:401-448: basic slices give a SliceExpr; integer indices and ``newaxis`` additionally drop / insert
    unit dimensions through a ReshapeExpr.  Boolean / integer-array indexing (FilterExpr) needs data-dependent
    output sizes and is not part of the device path."""
    from .slice import SliceExpr
    from .reshape import ReshapeExpr
    if not isinstance(idx, (int, tuple, slice)):
      from .program import NotDeviceMappable
      raise NotDeviceMappable('indexing with %r (filter.py) is not supported on the device path' % type(idx).__name__)
    del_dim = [x for x in range(len(idx)) if isinstance(idx[x], int)] if isinstance(idx, tuple) else []
    has_newaxis = isinstance(idx, tuple) and any(x is newaxis for x in idx)
    if not (isinstance(idx, int) or del_dim or has_newaxis):
      return SliceExpr(src=self, idx=idx)
    if isinstance(idx, tuple):
      sl = tuple(slice(x, None, None) if (isinstance(x, int) and x == -1) else x for x in idx if x is not newaxis)
    else:
      sl = slice(idx, None, None) if idx == -1 else idx
    ret = SliceExpr(src=self, idx=sl)
    new_shape = []
    if isinstance(idx, tuple):
      shape_ptr = idx_ptr = 0
      while shape_ptr < len(ret.shape) or idx_ptr < len(idx):
        if idx_ptr < len(idx) and idx[idx_ptr] is newaxis:
          new_shape.append(1)
        else:
          new_shape.append(ret.shape[shape_ptr])
          shape_ptr += 1
        idx_ptr += 1
      # positions of the integer indices in the shape WITH the inserted axes
      drop, pos = [], 0
      for x in idx:
        if isinstance(x, int):
          drop.append(pos)
        pos += 1
      new_shape = [s for i, s in enumerate(new_shape) if i not in drop]
    else:
      new_shape = list(ret.shape)[1:]
    return ReshapeExpr(array=ret, new_shape=tuple(new_shape))

  def __setitem__(self, k, val):
    raise Exception('Expressions are read-only.')

  @property
  def shape(self):
    """base.py:454-471."""
    cache = self.cache()
    if cache is not None:
      return cache.shape
    if self.shape_cache is None:
      try:
        self.shape_cache = tuple(self.compute_shape())
      except NotShapeable:
        self.shape_cache = evaluate(self).shape
    return self.shape_cache

  @property
  def ndim(self):
    return len(self.shape)

  @property
  def size(self):
    return int(np.prod(self.shape, dtype=np.int64))

  def optimized(self):
    """base.py:477-492."""
    if self.is_optimized:
      return self
    if self.optimized_expr is None:
      opt = optimized_dag(self)
      if opt is self:
        self.is_optimized = True
        return self
      self.optimized_expr = opt
      # a flag, not the reference's self-reference (`opt.optimized_expr = opt`): a cycle would keep the cached
      # result tiles -- device memory here -- alive until the cycle collector happens to run
      self.optimized_expr.is_optimized = True
    return self.optimized_expr

  def glom(self):
    return glom(self)


def expr_like(expr, **kw):
  """A new expression with the same id as ``expr`` (base.py:49-68)."""
  kw['expr_id'] = expr.expr_id
  kw['shape_cache'] = expr.shape_cache
  return expr.__class__(**kw)


class AsArray(Expr):
  """Promote a value to be array-like (base.py:508-530)."""
  members = ('val',)

  def visit(self, visitor):
    return self

  def compute_shape(self):
    if hasattr(self.val, 'shape'):
      return self.val.shape
    if np.isscalar(self.val):
      return ()
    raise NotShapeable

  def _evaluate(self, ctx, deps):
    return distarray.as_array(deps['val'])

  def pretty_str(self):
    return str(self.val)


class Val(Expr):
  """An existing value as an expression (base.py:533-556)."""
  members = ('val',)
  needs_cache = False

  def visit(self, visitor):
    return self

  def dependencies(self):
    return {}

  def compute_shape(self):
    return self.val.shape

  def _evaluate(self, ctx, deps):
    return self.val


class CollectionExpr(Expr):
  needs_cache = False
  members = ('vals',)

  def __getitem__(self, idx):
    return self.vals[idx]

  def __iter__(self):
    return iter(self.vals)

  def __len__(self):
    return len(self.vals)


class ListExpr(CollectionExpr):
  """base.py:602-626."""

  def dependencies(self):
    return dict(('v%d' % i, self.vals[i]) for i in range(len(self.vals)))

  def _evaluate(self, ctx, deps):
    return [deps['v%d' % i] for i in range(len(self.vals))]

  def visit(self, visitor):
    vals = [visitor.visit(v) for v in self.vals]
    if all(a is b for a, b in zip(vals, self.vals)):
      return self
    return ListExpr(vals=vals)


class TupleExpr(CollectionExpr):
  """base.py:629-650."""

  def dependencies(self):
    return dict(('v%d' % i, self.vals[i]) for i in range(len(self.vals)))

  def _evaluate(self, ctx, deps):
    return tuple(deps['v%d' % i] for i in range(len(self.vals)))

  def visit(self, visitor):
    vals = tuple(visitor.visit(v) for v in self.vals)
    if all(a is b for a, b in zip(vals, self.vals)):
      return self
    return TupleExpr(vals=vals)


def glom(value):
  """Evaluate and return a numpy.ndarray (base.py:652-662)."""
  if isinstance(value, Expr):
    value = evaluate(value)
  if isinstance(value, np.ndarray):
    return value
  return value.glom()


def optimized_dag(node):
  if not isinstance(node, Expr):
    raise TypeError
  from .optimize import optimize as _optimize
  return _optimize(node)


def evaluate(node):
  """base.py:679-690."""
  if isinstance(node, Expr):
    return node.evaluate()
  require_type(node, (np.ndarray, distarray.DistArray))
  return node


def eager(node):
  return Val(val=evaluate(node))


def lazify(val):
  """base.py:703-722."""
  if isinstance(val, Expr):
    return val
  if isinstance(val, list):
    return ListExpr(vals=val)
  if isinstance(val, tuple):
    return TupleExpr(vals=val)
  return Val(val=val)


def as_array(v):
  """base.py:725-734."""
  if isinstance(v, Expr):
    return v
  return AsArray(val=v)
