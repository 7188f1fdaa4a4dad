"""Tile-local expression IR (reference: spartan/expr/operator/local.py:35-152).

Global expressions are over arrays, local expressions are over tiles.  A LocalExpr tree is what the
fusion passes build (optimize.py:133-227) and what ``program.compile_tree`` lowers to the bytecode
the CUDA evaluator runs; nothing in this module evaluates on the host.
"""
import itertools

_var_id = itertools.count()


def make_var():
  """local.py:30-32."""
  return 'key_%d' % next(_var_id)


class LocalExpr(object):
  """An operation performed in the context of one tile (local.py:39-56)."""

  def __init__(self, deps=None):
    self.deps = list(deps or [])

  def add_dep(self, v):
    self.deps.append(v)

  def input_names(self):
    names = []
    for d in self.deps:
      for n in d.input_names():
        if n not in names:
          names.append(n)
    return names

  def __repr__(self):
    return self.pretty_str()


class LocalInput(LocalExpr):
  """An externally supplied input (local.py:58-70)."""

  def __init__(self, idx):
    LocalExpr.__init__(self)
    assert idx != ''
    self.idx = idx

  def pretty_str(self):
    return '%s' % self.idx

  def input_names(self):
    return [self.idx]


class FnCallExpr(LocalExpr):
  """A function call over tile values (local.py:73-127); ``kw`` carries constants (axis, dtype...)."""

  def __init__(self, fn, kw=None, deps=None, pretty_fn=None):
    LocalExpr.__init__(self, deps)
    assert fn is not None
    self.fn = fn
    self.kw = kw if kw is not None else {}
    self.pretty_fn = pretty_fn

  def fn_name(self):
    if self.pretty_fn:
      return self.pretty_fn
    return getattr(self.fn, '__name__', repr(self.fn))

  def pretty_str(self):
    return '%s(%s)' % (self.fn_name().split('.')[-1], ','.join(d.pretty_str() for d in self.deps))


class LocalMapExpr(FnCallExpr):
  _op_type = 'map'


class LocalMapLocationExpr(LocalMapExpr):
  _op_type = 'map_location'


class LocalReduceExpr(FnCallExpr):
  _op_type = 'reduce'
