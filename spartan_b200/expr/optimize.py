"""Expression-graph optimisation: the two fusion passes that define "fused element-wise chain" and
"fused map+reduce" (reference: spartan/expr/operator/optimize.py:107-227, pass order :1093-1099).

AutomaticTiling lives in tiling.py (re-targeted to NVLink bytes; off by default: FLAGS.opt_auto_tiling).  Out of scope
(SURVEY.md section 2.1): RotateSlice, ParakeetGeneration, CollapsedCachedExpressions.  A map is only fused into its consumer if its local tree can run on the
device (program.tree_is_mappable); otherwise it stays a separate node, exactly like an un-fusable
child in the reference.
"""
from ..config import FLAGS
from . import program
from .base import Expr, Val, AsArray, ListExpr, expr_like
from .local import LocalInput, LocalMapLocationExpr, LocalReduceExpr, make_var
from .map import MapExpr
from .ndarray import NdArrayExpr
from .reduce import ReduceExpr

_not_idempotent_list = set()


def not_idempotent(fn):
  """optimize.py:60-69: the result of ``fn`` must be evaluated once (random arrays), never re-fused."""
  def wrapped(*args, **kw):
    result = fn(*args, **kw)
    if isinstance(result, Expr):
      result.needs_cache = True
      _not_idempotent_list.add(result.expr_id)
    return result
  wrapped.__name__ = getattr(fn, '__name__', 'wrapped')
  return wrapped


class OptimizePass(object):
  """optimize.py:80-104."""

  def __init__(self):
    self.visited = {}

  def visit(self, op):
    if not isinstance(op, Expr):
      return op
    if op.expr_id in self.visited:
      return self.visited[op.expr_id]
    name = 'visit_%s' % op.typename()
    opt_op = getattr(self, name)(op) if hasattr(self, name) else op.visit(self)
    self.visited[opt_op.expr_id] = opt_op
    return opt_op


def fusable(v):
  """optimize.py:107-116 restricted to the node kinds that exist here."""
  from .write_array import WriteArrayExpr
  return isinstance(v, (MapExpr, ReduceExpr, NdArrayExpr, Val, AsArray, WriteArrayExpr))


def _fusable_map(child):
  return (isinstance(child, MapExpr) and child.expr_id not in _not_idempotent_list
          and not isinstance(child.op, LocalMapLocationExpr) and program.tree_is_mappable(child.op))


def merge_var(children, child_to_var, k, v):
  """optimize.py:119-130."""
  try:
    i = child_to_var.index(k)
    assert children[i] is v or children[i].expr_id == v.expr_id
  except ValueError:
    children.append(v)
    child_to_var.append(k)


class MapMapFusion(OptimizePass):
  """map(f, map(g, map(h, x))) -> map(f . g . h, x)   (optimize.py:133-187)."""
  name = 'map_fusion'

  def visit_MapExpr(self, expr):
    map_children = self.visit(expr.children)
    all_maps = all(fusable(v) for v in map_children)
    if (not all_maps or expr.expr_id in _not_idempotent_list or isinstance(expr.op, LocalMapLocationExpr)
        or not program.tree_is_mappable(expr.op)):
      return expr.visit(self)
    children, child_to_var = [], []
    combined_op = expr.op.__class__(fn=expr.op.fn, kw=expr.op.kw, pretty_fn=expr.op.pretty_fn)
    for child_expr in map_children:
      if _fusable_map(child_expr):
        for k, v in zip(child_expr.child_to_var, child_expr.children):
          merge_var(children, child_to_var, k, v)
        combined_op.add_dep(child_expr.op)
      else:
        children.append(child_expr)
        key = make_var()
        combined_op.add_dep(LocalInput(idx=key))
        child_to_var.append(key)
    if not program.fits_one_kernel(combined_op, children, child_to_var):
      return expr.visit(self)       # too deep / too many operands for one kernel: keep the children as nodes
    return expr_like(expr, children=ListExpr(vals=children), child_to_var=child_to_var, op=combined_op)


class ReduceMapFusion(OptimizePass):
  """reduce(f, map(g, X)) -> reduce(f . g, X)   (optimize.py:190-227)."""
  name = 'reduce_fusion'

  def visit_ReduceExpr(self, expr):
    old_children = self.visit(expr.children)
    for v in old_children:
      if not _fusable_map(v):
        return expr.visit(self)
    combined_op = LocalReduceExpr(fn=expr.op.fn, kw=expr.op.kw, deps=[expr.op.deps[0]])
    new_children, new_child_to_var = [], []
    for child_expr in old_children:
      for k, v in zip(child_expr.child_to_var, child_expr.children):
        merge_var(new_children, new_child_to_var, k, v)
      combined_op.add_dep(child_expr.op)
    return expr_like(expr, children=ListExpr(vals=new_children), child_to_var=new_child_to_var, axis=expr.axis,
                     dtype_fn=expr.dtype_fn, accumulate_fn=expr.accumulate_fn, op=combined_op,
                     tile_hint=expr.tile_hint)


def _auto_tiling():
  from .tiling import AutomaticTiling
  return AutomaticTiling


class _AutoTilingPass(object):
  """optimize.py:1094 registers AutomaticTiling before the fusion passes; resolved lazily (tiling.py imports dot.py)."""
  name = 'auto_tiling'

  def __new__(cls):
    return _auto_tiling()()


passes = [_AutoTilingPass, MapMapFusion, ReduceMapFusion]     # optimize.py:1093-1099 order (the in-scope passes)


def apply_pass(klass, dag):
  if not getattr(FLAGS, 'opt_' + klass.name):
    return dag
  return klass().visit(dag)


def optimize(dag):
  """optimize.py:1072-1081."""
  if not FLAGS.optimization:
    return dag
  for p in passes:
    dag = apply_pass(p, dag)
  return dag
