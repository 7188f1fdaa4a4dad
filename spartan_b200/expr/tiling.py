"""AutomaticTiling, re-targeted from TCP bytes to NVLink bytes (reference: spartan/expr/operator/optimize.py:459-1054,
solver spartan/expr/operator/tiling.cc:156-410).

The reference gives every array node of an expression up to three candidate tilings -- 0: split rows, 1: split
columns, 2: split both (optimize.py:546-575) -- charges every operator the network traffic its operands' tilings
cause (the ``cost_model`` tables, :490-511; per-operator rules in visit_MapExpr / visit_ReduceExpr / visit_DotExpr,
:577-760) and asks a min-cost-flow solver for the cheapest consistent choice (``calc_tiling``, :1009-1054), then
writes the winners into the ``tile_hint`` of the nodes that create arrays.

Same pass here with the cost of THIS backend: bytes that cross NVLink per evaluation on ``num_workers`` GPUs, given how
the evaluator actually moves data (spartan_b200/expr/map.py, reduce.py, dot.py):

  map       an operand that is not tiled like the largest operand is fetched piecewise by the owners of the output
            tiles: (W-1)/W of its bytes; operands tiled alike cost nothing;
  reduce    a tiling that splits the reduced axis leaves one partial per rank: an all-reduce of the output,
            2 (W-1)/W x output bytes per rank; any other tiling reduces locally;
  dot       owner computes: a column-tiled C needs its B columns where they are and every A strip once per rank
            ((W-1) x |A| in total, the copy-engine exchange of dot.py); a row-tiled C is the mirror image
            ((W-1) x |B|); operands tiled the other way are re-partitioned first ((W-1)/W of their bytes);
  transpose swaps the meaning of rows and columns, free (a view).

Expression DAGs are small, so instead of a flow network the free choices (arrays created without a tile_hint, dot
results without one) are searched exhaustively up to ``EXHAUSTIVE_LIMIT`` free nodes and by coordinate descent
beyond.  Fixed nodes (existing arrays, explicit hints) keep what they have.
"""
import itertools

import numpy as np

from .. import blob_ctx
from ..array import distarray
from .base import Expr, Val, AsArray, ListExpr, TupleExpr, expr_like
from .dot import DotExpr
from .map import MapExpr
from .ndarray import NdArrayExpr
from .reduce import ReduceExpr
from .transpose import TransposeExpr

ROW, COL, BLOCK = 0, 1, 2          # the reference's tiling ids (optimize.py:546-575)
EXHAUSTIVE_LIMIT = 9


def hint_for(shape, tiling, num_workers):
  """tile_hint that realises ``tiling`` for ``shape`` on ``num_workers`` ranks (round-robin placement then gives rank r
  row block r / column block r / column block r of a W x W grid)."""
  shape = tuple(int(s) for s in shape)
  W = max(1, int(num_workers))
  if len(shape) < 2:
    return tuple(max(1, -(-s // W)) for s in shape) if shape else None
  rows, cols = shape[0], shape[-1]
  if tiling == ROW:
    return (max(1, -(-rows // W)),) + shape[1:]
  if tiling == COL:
    return shape[:-1] + (max(1, -(-cols // W)),)
  return (max(1, -(-rows // W)),) + shape[1:-1] + (max(1, -(-cols // W)),)


def tiling_of_tiles(shape, tile_shape):
  """Classify an existing tiling: ROW if tiles span all columns, COL if they span all rows, else BLOCK."""
  if len(shape) < 2:
    return ROW
  full_rows = tile_shape[0] >= shape[0]
  full_cols = tile_shape[-1] >= shape[-1]
  if full_cols and not full_rows:
    return ROW
  if full_rows and not full_cols:
    return COL
  if full_rows and full_cols:
    return ROW            # one tile: every tiling is free to assume
  return BLOCK


def _nbytes(expr):
  try:
    shape = expr.shape
  except Exception:
    return 0
  dtype = getattr(expr, 'dtype', None)
  if isinstance(expr, Val):
    dtype = getattr(expr.val, 'dtype', dtype)
  try:
    item = np.dtype(dtype).itemsize if dtype is not None else 4
  except TypeError:
    item = 4
  return int(np.prod(shape, dtype=np.int64)) * item


class _Node(object):
  __slots__ = ('expr', 'kind', 'children', 'fixed', 'nbytes', 'shape', 'axis')

  def __init__(self, expr, kind, children=(), fixed=None, axis=None):
    self.expr, self.kind, self.children, self.fixed, self.axis = expr, kind, list(children), fixed, axis
    self.nbytes = _nbytes(expr)
    try:
      self.shape = tuple(expr.shape)
    except Exception:
      self.shape = ()


class AutomaticTiling(object):
  """optimize.py:459-1054 as a pass over this package's node kinds.  ``visit(dag)`` returns a DAG whose free creation
  nodes and dot nodes carry the chosen tile_hint; ``plan`` / ``cost`` are kept for inspection."""
  name = 'auto_tiling'

  def __init__(self, num_workers=None):
    self.W = int(num_workers if num_workers is not None else blob_ctx.get().num_workers)
    self.nodes = {}          # expr_id -> _Node
    self.order = []          # topological (children first)
    self.plan = {}
    self.cost = 0.0

  # ------------------------------------------------------------------ graph
  def _build(self, expr):
    if not isinstance(expr, Expr):
      return None
    if expr.expr_id in self.nodes:
      return expr.expr_id
    if isinstance(expr, (ListExpr, TupleExpr)):
      return None
    node = None
    if isinstance(expr, NdArrayExpr):
      fixed = None if expr.tile_hint is None else tiling_of_tiles(expr.shape, tuple(expr.tile_hint))
      node = _Node(expr, 'leaf', fixed=fixed)
    elif isinstance(expr, (Val, AsArray)):
      val = expr.val
      if isinstance(val, distarray.DistArrayImpl):
        node = _Node(expr, 'leaf', fixed=tiling_of_tiles(val.shape, val.tile_shape()))
      else:
        node = _Node(expr, 'scalar', fixed=ROW)
    elif isinstance(expr, MapExpr):
      kids = [self._build(c) for c in expr.children]
      node = _Node(expr, 'map', [k for k in kids if k is not None])
    elif isinstance(expr, ReduceExpr):
      kids = [self._build(c) for c in expr.children]
      node = _Node(expr, 'reduce', [k for k in kids if k is not None], axis=expr.axis)
    elif isinstance(expr, DotExpr):
      kids = [self._build(expr.matrix_a), self._build(expr.matrix_b)]
      fixed = None
      if expr.tile_hint is not None and len(expr.shape) == 2:
        fixed = tiling_of_tiles(expr.shape, tuple(expr.tile_hint))
      node = _Node(expr, 'dot', [k for k in kids if k is not None], fixed=fixed)
    elif isinstance(expr, TransposeExpr):
      kid = self._build(expr.array)
      node = _Node(expr, 'transpose', [kid] if kid is not None else [])
    else:
      kids = [self._build(v) for v in expr.dependencies().values() if isinstance(v, Expr)]
      hint = getattr(expr, 'tile_hint', None)
      fixed = tiling_of_tiles(expr.shape, tuple(hint)) if hint is not None and len(expr.shape) >= 2 else None
      node = _Node(expr, 'other', [k for k in kids if k is not None], fixed=fixed)
    self.nodes[expr.expr_id] = node
    self.order.append(expr.expr_id)
    return expr.expr_id

  def free_nodes(self):
    out = []
    for nid in self.order:
      n = self.nodes[nid]
      if n.fixed is None and len(n.shape) >= 2 and (n.kind == 'leaf' or n.kind == 'dot'):
        out.append(nid)
    return out

  # ------------------------------------------------------------------ cost
  def evaluate(self, choice):
    """(total NVLink bytes, {node: tiling}) for the free-node assignment ``choice`` ({expr_id: tiling})."""
    W = self.W
    frac = (W - 1.0) / W
    tiling, cost = {}, 0.0
    for nid in self.order:
      n = self.nodes[nid]
      kids = n.children
      if n.kind in ('leaf', 'scalar'):
        tiling[nid] = n.fixed if n.fixed is not None else choice.get(nid, ROW)
      elif n.kind == 'map':
        arrays = [k for k in kids if self.nodes[k].kind != 'scalar' and len(self.nodes[k].shape) >= 1]
        if not arrays:
          tiling[nid] = ROW
          continue
        largest = max(arrays, key=lambda k: self.nodes[k].nbytes)
        t = tiling[largest]
        tiling[nid] = t
        for k in arrays:
          if k != largest and len(self.nodes[k].shape) >= 2 and tiling[k] != t:
            cost += frac * self.nodes[k].nbytes          # fetched piecewise by the owners of the output tiles
      elif n.kind == 'reduce':
        t = tiling[kids[0]] if kids else ROW
        in_shape = self.nodes[kids[0]].shape if kids else ()
        axis = n.axis
        if axis is not None and axis < 0:
          axis += len(in_shape)
        splits = {ROW: (0,), COL: (len(in_shape) - 1,), BLOCK: (0, len(in_shape) - 1)}[t] if len(in_shape) >= 2 else (0,)
        if W > 1 and (axis is None or axis in splits):
          cost += 2.0 * frac * max(n.nbytes, 8)            # all-reduce of the output
        tiling[nid] = ROW
      elif n.kind == 'dot':
        a, b = kids[0], kids[1] if len(kids) > 1 else None
        t = n.fixed if n.fixed is not None else choice.get(nid, COL)
        tiling[nid] = t
        if W > 1 and len(n.shape) == 2 and b is not None:
          na, nb = self.nodes[a].nbytes, self.nodes[b].nbytes
          if t == ROW:
            cost += (W - 1.0) * nb                         # every rank needs all of B
            if tiling[a] != ROW:
              cost += frac * na
          else:                                            # COL / BLOCK: column blocks of C
            cost += (W - 1.0) * na                         # every rank needs every A strip once
            if tiling[b] == ROW:
              cost += frac * nb
      elif n.kind == 'transpose':
        t = tiling[kids[0]] if kids else ROW
        tiling[nid] = {ROW: COL, COL: ROW, BLOCK: BLOCK}[t]
      else:
        tiling[nid] = n.fixed if n.fixed is not None else (tiling[kids[0]] if kids else ROW)
    return cost, tiling

  def solve(self):
    free = self.free_nodes()
    options = {}
    for nid in free:
      options[nid] = (ROW, COL) if self.nodes[nid].kind == 'dot' else (ROW, COL, BLOCK)
    best_choice, best_cost, best_tiling = {}, None, None
    if len(free) <= EXHAUSTIVE_LIMIT:
      for combo in itertools.product(*[options[n] for n in free]):
        choice = dict(zip(free, combo))
        c, t = self.evaluate(choice)
        if best_cost is None or c < best_cost:
          best_choice, best_cost, best_tiling = choice, c, t
    else:
      choice = dict((n, options[n][0]) for n in free)
      best_cost, best_tiling = self.evaluate(choice)
      improved = True
      while improved:
        improved = False
        for n in free:
          for opt in options[n]:
            if opt == choice[n]:
              continue
            trial = dict(choice); trial[n] = opt
            c, t = self.evaluate(trial)
            if c < best_cost:
              choice, best_cost, best_tiling, improved = trial, c, t, True
      best_choice = choice
    self.plan, self.cost = best_choice, (best_cost or 0.0)
    return best_choice, self.cost, best_tiling

  # ------------------------------------------------------------------ rewrite
  def visit(self, dag):
    if self.W <= 1 or not isinstance(dag, Expr):
      return dag
    self._build(dag)
    choice, _, _ = self.solve()
    if not choice:
      return dag
    return _Rewrite(choice, self.W).visit(dag)


class _Rewrite(object):
  def __init__(self, choice, W):
    self.choice, self.W, self.visited = choice, W, {}

  def visit(self, op):
    if not isinstance(op, Expr):
      return op
    if op.expr_id in self.visited:
      return self.visited[op.expr_id]
    if op.expr_id in self.choice and isinstance(op, NdArrayExpr):
      out = expr_like(op, _shape=op._shape, sparse=op.sparse, dtype=op.dtype, reduce_fn=op.reduce_fn,
                      tile_hint=hint_for(op.shape, self.choice[op.expr_id], self.W))
    elif op.expr_id in self.choice and isinstance(op, DotExpr):
      out = expr_like(op, matrix_a=self.visit(op.matrix_a), matrix_b=self.visit(op.matrix_b),
                      tile_hint=hint_for(op.shape, self.choice[op.expr_id], self.W))
    else:
      out = op.visit(self)
    self.visited[op.expr_id] = out
    return out
