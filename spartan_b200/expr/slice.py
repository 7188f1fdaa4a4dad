"""Indexing with slices (reference: spartan/expr/operator/slice.py:88-137 SliceExpr; the view class lives in
spartan_b200/array/views.py)."""
from ..array import extent
from ..array.distarray import Broadcast
from ..array.views import Slice
from .base import Expr, NotShapeable, ListExpr, TupleExpr


class SliceExpr(Expr):
  """Represents an indexing operation (slice.py:88-137)."""
  members = ('src', 'idx', 'broadcast_to')

  def __init__(self, *args, **kw):
    super(SliceExpr, self).__init__(*args, **kw)
    assert not isinstance(self.src, ListExpr)
    assert not isinstance(self.idx, (ListExpr, TupleExpr))

  def dependencies(self):
    return {'src': self.src, 'idx': self.idx}

  def visit(self, visitor):
    return SliceExpr(src=visitor.visit(self.src), idx=self.idx, broadcast_to=self.broadcast_to,
                     expr_id=self.expr_id, shape_cache=self.shape_cache)

  def compute_shape(self):
    if isinstance(self.idx, (int, slice, tuple)):
      ex = extent.from_shape(self.src.shape)
      return extent.compute_slice(ex, self.idx).shape
    raise NotShapeable

  def pretty_str(self):
    return 'Slice[%d](%s, %s)' % (self.expr_id, self.src, self.idx)

  def _evaluate(self, ctx, deps):
    src, idx = deps['src'], deps['idx']
    assert not isinstance(idx, list)
    if self.broadcast_to is not None and tuple(src.shape) != tuple(self.broadcast_to):
      src = Broadcast(src, self.broadcast_to)
    return Slice(src, idx)
