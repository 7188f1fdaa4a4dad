"""Reshape operation and expr (reference: spartan/expr/operator/reshape.py:196-239)."""
from ..array.views import Reshape
from ..util import require_type, require_equal, require_unique
from .base import Expr, lazify


class ReshapeExpr(Expr):
  members = ('array', 'new_shape', 'tile_hint')

  def __str__(self):
    return 'Reshape[%d] %s to %s' % (self.expr_id, self.array, self.new_shape)

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
  """Reshape/retile ``array`` (reshape.py:212-239): reshape(a, (m, n)) or reshape(a, m, n); tile_hint= optional."""
  if len(args) == 1 and isinstance(args[0], (tuple, list)):
    new_shape = tuple(args[0])
  else:
    new_shape = tuple(args)
  require_type(new_shape, tuple)
  return ReshapeExpr(array=lazify(array), new_shape=new_shape, tile_hint=kargs.get('tile_hint'))


def ravel(v):
  """manipulation.py:13-22: flatten to one dimension."""
  v = lazify(v)
  n = 1
  for s in v.shape:
    n *= s
  return reshape(v, (n,))
