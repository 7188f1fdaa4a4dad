"""Transpose operation and expr (reference: spartan/expr/operator/transpose.py:70-100)."""
from ..array.views import Transpose
from .base import Expr, lazify


class TransposeExpr(Expr):
  members = ('array', 'tile_hint')

  def __str__(self):
    return 'Transpose[%d] %s' % (self.expr_id, self.array)

  def dependencies(self):
    return {'array': self.array}

  def visit(self, visitor):
    return TransposeExpr(array=visitor.visit(self.array), tile_hint=self.tile_hint, expr_id=self.expr_id,
                         shape_cache=self.shape_cache)

  def _evaluate(self, ctx, deps):
    return Transpose(deps['array'])

  def compute_shape(self):
    return tuple(self.array.shape[::-1])


def transpose(array, tile_hint=None):
  """Transpose ``array`` (all axes reversed, like np.transpose without ``axes``); transpose.py:86-100."""
  return TransposeExpr(array=lazify(array), tile_hint=tile_hint)
