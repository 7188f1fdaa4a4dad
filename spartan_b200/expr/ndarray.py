"""Array creation node (reference: spartan/expr/operator/ndarray.py)."""
import numpy as np

from ..array import distarray
from .base import Expr


class NdArrayExpr(Expr):
  members = ('_shape', 'sparse', 'dtype', 'tile_hint', 'reduce_fn')

  def pretty_str(self):
    return 'DistArray[%d](%s, %s, hint=%s)' % (self.expr_id, self.shape, np.dtype(self.dtype).name, self.tile_hint)

  def visit(self, visitor):
    return self

  def dependencies(self):
    return {}

  def compute_shape(self):
    return tuple(self._shape)

  def _evaluate(self, ctx, deps):
    # ndarray.py:33-41
    return distarray.create(self._shape, self.dtype, reducer=self.reduce_fn, tile_hint=self.tile_hint,
                            sparse=bool(self.sparse))


def ndarray(shape, dtype=np.float64, tile_hint=None, reduce_fn=None, sparse=False):
  """Lazily create a new distributed array (ndarray.py:43-58; default dtype np.float = float64)."""
  return NdArrayExpr(_shape=tuple(shape), dtype=dtype, tile_hint=tile_hint, reduce_fn=reduce_fn, sparse=sparse)
