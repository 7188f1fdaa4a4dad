"""from_numpy: the only way to feed deterministic inputs (reference:
spartan/expr/operator/write_array.py:46-94 WriteArrayExpr restricted to the from_numpy use, :424-445)."""
import numpy as np

from ..array import distarray, extent
from .base import Expr


class WriteArrayExpr(Expr):
  members = ('npa', 'tile_hint')

  def compute_shape(self):
    return self.npa.shape

  def visit(self, visitor):
    return self

  def dependencies(self):
    return {}

  def _evaluate(self, ctx, deps):
    arr = distarray.create(self.npa.shape, self.npa.dtype, tile_hint=self.tile_hint)
    # every rank holds npa and uploads (H2D) exactly the parts that land in its own tiles
    arr.update(extent.from_shape(self.npa.shape) if self.npa.ndim else extent.create((), (), ()), self.npa)
    return arr


def from_numpy(npa, tile_hint=None):
  """Make a distributed array from a numpy array (write_array.py:424-445)."""
  if not isinstance(npa, np.ndarray):
    raise TypeError('Expected ndarray, got: %s' % type(npa))
  return WriteArrayExpr(npa=npa, tile_hint=tile_hint)
