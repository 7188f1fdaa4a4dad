"""Pass criteria shared by the test modules (the role spartan/util.py's Assert plays in the reference's tests)."""
import numpy as np


def all_eq(got, want, tolerance=0):
  """Exact equality (tolerance 0) or |got - want| < tolerance, shapes included when both are arrays."""
  if isinstance(got, np.ndarray) and isinstance(want, np.ndarray):
    assert got.shape == want.shape, 'shapes differ: %s vs %s' % (got.shape, want.shape)
  g, w = np.asarray(got), np.asarray(want)
  if tolerance == 0:
    assert np.array_equal(g, w), 'not equal:\n%s\nvs\n%s' % (g, w)
  else:
    assert np.all(np.abs(g - w) < tolerance), 'not within %g:\n%s\nvs\n%s' % (tolerance, g, w)
