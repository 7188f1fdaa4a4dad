"""Helpers that define test pass criteria and a few shared utilities.
Mirrors the parts of spartan/util.py the hot path uses (Assert :222-326, divup :404-408, is_iterable)."""
import collections.abc
import math

import numpy as np


def divup(a, b):
  return int(math.ceil(float(a) / b))          # util.py:404-408


def is_iterable(x):
  return isinstance(x, collections.abc.Iterable) and not isinstance(x, (str, bytes, np.ndarray))


class Assert(object):
  """Assertion helpers with the reference's semantics (util.py:222-326)."""

  @staticmethod
  def all_eq(a, b, tolerance=0):
    if isinstance(a, np.ndarray) and isinstance(b, np.ndarray):
      assert a.shape == b.shape, 'Mismatched shapes: %s %s' % (a.shape, b.shape)
      if tolerance == 0:
        assert np.all(a == b), 'Failed: \n%s\n ==\n%s' % (a, b)
      else:
        assert np.all(np.abs(a - b) < tolerance), 'Failed: \n%s\n ==\n%s' % (a, b)
      return
    if np.isscalar(a) or np.isscalar(b) or np.ndim(a) == 0 or np.ndim(b) == 0:
      if tolerance == 0:
        assert a == b, 'Failed: \n%s\n ==\n%s' % (a, b)
      else:
        assert abs(a - b) < tolerance, 'Failed: \n%s\n ==\n%s' % (a, b)
      return
    for i, j in zip(a, b):
      assert i == j, 'Failed: \n%s\n ==\n%s' % (a, b)

  @staticmethod
  def all_close(a, b):
    assert isinstance(a, np.ndarray) and isinstance(b, np.ndarray)
    assert a.shape == b.shape, 'Mismatched shapes: %s %s' % (a.shape, b.shape)
    assert np.allclose(a, b), 'Failed: \n%s close to \n%s' % (a, b)

  @staticmethod
  def float_close(a, b):
    Assert.all_close(np.array(a), np.array(b))

  @staticmethod
  def eq(a, b, fmt='', *args):
    assert a == b, 'Failed: %s == %s (%s)' % (a, b, fmt % args if args else fmt)

  @staticmethod
  def ne(a, b, fmt='', *args):
    assert a != b, 'Failed: %s != %s (%s)' % (a, b, fmt % args if args else fmt)

  @staticmethod
  def le(a, b, fmt='', *args):
    assert a <= b, 'Failed: %s <= %s' % (a, b)

  @staticmethod
  def true(expr):
    assert expr, 'Failed: %s == True' % (expr,)

  @staticmethod
  def isinstance(expr, klass):
    assert isinstance(expr, klass), 'Failed: isinstance(%s, %s) [type = %s]' % (expr, klass, type(expr))

  @staticmethod
  def not_null(expr):
    assert expr is not None, expr

  @staticmethod
  def no_duplicates(collection):
    d = collections.defaultdict(int)
    for item in collection:
      d[item] += 1
    bad = [(k, v) for k, v in d.items() if v > 1]
    assert len(bad) == 0, 'Duplicates found: %s' % bad

  @staticmethod
  def raises_exception(exception, function, *args, **kwargs):
    try:
      function(*args, **kwargs)
    except exception:
      return
    assert False, '%s expected, no error was raised.' % exception.__name__
