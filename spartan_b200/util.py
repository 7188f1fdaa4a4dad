"""Small shared helpers of the host layer: integer ceil-division with the reference's float semantics
(spartan/util.py:404-408, needed bit-exactly by change_partition_axis) and argument checks that raise ordinary
Python exceptions.  Test pass criteria live under tests/, not here."""
import collections.abc
import math

import numpy as np


def divup(a, b):
  return int(math.ceil(float(a) / b))          # util.py:404-408: float ceil, exact below 2**53


def is_iterable(x):
  return isinstance(x, collections.abc.Iterable) and not isinstance(x, (str, bytes, np.ndarray))


def require_type(value, types, what='argument'):
  if not isinstance(value, types):
    raise TypeError('%s must be %s, got %s' % (what, types, type(value).__name__))
  return value


def require_equal(a, b, what=''):
  if a != b:
    raise ValueError('%s: %s != %s' % (what or 'mismatch', a, b))


def require_unique(items, what='items'):
  seen = set()
  dup = [x for x in items if x in seen or seen.add(x)]
  if dup:
    raise ValueError('duplicate %s: %s' % (what, dup))
