"""Extent algebra vs vectors produced by RUNNING the reference's extent.pyx
(oracle/ref_extent/make_extent_vectors.py -> tests/golden/extent_vectors.json).
Checked for both the oracle restatement and the product's C-ABI extent functions."""
import json
import os

import pytest

from spartan_oracle import extent as oex

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'extent_vectors.json')))


def _impls():
  impls = [('oracle', oex)]
  try:
    from spartan_b200.array import extent as pex
    impls.append(('product', pex))
  except ImportError:
    pass
  return impls


def mk(E, t):
  return None if t is None else E.create(t[0], t[1], t[2])


def tup(ex):
  return None if ex is None else [list(ex.ul), list(ex.lr), None if ex.array_shape is None else list(ex.array_shape)]


@pytest.mark.parametrize('name,E', _impls())
def test_golden(name, E):
  for c in GOLD['create_shape']:
    ex = E.create(c['ul'], c['lr'], c['shape'])
    assert (ex is not None) == c['valid']
    if ex is not None:
      assert list(ex.shape) == c['ex_shape']
  for c in GOLD['intersection']:
    assert tup(E.intersection(mk(E, c['a']), mk(E, c['b']))) == c['out'], c
  for c in GOLD['drop_axis']:
    assert tup(E.drop_axis(mk(E, c['a']), c['axis'])) == c['out'], c
  for c in GOLD['ravelled_pos']:
    assert E.ravelled_pos(c['idx'], c['shape']) == c['out']
  for c in GOLD['to_global_axis']:
    assert mk(E, c['a']).to_global(c['idx'], c['axis']) == c['out']
  for c in GOLD['offset_slice']:
    sl = E.offset_slice(mk(E, c['a']), mk(E, c['s']))
    assert [[x.start, x.stop] for x in sl] == c['out']
  for c in GOLD['offset_from']:
    assert tup(E.offset_from(mk(E, c['a']), mk(E, c['s']))) == c['out']
  for c in GOLD['compute_slice']:
    idx = tuple(slice(a, b) for a, b in c['idx'])
    assert tup(E.compute_slice(mk(E, c['a']), idx)) == c['out'], c
  for c in GOLD['from_slice']:
    idx = tuple(slice(a, b) for a, b in c['idx'])
    assert tup(E.from_slice(idx, tuple(c['shape']))) == c['out'], c
  for c in GOLD['find_shape']:
    assert list(E.find_shape([mk(E, t) for t in c['exs']])) == c['out']
  for c in GOLD['change_partition_axis']:
    assert tup(E.change_partition_axis(mk(E, c['a']), c['axis'])) == c['out'], c
