"""Generates tests/golden/extent_vectors.json by RUNNING the reference's extent.pyx (built by
build_ref_extent.py) on seeded random inputs.

Only functions whose arithmetic is division-free are recorded: Cython compiles the Python-2 source for a
Python-3 runtime with ``/`` on Python ints as true division, so ``unravelled_pos``, ``to_global(axis=None)``
and ``find_rect`` do not behave as they did under Python 2 and are pinned by the reference's own test
values instead (tests/test_oracle_reference_vectors.py).

  python oracle/ref_extent/build_ref_extent.py && python oracle/ref_extent/make_extent_vectors.py
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'stub'))
sys.path.insert(0, os.path.join(HERE, '..', '_ref'))
import ref_extent as R  # noqa: E402


def tup(ex):
  return None if ex is None else [list(ex.ul), list(ex.lr), None if ex.array_shape is None else list(ex.array_shape)]


def rand_extent(rnd, shape, allow_empty=False):
  ul, lr = [], []
  for d in shape:
    a = rnd.randint(0, d - 1)
    b = rnd.randint(a if allow_empty else a + 1, d)
    ul.append(a); lr.append(b)
  return ul, lr


def main():
  rnd = random.Random(20260925)
  out = {'intersection': [], 'create_shape': [], 'drop_axis': [], 'offset_slice': [], 'offset_from': [],
         'compute_slice': [], 'from_slice': [], 'ravelled_pos': [], 'change_partition_axis': [],
         'find_shape': [], 'to_global_axis': []}
  for _ in range(300):
    nd = rnd.randint(1, 4)
    shape = [rnd.randint(1, 12) for _ in range(nd)]
    a_ul, a_lr = rand_extent(rnd, shape, allow_empty=True)
    b_ul, b_lr = rand_extent(rnd, shape, allow_empty=True)
    a = R.create(a_ul, a_lr, shape); b = R.create(b_ul, b_lr, shape)
    out['create_shape'].append({'ul': a_ul, 'lr': a_lr, 'shape': shape,
                                'valid': a is not None, 'ex_shape': None if a is None else list(a.shape)})
    if a is not None and b is not None:
      out['intersection'].append({'a': tup(a), 'b': tup(b), 'out': tup(R.intersection(a, b))})
    if a is not None:
      for axis in list(range(-nd, nd)) + [None]:
        out['drop_axis'].append({'a': tup(a), 'axis': axis, 'out': tup(R.drop_axis(a, axis))})
      out['ravelled_pos'].append({'idx': a_ul, 'shape': shape, 'out': int(R.ravelled_pos(a_ul, shape))})
      ax = rnd.randint(0, nd - 1)
      idx = rnd.randint(0, 5)
      out['to_global_axis'].append({'a': tup(a), 'idx': idx, 'axis': ax, 'out': int(a.to_global(idx, ax))})
      # sub-extent inside a
      s_ul = [rnd.randint(u, l - 1) for u, l in zip(a.ul, a.lr)]
      s_lr = [rnd.randint(u + 1, l) for u, l in zip(s_ul, a.lr)]
      s = R.create(s_ul, s_lr, shape)
      sl = R.offset_slice(a, s)
      out['offset_slice'].append({'a': tup(a), 's': tup(s), 'out': [[x.start, x.stop] for x in sl]})
      out['offset_from'].append({'a': tup(a), 's': tup(s), 'out': tup(R.offset_from(a, s))})
      idx = tuple(slice(rnd.randint(0, 3), rnd.randint(1, 6)) for _ in range(rnd.randint(1, nd)))
      out['compute_slice'].append({'a': tup(a), 'idx': [[x.start, x.stop] for x in idx],
                                   'out': tup(R.compute_slice(a, idx))})
    idx = tuple(slice(rnd.randint(0, 2), rnd.randint(1, 12)) for _ in range(nd))
    try:
      res = tup(R.from_slice(idx, tuple(shape)))
      out['from_slice'].append({'idx': [[x.start, x.stop] for x in idx], 'shape': shape, 'out': res})
    except AssertionError:
      pass
    exs = [x for x in (a, b) if x is not None]
    if exs:
      out['find_shape'].append({'exs': [tup(x) for x in exs], 'out': [int(v) for v in R.find_shape(exs)]})
  # change_partition_axis: row strips -> column strips and back, vectors, already-aligned
  for _ in range(200):
    rows, cols = rnd.randint(1, 40), rnd.randint(1, 40)
    shape = (rows, cols)
    if rnd.random() < 0.5:
      step = rnd.randint(1, rows); r0 = rnd.randrange(0, rows, step)
      ex = R.create((r0, 0), (min(rows, r0 + step), cols), shape)
    else:
      step = rnd.randint(1, cols); c0 = rnd.randrange(0, cols, step)
      ex = R.create((0, c0), (rows, min(cols, c0 + step)), shape)
    for axis in (0, 1, -1):
      out['change_partition_axis'].append({'a': tup(ex), 'axis': axis, 'out': tup(R.change_partition_axis(ex, axis))})
  for _ in range(30):
    n = rnd.randint(1, 30); step = rnd.randint(1, n); s0 = rnd.randrange(0, n, step)
    ex = R.create((s0,), (min(n, s0 + step),), (n,))
    for axis in (0, 1):
      out['change_partition_axis'].append({'a': tup(ex), 'axis': axis, 'out': tup(R.change_partition_axis(ex, axis))})
  path = os.path.join(HERE, '..', '..', 'tests', 'golden', 'extent_vectors.json')
  with open(path, 'w') as f:
    json.dump(out, f, separators=(',', ':'))
  print({k: len(v) for k, v in out.items()}, os.path.getsize(path))


if __name__ == '__main__':
  main()
