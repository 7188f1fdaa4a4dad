"""Pins the ORACLE against the reference's own tests for the hot path.

Every test below is the Python-3 restatement of an assertion in /root/reference/tests (file:line cited);
the reference compares against NumPy, so the expected side is recomputed with NumPy here exactly as the
reference test does.  Run on CPU (no GPU, no reference tree needed).
"""
import os
import random

import numpy as np
import pytest

import spartan_oracle
from spartan_oracle import expr, extent, distarray

TEST_SIZE = 50


def all_eq(a, b, tolerance=0):
  # spartan/util.py:236-257 Assert.all_eq
  a = np.asarray(a); b = np.asarray(b)
  assert a.shape == b.shape, (a.shape, b.shape)
  if tolerance == 0:
    assert np.all(a == b)
  else:
    assert np.all(np.abs(a - b) < tolerance)


@pytest.fixture(autouse=True, params=[1, 3, 8])
def ctx(request):
  spartan_oracle.initialize(request.param)
  expr.eval_cache.clear()
  return request.param


# ---- tests/test_extent.py:6-50
def test_intersection():
  a = extent.create((0, 0), (10, 10), None)
  b = extent.create((5, 5), (6, 6), None)
  assert extent.intersection(a, b) == extent.create((5, 5), (6, 6), None)
  assert extent.intersection(b, a) == extent.create((5, 5), (6, 6), None)
  a = extent.create((5, 5), (10, 10), None)
  b = extent.create((4, 6), (6, 8), None)
  assert extent.intersection(a, b) == extent.create((5, 6), (6, 8), None)
  a = extent.create((5, 5), (5, 5), None)
  b = extent.create((1, 1), (2, 2), None)
  assert extent.intersection(a, b) is None


def test_ravelled_pos():
  a = extent.create((2, 2), (7, 7), (10, 10))
  for i in range(10):
    for j in range(10):
      assert extent.ravelled_pos((i, j), a.array_shape) == 10 * i + j
  assert a.to_global(0, axis=None) == 22
  assert a.to_global(10, axis=None) == 42
  assert a.to_global(11, axis=None) == 43
  assert a.to_global(20, axis=None) == 62


def test_unravel():
  rnd = random.Random(0)
  for _ in range(100):
    shp = (20, 77)
    ul = (rnd.randint(0, 19), rnd.randint(0, 76))
    lr = (rnd.randint(ul[0] + 1, 20), rnd.randint(ul[1] + 1, 77))
    a = extent.create(ul, lr, shp)
    assert a.ul == extent.unravelled_pos(a.ravelled_pos(), a.array_shape)


# ---- default tilings, restated from distarray.py:26-110
def test_default_tilings():
  assert distarray.good_tile_shape((4096, 4096), 1) == [4096, 4096]
  assert distarray.good_tile_shape((4096, 4096), 3) == [1365, 4096]
  assert len(distarray.compute_extents((4096, 4096), None, 3)) == 4
  assert distarray.good_tile_shape((4096, 4096), 8) == [512, 4096]
  assert distarray.good_tile_shape((32768, 32768), 8) == [4096, 32768]
  assert distarray.good_tile_shape((50, 50, 50), 3) == [16, 50, 50]
  assert distarray.good_tile_shape((10,), 3) == [3]
  assert list(distarray.compute_extents((10,), None, 3).values()) == [0, 1, 2, 0]
  assert distarray.good_tile_shape((10 ** 7, 256), 8) == [1250000, 256]


# ---- BASELINE config 1 golden value
def test_ones_sum_golden(ctx):
  assert expr.ones((4096, 4096)).sum().glom() == 16777216.0
  assert expr.ones((4096, 4096)).sum().optimized().glom() == 16777216.0


# ---- tests/test_reduce.py:14-107
def test_sum_3d():
  x = expr.arange((TEST_SIZE, TEST_SIZE, TEST_SIZE), dtype=np.int64)
  nx = np.arange(TEST_SIZE ** 3, dtype=np.int64).reshape((TEST_SIZE,) * 3)
  for axis in [None, 0, 1, 2]:
    all_eq(x.sum(axis).glom(), nx.sum(axis))


def test_sum_2d_1d():
  x = expr.arange((TEST_SIZE, TEST_SIZE), dtype=np.int64)
  nx = np.arange(TEST_SIZE * TEST_SIZE, dtype=np.int64).reshape((TEST_SIZE, TEST_SIZE))
  for axis in [None, 0, 1]:
    all_eq(x.sum(axis).glom(), nx.sum(axis))
  all_eq(expr.arange((TEST_SIZE,), dtype=np.int64).sum().glom(), np.arange(TEST_SIZE).sum())


def test_argmin_argmax():
  nx1 = np.arange(TEST_SIZE, dtype=np.int64)
  all_eq(expr.arange((TEST_SIZE,), dtype=np.int64).argmin().glom(), nx1.argmin())
  all_eq(expr.arange((TEST_SIZE,), dtype=np.int64).argmax().glom(), nx1.argmax())
  nx2 = np.arange(TEST_SIZE * TEST_SIZE, dtype=np.int64).reshape((TEST_SIZE, TEST_SIZE))
  x2 = expr.arange((TEST_SIZE, TEST_SIZE), dtype=np.int64)
  all_eq(x2.argmin(axis=1).glom(), nx2.argmin(axis=1))
  all_eq(x2.argmax(axis=1).glom(), nx2.argmax(axis=1))
  nx3 = np.arange(TEST_SIZE ** 3, dtype=np.int64).reshape((TEST_SIZE,) * 3)
  x3 = expr.arange((TEST_SIZE,) * 3, dtype=np.int64)
  for axis in [None, 0, 1, 2]:
    all_eq(x3.argmin(axis).glom(), nx3.argmin(axis))
    all_eq(x3.argmax(axis).glom(), nx3.argmax(axis))


def test_simple_sum():
  for axis in [0, 1, None]:
    a = expr.ones((TEST_SIZE, TEST_SIZE)) + expr.ones((TEST_SIZE, TEST_SIZE))
    all_eq(a.sum(axis=axis).glom(), 2 * np.ones((TEST_SIZE, TEST_SIZE)).sum(axis))


def test_count_nonzero_zero():
  assert expr.count_nonzero(expr.ones((TEST_SIZE,))).glom() == TEST_SIZE
  assert expr.count_nonzero(expr.zeros((TEST_SIZE,))).glom() == 0
  assert expr.count_zero(expr.ones((TEST_SIZE,))).glom() == 0
  assert expr.count_zero(expr.zeros((TEST_SIZE,))).glom() == TEST_SIZE


# ---- tests/test_maptiles.py:12-62, tests/test_elementwise.py:9-22
def test_map_chains():
  all_eq((expr.ones((20, 20)) + expr.ones((20, 20))).glom(), 2 * np.ones((20, 20)))
  a = expr.ones((10, 10)); b = expr.ones((10, 10)); c = expr.ones((10, 10))
  all_eq((a + b + c).glom(), np.ones((10, 10)) * 3)
  all_eq((a + b + a + b + a + b + a + b + a + b).glom(), np.ones((10, 10)) * 10)
  all_eq((a + b + a + b + a + b + a + b + a + b).optimized().glom(), np.ones((10, 10)) * 10)


def test_ln():
  a = 1.0 + expr.ones((100,), dtype=np.float32)
  b = 1.0 + np.ones(100).astype(np.float32)
  got = expr.ln(a).glom()
  assert got.dtype == np.float32          # NumPy-1 value-based casting: python float does not widen
  assert np.allclose(got, np.log(b))


def test_broadcast():
  a = expr.ones((2, 1)); b = expr.ones((2, 5))
  all_eq((a / b).glom(), np.ones((2, 5)))
  all_eq((b / a).glom(), np.ones((2, 5)))


def test_maximum():
  rng = np.random.RandomState(0)
  np_a = rng.randn(10, 10); np_b = rng.randn(10, 10)
  sp_a = expr.from_numpy(np_a); sp_b = expr.from_numpy(np_b)
  all_eq(expr.maximum(sp_a, sp_b).glom(), np.maximum(np_a, np_b))
  all_eq(expr.maximum(sp_a, 0).glom(), np.maximum(np_a, 0))


# ---- tests/test_dot.py:8-103, tests/test_matmul.py:12-22
def test_dot_2d_2d():
  all_eq(expr.dot(expr.arange((132, 100)), expr.arange((100, 77))).glom(),
         np.dot(np.arange(13200).reshape(132, 100), np.arange(7700).reshape(100, 77)))
  all_eq(expr.dot(expr.arange((67, 100)), expr.arange((100, 77))).glom(),
         np.dot(np.arange(6700).reshape(67, 100), np.arange(7700).reshape(100, 77)))
  all_eq(expr.dot(expr.arange((77, 100)), np.arange(8800).reshape(100, 88)).glom(),
         np.dot(np.arange(7700).reshape(77, 100), np.arange(8800).reshape(100, 88)))


def test_dot_vec():
  all_eq(expr.dot(expr.arange(stop=100), expr.arange(stop=100)).glom(), [np.dot(np.arange(100), np.arange(100))])
  all_eq(expr.dot(expr.arange((100, 77)), expr.arange(stop=77)).glom(),
         np.dot(np.arange(7700).reshape(100, 77), np.arange(77)))
  all_eq(expr.dot(expr.arange((77, 100)), expr.arange(stop=100)).glom(),
         np.dot(np.arange(7700).reshape(77, 100), np.arange(100)))
  all_eq(expr.dot(expr.arange(stop=100), np.arange(100)).glom(), [np.dot(np.arange(100), np.arange(100))])
  all_eq(expr.dot(expr.arange((77, 100)), np.arange(100)).glom(),
         np.dot(np.arange(7700).reshape(77, 100), np.arange(100)))


def test_matmul():
  x = expr.arange((100, 50), dtype=np.int64).astype(np.float64)
  y = expr.arange((50, 100), dtype=np.int64).astype(np.float64)
  nx = np.arange(5000, dtype=np.int64).reshape(100, 50).astype(np.float64)
  ny = np.arange(5000, dtype=np.int64).reshape(50, 100).astype(np.float64)
  all_eq(expr.dot(x, y).glom(), np.dot(nx, ny))


# ---- tests/test_creation.py:17-68 (arange)
def test_arange():
  with pytest.raises(ValueError):
    expr.arange()
  all_eq(expr.arange((10,)).glom(), np.arange(10))
  all_eq(expr.arange((3, 5)).glom(), np.arange(15).reshape((3, 5)))
  all_eq(expr.arange((10,), -1).glom(), np.arange(-1, 9))
  all_eq(expr.arange((3, 5), -1).glom(), np.arange(-1, 14).reshape((3, 5)))
  all_eq(expr.arange((10,), step=2).glom(), np.arange(0, 20, 2))
  all_eq(expr.arange((3, 5), 1, step=2).glom(), np.arange(1, 31, 2).reshape((3, 5)))
  all_eq(expr.arange(stop=10).glom(), np.arange(10))
  all_eq(expr.arange(-1, 19, 2).glom(), np.arange(-1, 19, 2))


# ---- tests/test_statistics.py:16-30, tests/test_mathematics.py:9-15, tests/test_logic.py:9-23
def test_min_max_prod_logic():
  src = np.asarray([1, 1, 1, 2, 2, 5, 5, 10])
  all_eq(expr.max(expr.from_numpy(src)).glom(), np.max(src))
  all_eq(expr.min(expr.from_numpy(src)).glom(), np.min(src))
  src = np.arange(100).reshape(10, 10)
  all_eq(expr.min(expr.from_numpy(src), axis=1).glom(), np.min(src, axis=1))
  nA = np.arange(40000, dtype=np.int32).reshape(100, 400)
  A = expr.from_numpy(nA)
  got = A.prod().glom()
  assert got.dtype == np.int64
  all_eq(got, nA.astype(np.int64).prod())
  nC = nA.T.copy() // 1000
  all_eq(expr.all(expr.from_numpy(nC)).glom(), np.all(nC))
  all_eq(expr.any(expr.from_numpy(nC)).glom(), np.any(nC))


# ---- tests/test_optimization.py:124-163 (slices dropped: views are out of scope), tolerance 1e-6
def test_optimization_reduced():
  rng = np.random.RandomState(1)
  na = rng.rand(300, 300); nb = rng.rand(300, 300)
  a = expr.from_numpy(na); b = expr.from_numpy(nb)
  c = a - b; d = a + c; h = c - d; i = c + h
  m = h + i; n = i - m; o = n - m; q = n + o; r = q - m
  s = expr.sum(r)
  nc = na - nb; nd = na + nc; nh = nc - nd; ni = nc + nh
  nm = nh + ni; nn = ni - nm; no = nn - nm; nq = nn + no; nr = nq - nm
  opt = s.optimized()
  assert isinstance(opt, expr.ReduceExpr)
  assert all(not isinstance(ch, expr.MapExpr) for ch in opt.children)   # fully fused into the reduce
  all_eq(np.sum(nr), opt.glom(), tolerance=1e-6 * abs(np.sum(nr)) + 1e-6)


def test_fusion_structure():
  # optimize.py:133-227: map chain collapses to one MapExpr over the leaf arrays
  x = expr.ones((8, 8)); y = expr.ones((8, 8))
  e = (x * 2 + y).sum(axis=0)
  opt = e.optimized()
  assert isinstance(opt, expr.ReduceExpr)
  assert len(opt.op.deps) == 2 and isinstance(opt.op.deps[1], expr.LocalMapExpr)
  all_eq(opt.glom(), e.glom())
  all_eq(opt.glom(), np.full((8,), 24, np.float32))
  assert opt.glom().dtype == np.float32


# ---- tests/test_slice.py:25-80 (the shuffle variant needs the shuffle operator, outside the hot path)
def test_slice_reference_cases():
  T = 10
  x = expr.arange((T, T)); nx = np.arange(T * T).reshape(T, T)
  all_eq(x[5:8, 5:8].evaluate().glom(), nx[5:8, 5:8])                       # test_slice_get
  all_eq(expr.map(x[5:8, 5:8], lambda tile: tile + 1).glom(), nx[5:8, 5:8] + 1)   # test_slice_map
  x3 = expr.arange((10, 10, 10), dtype=np.int64); n3 = np.arange(1000).reshape((10, 10, 10))
  all_eq(expr.map(x3[:, :, 0], lambda tile: tile + 13).glom().reshape(10, 10), n3[:, :, 0] + 13)   # test_slice_map2
  xr = expr.arange((T, T, T), dtype=np.int64); nr = np.arange(T ** 3).reshape((T, T, T))
  all_eq(xr[:, :, 0].sum().glom(), nr[:, :, 0].sum())                       # test_slice_reduce
  a = expr.arange((T,), dtype=np.int64); na = np.arange(T)
  all_eq((a[1:] - a[:-1]).glom(), na[1:] - na[:-1])                         # test_slice_sub
  all_eq((a[1:] - a[:-1]).optimized().glom(), na[1:] - na[:-1])
  assert extent.from_slice((slice(None), slice(None), 0), [100, 100, 100]).shape == (100, 100, 1)   # test_from_slice


# ---- tests/test_transpose.py:9-37 (dense cases)
def test_transpose_reference_cases():
  all_eq(expr.transpose(expr.arange((372, 134))).glom(), np.transpose(np.arange(372 * 134).reshape(372, 134)))
  all_eq(expr.transpose(expr.arange((31, 32, 33))).glom(), np.transpose(np.arange(31 * 32 * 33).reshape(31, 32, 33)))
  rng = np.random.RandomState(0)
  n1 = rng.random_sample((401, 97)); n2 = rng.random_sample((401, 97))
  got = expr.dot(expr.from_numpy(n1), expr.transpose(expr.from_numpy(n2))).glom()
  assert np.all(np.isclose(np.dot(n1, np.transpose(n2)), got))


# ---- tests/test_reshape.py:9-88,98-121 (dense cases)
def test_reshape_reference_cases():
  all_eq(expr.reshape(expr.arange((10, 10)), (100,)).glom(), expr.arange((100,)).glom())          # reshape1
  b = expr.reshape(expr.arange((1000,), tile_hint=[100]), (10, 100)).evaluate()                   # reshape2
  expr.reshape(b, (1000,)).evaluate()
  d = expr.reshape(expr.reshape(expr.reshape(expr.arange((100, 100)), (10000,)), (10000, 1)), (1, 10000))
  all_eq(d.glom(), expr.arange((1, 10000)).glom())                                                  # reshape3
  f = expr.arange((10000,))
  for shp in ((10, 1000), (1000, 10), (20, 500), (500, 20), (1, 10000)):
    f = expr.reshape(f, shp)
  all_eq(f.glom(), expr.arange((1, 10000)).glom())                                                  # reshape4
  for n, s1, s2 in ((35511, (133, 267), (267, 133)), (12319, (127, 97), (97, 127))):               # reshape5, 6
    d = expr.reshape(expr.reshape(expr.reshape(expr.arange((n,)), s1), s2), (1, n))
    all_eq(d.glom(), expr.arange((1, n)).glom())
  targets = [(23, 120, 100), (12, 230, 100), (276000, 1), (1, 276000)]                             # reshape7
  for src in ((100, 23, 120), (12, 23, 1000), (1, 276000), (276000, 1), (276000,)):
    a = expr.arange(src)
    for shp in targets:
      all_eq(expr.reshape(a, shp).glom(), np.arange(276000).reshape(shp))
  rng = np.random.RandomState(1)                                                                    # reshape_dot
  n1 = rng.random_sample((357, 93)); n2 = rng.random_sample((31, 357))
  all_eq(np.dot(np.reshape(n1, (1071, 31)), n2),
         expr.dot(expr.reshape(expr.from_numpy(n1), (1071, 31)), expr.from_numpy(n2)).glom(), 10e-9)
  n1 = rng.random_sample((357, 718)); n2 = rng.random_sample((718,))
  all_eq(np.dot(n1, np.reshape(n2, (718, 1))),
         expr.dot(expr.from_numpy(n1), expr.reshape(expr.from_numpy(n2), (718, 1))).glom(), 10e-9)
  n1 = rng.random_sample((718,)); n2 = rng.random_sample((1, 357))
  all_eq(np.dot(np.reshape(n1, (718, 1)), n2),
         expr.dot(expr.reshape(expr.from_numpy(n1), (718, 1)), expr.from_numpy(n2)).glom(), 10e-9)


# ---- tests/test_fio.py:24-31 (dense save / load round trips, zipped and not)
def test_fio_reference_roundtrip(tmp_path):
  from spartan_oracle import fio
  t1 = expr.arange((123, 45), dtype=np.float64) if False else expr.from_numpy(np.random.RandomState(0).rand(123, 45))
  for iszip in (False, True):
    assert fio.save(t1, 'fiotest1', str(tmp_path), iszip) is True
    all_eq(t1.glom(), fio.load('fiotest1', str(tmp_path), iszip).glom())


# ---- tests/test_creation.py:10-16 (eye / identity), :65-92 (diagonal / diag)
def test_creation_eye_diag_reference_cases():
  all_eq(expr.eye(100, 10).glom(), np.eye(100, 10))
  all_eq(expr.identity(100).glom(), np.identity(100))
  rng = np.random.RandomState(2)
  for shp in ((2, 2), (15, 10), (16, 16)):
    x = rng.randn(*shp)
    all_eq(expr.diagonal(expr.from_numpy(x)).glom(), np.diagonal(x))
  dim = random.randint(1, 99)
  x = rng.randn(dim, dim)
  all_eq(expr.diag(expr.from_numpy(x)).glom(), np.diag(x))
  all_eq(expr.diag(expr.diag(expr.from_numpy(x))).glom(), np.diag(np.diag(x)))


# ---- tests/test_statistics.py:32-70 (std; Assert.float_close / all_close)
def test_std_reference_cases():
  rng = np.random.RandomState(4)
  for shp in ((10,), (10, 10), (17, 17)):
    x = rng.randn(*shp)
    assert abs(expr.std(expr.from_numpy(x)).glom() - np.std(x)) < 1e-6
  for shp in ((10, 10), (15, 13), (13, 15), (17, 17)):
    x = rng.randn(*shp)
    sx = expr.from_numpy(x)
    assert np.allclose(expr.std(sx, 0).glom(), np.std(x, 0))
    assert np.allclose(expr.std(sx, 1).glom(), np.std(x, 1))


# ---- k-means and sparse dot: vectors produced by running the reference's own mapper functions
#      (oracle/ref_apps/make_app_vectors.py executes the definitions of k_means_.py:61-97 and dot.py:195-217 taken
#      from the reference's files; their inputs and outputs are committed as tests/golden/app_vectors.json)
def _app_vectors():
  import json
  with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'app_vectors.json')) as f:
    return json.load(f)


def test_kmeans_mappers_match_reference_vectors():
  from spartan_oracle import apps
  vec = _app_vectors()
  assert vec['kmeans'], 'no k-means vectors'
  for case in vec['kmeans']:
    pts, centers, k = np.array(case['points']), np.array(case['centers']), case['k']
    for t in case['tiles']:
      r0, r1 = t['rows']
      labels = apps.kmeans_dist_mapper(pts[r0:r1], centers)
      assert labels.tolist() == t['labels']
      assert apps.kmeans_count_mapper(labels, k).tolist() == t['counts']
      assert np.array_equal(apps.kmeans_center_mapper(pts[r0:r1], labels, k), np.array(t['sums']))
      assert t['label_extent'] == [[r0], [r1], [case['n']]]
    # one full iteration of the driver (k_means_.py:130-160) = the per-tile results summed (Q7) and divided
    new_centers, labels = apps.kmeans_fit(pts, centers, 1, case['tile_rows'])
    counts = np.sum([t['counts'] for t in case['tiles']], axis=0).astype(np.float64)
    sums = np.sum([t['sums'] for t in case['tiles']], axis=0)
    assert labels.tolist() == [v for t in case['tiles'] for v in t['labels']]
    ok = counts > 0
    assert np.allclose(new_centers[ok], sums[ok] / counts[ok, None], rtol=1e-15, atol=0)


def test_sparse_dot_mapper_matches_reference_vectors():
  import scipy.sparse
  from spartan_oracle import apps
  vec = _app_vectors()
  assert vec['spmv'], 'no sparse-dot vectors'
  for case in vec['spmv']:
    n, strip = case['n'], case['strip']
    m = scipy.sparse.coo_matrix((np.array(case['vals'], np.float32), (case['rows'], case['cols'])), shape=(n, n))
    x = np.array(case['x'], np.float32)
    want = np.zeros(n, np.float32)
    for s in case['strips']:
      c0, c1 = s['cols']
      assert s['target_extent'] == [[0, 0], [n, 1], [n, 1]] and s['partial_dtype'] == 'float32'
      part = scipy.sparse.csc_matrix(m)[:, c0:c1].tocsr().dot(x[c0:c1]).astype(np.float32)
      assert np.array_equal(part, np.array(s['partial'], np.float32))     # the oracle's per-strip product, bit for bit
      want = np.add(want, np.array(s['partial'], np.float32))              # np.add merge of the strips (tile.pyx:263-268)
    assert np.array_equal(apps.spmv_strips(m, x, strip), want)
