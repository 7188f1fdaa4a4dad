"""GPU parity: the CUDA path (through the Python host -> C ABI -> sm_100a kernels) against the ORACLE on
the same seeded inputs.  These are the reference's own hot-path tests (file:line cited) re-expressed for
pytest, plus edge cases.  Integer / index work is bit-exact; float tolerances are written in each test.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import spartan_oracle
from spartan_oracle import expr as oexpr

import spartan_b200 as sp
from helpers import all_eq

TEST_SIZE = 50


@pytest.fixture(autouse=True)
def ctx():
  import torch
  if not torch.cuda.is_available():
    pytest.skip('needs a CUDA device')
  c = sp.initialize()
  spartan_oracle.initialize(c.num_workers)
  oexpr.eval_cache.clear()
  yield c


def both(build):
  """Evaluates the same expression builder with the product and with the oracle."""
  return build(sp).glom(), build(oexpr).glom()


# ------------------------------------------------------------------ BASELINE config 1 golden value
def test_ones_sum_golden():
  assert sp.ones((4096, 4096)).sum().glom() == 16777216.0
  assert sp.ones((4096, 4096)).sum().optimized().glom() == 16777216.0


# ------------------------------------------------------------------ tests/test_maptiles.py:12-62
def test_map_chains():
  all_eq((sp.ones((20, 20)) + sp.ones((20, 20))).glom(), 2 * np.ones((20, 20)))
  a = sp.ones((10, 10)); b = sp.ones((10, 10)); c = sp.ones((10, 10))
  all_eq((a + b + c).glom(), np.ones((10, 10)) * 3)
  many = (a + b + a + b + a + b + a + b + a + b)
  all_eq(many.glom(), np.ones((10, 10)) * 10)
  all_eq(many.optimized().glom(), np.ones((10, 10)) * 10)
  assert many.glom().dtype == np.float32


def test_ln():
  a = 1.0 + sp.ones((100,), dtype=np.float32)
  got = sp.ln(a).glom()
  assert got.dtype == np.float32
  np.testing.assert_allclose(got, np.log(1.0 + np.ones(100, np.float32)), rtol=1e-6)   # logf <= 1 ulp


def test_broadcast():
  a = sp.ones((2, 1)); b = sp.ones((2, 5))
  all_eq((a / b).glom(), np.ones((2, 5)))
  all_eq((b / a).glom(), np.ones((2, 5)))


def test_maximum_and_scalar_broadcast():
  # tests/test_elementwise.py:9-22 -- exact (max is exact in any precision)
  rng = np.random.RandomState(0)
  np_a = rng.randn(10, 10); np_b = rng.randn(10, 10)
  all_eq(sp.maximum(sp.from_numpy(np_a), sp.from_numpy(np_b)).glom(), np.maximum(np_a, np_b))
  all_eq(sp.maximum(sp.from_numpy(np_a), 0).glom(), np.maximum(np_a, 0))


@pytest.mark.parametrize('shape', [(1,), (7,), (33, 5), (128, 257), (3, 4, 5), (1000, 1)])
@pytest.mark.parametrize('dtype', [np.float32, np.float64, np.int32, np.int64])
def test_elementwise_bit_exact(shape, dtype):
  """+, -, * chains round once per op exactly like the NumPy ufunc chain (-fmad=false): bit-exact."""
  rng = np.random.RandomState(1)
  if np.dtype(dtype).kind == 'f':
    x = rng.randn(*shape).astype(dtype); y = rng.randn(*shape).astype(dtype)
  else:
    x = rng.randint(-1000, 1000, size=shape).astype(dtype); y = rng.randint(1, 1000, size=shape).astype(dtype)
  for build in (lambda m, X, Y: X * 2 + Y, lambda m, X, Y: (X - Y) * (X + Y) - X, lambda m, X, Y: m.maximum(X, Y) - m.minimum(X, 3),
                lambda m, X, Y: m.abs(X) + m.square(Y)):
    got = build(sp, sp.from_numpy(x), sp.from_numpy(y)).optimized().glom()
    ref = build(oexpr, oexpr.from_numpy(x), oexpr.from_numpy(y)).optimized().glom()
    assert got.dtype == ref.dtype, (got.dtype, ref.dtype)
    all_eq(got, ref)


def test_compare_logic_astype():
  rng = np.random.RandomState(2)
  x = rng.randn(64, 33).astype(np.float32); y = rng.randn(64, 33).astype(np.float32)
  for build in (lambda m, X, Y: X > Y, lambda m, X, Y: m.logical_and(X > 0, Y < 0), lambda m, X, Y: (X == X) | (Y != Y),
                lambda m, X, Y: m.astype(X * 10, np.int64), lambda m, X, Y: m.astype(X > 0, np.float32) * Y):
    got = build(sp, sp.from_numpy(x), sp.from_numpy(y)).glom()
    ref = build(oexpr, oexpr.from_numpy(x), oexpr.from_numpy(y)).glom()
    assert got.dtype == ref.dtype
    all_eq(got, ref)


def test_transcendentals_tolerance():
  """exp / log / sqrt / divide: device libm is <= 2 ulp; tolerance 4 ulp of float32 (rtol 5e-7)."""
  rng = np.random.RandomState(3)
  x = (rng.rand(1000).astype(np.float32) + 0.1)
  for build in (lambda m, X: m.exp(X), lambda m, X: m.log(X), lambda m, X: m.sqrt(X), lambda m, X: 1.0 / X,
                lambda m, X: m.power(X, 2.5)):
    got = build(sp, sp.from_numpy(x)).glom(); ref = build(oexpr, oexpr.from_numpy(x)).glom()
    assert got.dtype == ref.dtype == np.float32
    np.testing.assert_allclose(got, ref, rtol=5e-7)


def test_row_and_column_vector_broadcast():
  rng = np.random.RandomState(4)
  x = rng.randn(96, 40).astype(np.float32); r = rng.randn(1, 40).astype(np.float32); c = rng.randn(96, 1).astype(np.float32)
  got = (sp.from_numpy(x) * sp.from_numpy(r) + sp.from_numpy(c)).optimized().glom()
  all_eq(got, x * r + c)


# ------------------------------------------------------------------ tests/test_reduce.py:14-107
def test_sum_3d_int64_exact():
  nx = np.arange(TEST_SIZE ** 3, dtype=np.int64).reshape((TEST_SIZE,) * 3)
  for axis in [None, 0, 1, 2]:
    x = sp.arange((TEST_SIZE, TEST_SIZE, TEST_SIZE), dtype=np.int64)
    all_eq(x.sum(axis).glom(), nx.sum(axis))


def test_sum_2d_1d():
  nx = np.arange(TEST_SIZE * TEST_SIZE, dtype=np.int64).reshape((TEST_SIZE, TEST_SIZE))
  for axis in [None, 0, 1]:
    all_eq(sp.arange((TEST_SIZE, TEST_SIZE), dtype=np.int64).sum(axis).glom(), nx.sum(axis))
  all_eq(sp.arange((TEST_SIZE,), dtype=np.int64).sum().glom(), np.arange(TEST_SIZE).sum())


def test_simple_sum():
  for axis in [0, 1, None]:
    a = sp.ones((TEST_SIZE, TEST_SIZE)) + sp.ones((TEST_SIZE, TEST_SIZE))
    all_eq(a.sum(axis=axis).glom(), 2 * np.ones((TEST_SIZE, TEST_SIZE)).sum(axis))


def test_count_nonzero_zero():
  assert sp.count_nonzero(sp.ones((TEST_SIZE,))).glom() == TEST_SIZE
  assert sp.count_nonzero(sp.zeros((TEST_SIZE,))).glom() == 0
  assert sp.count_zero(sp.ones((TEST_SIZE,))).glom() == 0
  assert sp.count_zero(sp.zeros((TEST_SIZE,))).glom() == TEST_SIZE
  assert sp.count_nonzero(sp.ones((5000, 5000))).glom() == 25000000     # > 2^24: must not saturate


@pytest.mark.parametrize('shape,axis', [((257, 129), 0), ((257, 129), 1), ((257, 129), None), ((5, 6, 7), 1),
                                        ((100000,), None), ((3, 100000), 1), ((100000, 3), 0), ((1, 1), None)])
def test_reductions_vs_oracle(shape, axis):
  rng = np.random.RandomState(5)
  xi = rng.randint(-50, 50, size=shape).astype(np.int64)
  xf = rng.rand(*shape).astype(np.float32)
  for name in ('sum', 'min', 'max'):
    got = getattr(sp, name)(sp.from_numpy(xi), axis).glom(); ref = getattr(oexpr, name)(oexpr.from_numpy(xi), axis).glom()
    assert got.dtype == ref.dtype
    all_eq(np.asarray(got), np.asarray(ref))                       # integers: bit-exact
    got = getattr(sp, name)(sp.from_numpy(xf), axis).glom(); ref = getattr(oexpr, name)(oexpr.from_numpy(xf), axis).glom()
    assert got.dtype == ref.dtype == np.float32
    if name == 'sum':     # different (tree) summation order: 1e-5 relative, the north-star tolerance
      np.testing.assert_allclose(got, ref, rtol=1e-5)
    else:
      all_eq(np.asarray(got), np.asarray(ref))
  small = rng.randint(1, 3, size=shape).astype(np.int32)
  got = sp.prod(sp.from_numpy(small), axis).glom(); ref = oexpr.prod(oexpr.from_numpy(small), axis).glom()
  assert got.dtype == ref.dtype == np.int64
  all_eq(np.asarray(got), np.asarray(ref))
  b = rng.randint(0, 2, size=shape).astype(np.int32)
  for name in ('all', 'any'):
    got = getattr(sp, name)(sp.from_numpy(b), axis).glom(); ref = getattr(oexpr, name)(oexpr.from_numpy(b), axis).glom()
    assert got.dtype == np.bool_
    all_eq(np.asarray(got), np.asarray(ref))


def test_min_max_prod_logic_reference():
  # tests/test_statistics.py:16-30, tests/test_mathematics.py:9-15, tests/test_logic.py:9-23
  src = np.asarray([1, 1, 1, 2, 2, 5, 5, 10])
  all_eq(sp.max(sp.from_numpy(src)).glom(), np.max(src))
  all_eq(sp.min(sp.from_numpy(src)).glom(), np.min(src))
  src = np.arange(100).reshape(10, 10)
  all_eq(sp.min(sp.from_numpy(src), axis=1).glom(), np.min(src, axis=1))
  nA = np.arange(40000, dtype=np.int32).reshape(100, 400)
  got = sp.from_numpy(nA).prod().glom()
  assert got.dtype == np.int64
  all_eq(got, nA.astype(np.int64).prod())
  nC = (nA.T.copy() // 1000)
  C = sp.from_numpy(nA.T.copy()) / 1000          # Python-2 era integer divide floors
  all_eq(C.glom(), nC)
  all_eq(sp.all(C).glom(), np.all(nC))
  all_eq(sp.any(C).glom(), np.any(nC))


def test_fused_map_reduce_vs_oracle():
  """(x*2+y).sum(axis=0), BASELINE config 3 at a size the oracle finishes instantly; 1e-5 relative."""
  rng = np.random.default_rng(2)
  x = rng.random((1024, 2048), dtype=np.float32); y = rng.random((1024, 2048), dtype=np.float32)
  for hint in (None, (128, 2048), (256, 512)):
    e = (sp.from_numpy(x, tile_hint=hint) * 2 + sp.from_numpy(y, tile_hint=hint)).sum(axis=0)
    got_f = e.optimized().glom(); got_u = e.glom()
    ref = (oexpr.from_numpy(x, tile_hint=hint) * 2 + oexpr.from_numpy(y, tile_hint=hint)).sum(axis=0).optimized().glom()
    exact = (x.astype(np.float64) * 2 + y).sum(axis=0)
    assert got_f.dtype == ref.dtype == np.float32
    np.testing.assert_allclose(got_f, ref, rtol=1e-5)
    np.testing.assert_allclose(got_u, ref, rtol=1e-5)
    np.testing.assert_allclose(got_f, exact, rtol=1e-5)


def test_optimization_reduced():
  # tests/test_optimization.py:124-163 without the slices (views are out of scope), fp64, tolerance 1e-6
  rng = np.random.RandomState(1)
  na = rng.rand(300, 300); nb = rng.rand(300, 300)
  a = sp.from_numpy(na); b = sp.from_numpy(nb)
  c = a - b; d = a + c; h = c - d; i = c + h
  m = h + i; n = i - m; o = n - m; q = n + o; r = q - m
  s = sp.sum(r)
  nc = na - nb; nd = na + nc; nh = nc - nd; ni = nc + nh
  nm = nh + ni; nn = ni - nm; no = nn - nm; nq = nn + no; nr = nq - nm
  ns = np.sum(nr)
  got = s.optimized().glom()
  assert abs(got - ns) < 1e-6 * max(1.0, abs(ns))


# ------------------------------------------------------------------ creation (tests/test_creation.py:17-68)
def test_arange():
  pytest.raises(ValueError, sp.arange)
  all_eq(sp.arange((10,)).glom(), np.arange(10))
  all_eq(sp.arange((3, 5)).glom(), np.arange(15).reshape((3, 5)))
  all_eq(sp.arange((10,), -1).glom(), np.arange(-1, 9))
  all_eq(sp.arange((3, 5), -1).glom(), np.arange(-1, 14).reshape((3, 5)))
  all_eq(sp.arange((10,), step=2).glom(), np.arange(0, 20, 2))
  all_eq(sp.arange((3, 5), 1, step=2).glom(), np.arange(1, 31, 2).reshape((3, 5)))
  all_eq(sp.arange(stop=10).glom(), np.arange(10))
  all_eq(sp.arange(-1, 19, 2).glom(), np.arange(-1, 19, 2))
  all_eq(sp.arange((64, 48), dtype=np.int64, tile_hint=(16, 16)).glom(), np.arange(64 * 48).reshape(64, 48))


def test_from_numpy_roundtrip_and_tilings():
  rng = np.random.RandomState(6)
  for shape, hint in [((37, 53), None), ((37, 53), (10, 53)), ((37, 53), (37, 7)), ((37, 53), (8, 9)), ((5,), (2,)), ((2, 3, 4), (1, 3, 2))]:
    x = rng.randn(*shape)
    all_eq(sp.from_numpy(x, tile_hint=hint).glom(), x)
    all_eq((sp.from_numpy(x, tile_hint=hint) + 1).glom(), x + 1)


def test_rand_properties():
  """Device RNG (Philox): uniform [0,1), deterministic per seed, independent of the tiling."""
  a = sp.rand(512, 256, seed=7, dtype=np.float32).glom()
  b = sp.rand(512, 256, seed=7, dtype=np.float32, tile_hint=(64, 64)).glom()
  all_eq(a, sp.rand(512, 256, seed=7, dtype=np.float32, tile_hint=(512, 64)).glom())     # full-height column tiles
  all_eq(a, sp.rand(512, 256, seed=7, dtype=np.float32, tile_hint=(100, 256)).glom())    # full-width row tiles
  c = sp.rand(512, 256, seed=8, dtype=np.float32).glom()
  assert a.dtype == np.float32 and a.min() >= 0.0 and a.max() < 1.0
  all_eq(a, b)
  assert not np.array_equal(a, c)
  assert abs(a.mean() - 0.5) < 5e-3 and abs(a.var() - 1 / 12.0) < 5e-3
  n = sp.randn(512, 256, seed=9, dtype=np.float32).glom()
  assert abs(n.mean()) < 1e-2 and abs(n.std() - 1.0) < 1e-2
  assert sp.rand(8, 8).glom().dtype == np.float64      # srandom.py:83 default dtype


# ------------------------------------------------------------------ dot (tests/test_dot.py:8-103, test_matmul.py:12-22)
def test_dot_reference_cases_exact():
  all_eq(sp.dot(sp.arange((132, 100)), sp.arange((100, 77))).glom(),
                np.dot(np.arange(13200.).reshape(132, 100), np.arange(7700.).reshape(100, 77)))
  all_eq(sp.dot(sp.arange((67, 100)), sp.arange((100, 77))).glom(),
                np.dot(np.arange(6700.).reshape(67, 100), np.arange(7700.).reshape(100, 77)))
  all_eq(sp.dot(sp.arange((77, 100)), np.arange(8800.).reshape(100, 88)).glom(),
                np.dot(np.arange(7700.).reshape(77, 100), np.arange(8800.).reshape(100, 88)))
  all_eq(sp.dot(sp.arange(stop=100), sp.arange(stop=100)).glom(), np.asarray([np.dot(np.arange(100.), np.arange(100.))]))
  all_eq(sp.dot(sp.arange((100, 77)), sp.arange(stop=77)).glom(), np.dot(np.arange(7700.).reshape(100, 77), np.arange(77.)))
  all_eq(sp.dot(sp.arange((77, 100)), sp.arange(stop=100)).glom(), np.dot(np.arange(7700.).reshape(77, 100), np.arange(100.)))
  all_eq(sp.dot(sp.arange((77, 100)), np.arange(100.)).glom(), np.dot(np.arange(7700.).reshape(77, 100), np.arange(100.)))
  x = sp.arange((100, 50), dtype=np.int64).astype(np.float64); y = sp.arange((50, 100), dtype=np.int64).astype(np.float64)
  all_eq(sp.dot(x, y).glom(), np.dot(np.arange(5000.).reshape(100, 50), np.arange(5000.).reshape(50, 100)))
  xi = sp.arange((40, 30), dtype=np.int64); yi = sp.arange((30, 20), dtype=np.int64)
  all_eq(sp.dot(xi, yi).glom(), np.dot(np.arange(1200).reshape(40, 30), np.arange(600).reshape(30, 20)))


@pytest.mark.parametrize('M,N,K,hint', [(256, 512, 384, None), (300, 700, 1000, (128, 256)), (1024, 1024, 2048, (512, 512)),
                                         (130, 70, 4100, (64, 64))])
def test_dot_fp32_tensor_core_vs_oracle(M, N, K, hint):
  """fp32 dot on tcgen05.  Oracle = np.dot (what the reference tests assert).
  tf32x3: max |C - C_ref| <= 1e-5 * max|C_ref| on zero-mean data (normwise; the element-wise relative
  error of a cancelling sum is unbounded in any arithmetic); tf32x1 on uniform data: 2e-4."""
  rng = np.random.default_rng(0)
  a = rng.standard_normal((M, K), dtype=np.float32); b = rng.standard_normal((K, N), dtype=np.float32)
  ref = np.dot(a.astype(np.float64), b.astype(np.float64))
  old = sp.FLAGS.dot_precision
  # the oracle's TILED dot (K-joins + np.add merges, dot.py:195-217 / tile.pyx:263-268) on the same tiling
  spartan_oracle.initialize(1)
  oref = oexpr.dot_grid(oexpr.from_numpy(a, tile_hint=hint), oexpr.from_numpy(b, tile_hint=hint), tile_hint=hint).glom()
  assert np.abs(oref - ref).max() <= 1e-5 * np.abs(ref).max()
  try:
    for prec in ('bf16x3', 'tf32x3'):        # bf16x3 is the default (and benchmarked) mode
      sp.FLAGS.dot_precision = prec
      got = sp.dot(sp.from_numpy(a, tile_hint=hint), sp.from_numpy(b, tile_hint=hint), tile_hint=hint).glom()
      assert got.dtype == np.float32
      assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max(), prec
      assert np.abs(got - np.dot(a, b)).max() <= 1e-5 * np.abs(ref).max(), prec
      assert np.abs(got - oref).max() <= 1e-5 * np.abs(ref).max(), prec
    sp.FLAGS.dot_precision = 'simt'
    got = sp.dot(sp.from_numpy(a, tile_hint=hint), sp.from_numpy(b, tile_hint=hint), tile_hint=hint).glom()
    assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max()
    sp.FLAGS.dot_precision = 'tf32x1'
    au = rng.random((M, K), dtype=np.float32); bu = rng.random((K, N), dtype=np.float32)
    got = sp.dot(sp.from_numpy(au, tile_hint=hint), sp.from_numpy(bu, tile_hint=hint), tile_hint=hint).glom()
    refu = np.dot(au.astype(np.float64), bu.astype(np.float64))
    assert np.abs(got - refu).max() <= 2e-4 * np.abs(refu).max()
  finally:
    sp.FLAGS.dot_precision = old


@pytest.mark.parametrize('prec', ['bf16x3', 'tf32x3'])
def test_dot_full_depth_zero_mean(prec):
  """The benchmark's contraction depth on the hard dataset: K = 32768, zero-mean operands (sums cancel, so operand
  rounding is not averaged away), 2048 x 2048 output, float64 on the host as the reference.  Tolerance: north_star's
  1e-5, normwise."""
  rng = np.random.default_rng(7)
  M = N = 2048
  K = 32768
  a = rng.standard_normal((M, K), dtype=np.float32); b = rng.standard_normal((K, N), dtype=np.float32)
  ref = a.astype(np.float64) @ b.astype(np.float64)
  old = sp.FLAGS.dot_precision
  try:
    sp.FLAGS.dot_precision = prec
    got = sp.dot(sp.from_numpy(a, tile_hint=(512, 4096)), sp.from_numpy(b, tile_hint=(4096, 512)), tile_hint=(512, 512)).glom()
  finally:
    sp.FLAGS.dot_precision = old
  err = np.abs(got - ref).max() / np.abs(ref).max()
  assert err <= 1e-5, (prec, err)
  # uniform [0, 1) operands of the same shape (the easy case the bench times)
  a = rng.random((M, K), dtype=np.float32); b = rng.random((K, N), dtype=np.float32)
  ref = a.astype(np.float64) @ b.astype(np.float64)
  try:
    sp.FLAGS.dot_precision = prec
    got = sp.dot(sp.from_numpy(a), sp.from_numpy(b)).glom()
  finally:
    sp.FLAGS.dot_precision = old
  assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max()


def test_dot_prepared_operand_cache_tracks_updates():
  """Prepared operands of unchanged arrays are reused (same bits as the uncached call); an in-place update of either
  operand, or a merge into one of its tiles, invalidates them."""
  from spartan_b200 import device_ops
  from spartan_b200.expr.base import lazify
  rng = np.random.default_rng(3)
  a = rng.standard_normal((640, 512), dtype=np.float32); b = rng.standard_normal((512, 384), dtype=np.float32)
  A = sp.from_numpy(a, tile_hint=(128, 128)).evaluate(); B = sp.from_numpy(b, tile_hint=(128, 128)).evaluate()
  device_ops.prepared_cache.clear()
  first = sp.dot(lazify(A), lazify(B)).glom()
  h0 = device_ops.prepared_cache.hits
  again = sp.dot(lazify(A), lazify(B)).glom()
  assert device_ops.prepared_cache.hits >= h0 + 2, 'second evaluation did not hit the prepared-operand cache'
  all_eq(first, again)
  sp.FLAGS.dot_prepared_cache = False
  try:
    all_eq(sp.dot(lazify(A), lazify(B)).glom(), first)
  finally:
    sp.FLAGS.dot_prepared_cache = True
  a2 = rng.standard_normal((640, 512), dtype=np.float32)
  A.update(sp.extent.from_shape(A.shape), a2)
  got = sp.dot(lazify(A), lazify(B)).glom()
  ref = a2.astype(np.float64) @ b.astype(np.float64)
  assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max()
  B.update(sp.extent.create((0, 0), (100, 384), B.shape), np.zeros((100, 384), np.float32))
  b2 = b.copy(); b2[:100] = 0
  got = sp.dot(lazify(A), lazify(B)).glom()
  ref = a2.astype(np.float64) @ b2.astype(np.float64)
  assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max()


_COMBINERS = [np.add, np.minimum, np.maximum, np.multiply]


@pytest.mark.parametrize('reducer', _COMBINERS, ids=[r.__name__ for r in _COMBINERS])
@pytest.mark.parametrize('dtype,udtype', [(np.float32, np.float32), (np.float32, np.float64), (np.int64, np.int64),
                                          (np.int64, np.int32), (np.float64, np.float32)])
def test_combiner_update_regions_vs_oracle(reducer, dtype, udtype):
  """The combiner itself (Tile.merge, tile.pyx:200-297, driven through DistArrayImpl.update, distarray.py:372-422):
  first write to an element replaces, later writes reduce -- on partial regions that overlap each other and cross tile
  boundaries, on whole tiles (fast path incl. its first-element test and its cast to the tile dtype, :263-268), and on
  the whole array; update data of another dtype than the tile.  Same update sequence into the oracle's tiles and the
  device tiles; every element ends up written, results must be bit-identical."""
  shape, hint = (37, 53), (16, 20)
  rng = np.random.default_rng(11)

  def data(shp):
    # Updates of another dtype than the tile carry small integers: the reference's full-tile reduce lets the tile's
    # dtype drift to the wider one (`old_tile.data = reducer(old_tile.data, update)`, tile.pyx:265, no cast) while a
    # device tile keeps its dtype, so only values exact in both are comparable bit for bit.
    if np.dtype(udtype).kind == 'f' and np.dtype(udtype) == np.dtype(dtype):
      return (rng.random(shp) * 4 - 1).astype(udtype)
    return rng.integers(-3, 4, size=shp).astype(udtype)

  regions = [((3, 5), (30, 41)),        # partial, crosses tile boundaries, misses every tile's first element but one
             ((10, 0), (37, 25)),       # overlaps the first
             ((16, 20), (32, 40)),      # exactly one tile: full-tile fast path on a tile whose first element was written
             ((0, 40), (16, 53)),       # one whole edge tile, never written before: fast path replaces
             ((0, 0), (37, 53)),        # whole array: one full-tile update per tile
             ((1, 1), (36, 52)),        # partial again
             ((0, 0), (16, 20))]        # a whole tile once more
  spartan_oracle.initialize(1)
  from spartan_oracle import distarray as odist, extent as oext
  oarr = odist.create(shape, dtype, reducer=reducer, tile_hint=hint)
  darr = sp.distarray.create(shape, dtype, reducer=reducer, tile_hint=hint)
  for ul, lr in regions:
    upd = data(tuple(b - a for a, b in zip(ul, lr)))
    oarr.update(oext.create(ul, lr, shape), upd)
    darr.update(sp.extent.create(ul, lr, shape), upd)
  want = oarr.glom()
  got = darr.glom()
  assert got.dtype == want.dtype
  all_eq(got, want)


def test_combiner_full_tile_before_first_element_replaces():
  """tile.pyx:263-268: a full-tile update reduces only if the tile's first element has been written; after a partial
  write that misses it, the full-tile update REPLACES the tile (reference behaviour, kept)."""
  spartan_oracle.initialize(1)
  from spartan_oracle import distarray as odist, extent as oext
  shape = (8, 8)
  oarr = odist.create(shape, np.float32, reducer=np.add, tile_hint=shape)
  darr = sp.distarray.create(shape, np.float32, reducer=np.add, tile_hint=shape)
  part = np.full((4, 4), 5, np.float32); whole = np.arange(64, dtype=np.float32).reshape(8, 8)
  for arr, ex in ((oarr, oext), (darr, sp.extent)):
    arr.update(ex.create((2, 2), (6, 6), shape), part)
    arr.update(ex.create((0, 0), (8, 8), shape), whole)
    arr.update(ex.create((0, 0), (8, 8), shape), whole)
  all_eq(darr.glom(), oarr.glom())
  all_eq(darr.glom(), 2 * whole)


@pytest.mark.parametrize('M,K,N,ha,hb,hc', [(96, 64, 80, (24, 64), (16, 80), (48, 40)),      # rows > cols: outer route
                                            (64, 200, 48, (16, 200), (50, 48), None),       # map2 route, A by rows
                                            (40, 120, 70, (40, 30), (30, 70), (20, 35)),    # map2 route, A by columns
                                            (300, 40, 70, (75, 40), (10, 70), (100, 70))])
def test_dot_through_map2_and_outer_vs_oracle(M, K, N, ha, hb, hc):
  """dot as an instance of the join operators, routed like dot.py:281-294 (outer for rows > cols, else map2 on
  (axis 1, axis 0)): strips fetched by change_partition_axis, rank-k partials np.add-merged into the target.  The
  oracle evaluates the same route with NumPy tiles; integer-valued float64 data makes the comparison exact."""
  rng = np.random.default_rng(M + K)
  a = rng.integers(-4, 5, size=(M, K)).astype(np.float64); b = rng.integers(-4, 5, size=(K, N)).astype(np.float64)
  for workers in (1, 3):
    spartan_oracle.initialize(workers)
    want = oexpr.dot(oexpr.from_numpy(a, tile_hint=ha), oexpr.from_numpy(b, tile_hint=hb), tile_hint=hc).glom()
    all_eq(want, a @ b)
  got = sp.dot_as_join(sp.from_numpy(a, tile_hint=ha), sp.from_numpy(b, tile_hint=hb), tile_hint=hc).glom()
  all_eq(got, want)
  all_eq(sp.dot(sp.from_numpy(a, tile_hint=ha), sp.from_numpy(b, tile_hint=hb), tile_hint=hc).glom(), want)
  # fp32 on the tensor cores through the same join
  af, bf = a.astype(np.float32), b.astype(np.float32)
  got = sp.dot_as_join(sp.from_numpy(af, tile_hint=ha), sp.from_numpy(bf, tile_hint=hb), tile_hint=hc).glom()
  assert got.dtype == np.float32
  all_eq(got, (a @ b).astype(np.float32))       # small integers: every mode is exact


def test_map2_user_tile_function_with_reducer_target():
  """A third join, not one the library hard-codes: per row strip, the column sums of x*y, merged with np.add into a
  (cols,) target -- a device tile function built from the fused map+reduce kernel, against the oracle's map2 running the
  NumPy version of the same function."""
  from spartan_b200 import device_ops, _lib
  from spartan_b200.array import extent as dext
  from spartan_oracle import extent as oext
  rng = np.random.default_rng(5)
  x = rng.integers(-5, 6, size=(90, 40)).astype(np.float32); y = rng.integers(-5, 6, size=(90, 40)).astype(np.float32)

  def target_extents(extents):
    return [dext.create((0,), (extents[0].shape[1],), (extents[0].array_shape[1],))]

  @sp.device_tile_function(target_extents)
  def strip_colsum(extents, tiles):
    import torch
    out = torch.empty((tiles[0].shape[1],), dtype=torch.float32, device=tiles[0].device)
    prog = device_ops.make_program([('IN', 0), ('IN', 1), ('MUL', 0)], _lib.SP_F32)
    device_ops.run_map_reduce(prog, [tiles[0], tiles[1]], tuple(tiles[0].shape), 0, _lib.SP_RED_SUM, out, False)
    return [out]

  def strip_colsum_np(extents, tiles):
    yield oext.create((0,), (extents[0].shape[1],), (extents[0].array_shape[1],)), (tiles[0] * tiles[1]).sum(axis=0)

  spartan_oracle.initialize(3)
  want = oexpr.map2((oexpr.from_numpy(x, tile_hint=(20, 40)), oexpr.from_numpy(y, tile_hint=(30, 40))), (0, 0),
                    fn=strip_colsum_np, shape=(40,), reducer=np.add).glom()
  got = sp.map2((sp.from_numpy(x, tile_hint=(20, 40)), sp.from_numpy(y, tile_hint=(30, 40))), (0, 0), fn=strip_colsum,
                shape=(40,), reducer=np.add).glom()
  all_eq(got, want)
  all_eq(got, (x * y).sum(axis=0))
  with pytest.raises(sp.NotDeviceMappable):
    sp.map2([sp.from_numpy(x)], [0], fn=strip_colsum_np, shape=(40,))


def test_dot_linearity_large():
  """Size-independent property at a shape the oracle would need minutes for: dot(A, B+C) = dot(A,B)+dot(A,C)
  and dot(A, e_j-columns) reproduces A columns exactly (identity is exact in TF32)."""
  n = 4096
  A = sp.rand(n, n, seed=1, dtype=np.float32, tile_hint=(1024, 1024))
  I = sp.from_numpy(np.eye(n, dtype=np.float32), tile_hint=(1024, 1024))
  got = sp.dot(A, I, tile_hint=(1024, 1024)).glom()
  a = A.glom()
  # default mode bf16x3 represents each operand with 16 mantissa bits: |err| <= 2^-17 per element here
  np.testing.assert_allclose(got, a, rtol=1e-5, atol=1e-7)


# ------------------------------------------------------------------ streaming (smem-ring) fast path
@pytest.mark.parametrize('shape', [(300, 512), (64, 4100), (1, 1 << 20), (1000, 1024), (7, 3, 2048)])
@pytest.mark.parametrize('dtype', [np.float32, np.float64, np.int64])
def test_streaming_map_bit_exact(shape, dtype):
  """Rows >= 1 KiB take the TMA-bulk/shared-memory-ring kernels; results must stay bit-exact."""
  rng = np.random.RandomState(11)
  if np.dtype(dtype).kind == 'f':
    x = rng.randn(*shape).astype(dtype); y = rng.randn(*shape).astype(dtype); z = rng.randn(*shape).astype(dtype)
  else:
    x = rng.randint(-99, 99, size=shape).astype(dtype); y = rng.randint(1, 99, size=shape).astype(dtype)
    z = rng.randint(-9, 9, size=shape).astype(dtype)
  row = x[..., :1, :] * 0 + np.arange(shape[-1]).astype(dtype)            # broadcast along rows
  col = (x[..., :, :1] * 0 + 3).astype(dtype)                            # broadcast along columns
  builds = (lambda m, X, Y, Z, R, C: X * 2 + Y,
            lambda m, X, Y, Z, R, C: (X - Y) * Z + X,                    # three streamed operands
            lambda m, X, Y, Z, R, C: X * R + C,
            lambda m, X, Y, Z, R, C: m.maximum(X, Y) - m.abs(Z))
  for build in builds:
    got = build(sp, *[sp.from_numpy(a) for a in (x, y, z, row, col)]).optimized().glom()
    ref = build(oexpr, *[oexpr.from_numpy(a) for a in (x, y, z, row, col)]).optimized().glom()
    assert got.dtype == ref.dtype
    all_eq(got, ref)


@pytest.mark.parametrize('shape', [(1000, 1024), (129, 4100), (4096, 512), (5, 300, 256)])
def test_streaming_reduce_axis0(shape):
  rng = np.random.RandomState(12)
  xi = rng.randint(-50, 50, size=shape).astype(np.int64); yi = rng.randint(-50, 50, size=shape).astype(np.int64)
  xf = rng.rand(*shape).astype(np.float32); yf = rng.rand(*shape).astype(np.float32)
  xd = rng.rand(*shape)
  axis = len(shape) - 2
  for name in ('sum', 'min', 'max'):
    got = getattr(sp, name)(sp.from_numpy(xi) * 3 - sp.from_numpy(yi), axis).optimized().glom()
    ref = getattr(oexpr, name)(oexpr.from_numpy(xi) * 3 - oexpr.from_numpy(yi), axis).optimized().glom()
    all_eq(got, ref)
  got = (sp.from_numpy(xf) * 2 + sp.from_numpy(yf)).sum(axis=axis).optimized().glom()
  ref = (xf.astype(np.float64) * 2 + yf).sum(axis=axis)
  np.testing.assert_allclose(got, ref, rtol=1e-5)
  got = sp.from_numpy(xd).sum(axis=axis).glom()
  np.testing.assert_allclose(got, xd.sum(axis=axis), rtol=1e-12)
  all_eq(sp.max(sp.from_numpy(xf), axis).glom(), xf.max(axis=axis))


@pytest.mark.parametrize('shape', [(1000, 1024), (129, 4100), (4096, 512), (37, 70000), (5, 300, 2048)])
def test_streaming_reduce_trailing_axis(shape):
  """Reductions over the LAST axis of long rows run on the streaming kernel (MODE 2: lanes folded by shuffle, segments
  and panels in fixed order); integers bit-exact, floats against float64, and arg-reductions exercise the position leaf
  whose strides are re-mapped for that kernel."""
  rng = np.random.RandomState(13)
  xi = rng.randint(-50, 50, size=shape).astype(np.int64); yi = rng.randint(-50, 50, size=shape).astype(np.int64)
  xf = rng.rand(*shape).astype(np.float32); yf = rng.rand(*shape).astype(np.float32)
  xd = rng.rand(*shape)
  axis = len(shape) - 1
  for name in ('sum', 'min', 'max'):
    got = getattr(sp, name)(sp.from_numpy(xi) * 3 - sp.from_numpy(yi), axis).optimized().glom()
    all_eq(got, getattr(np, name)(xi * 3 - yi, axis=axis))
  got = (sp.from_numpy(xf) * 2 + sp.from_numpy(yf)).sum(axis=axis).optimized().glom()
  np.testing.assert_allclose(got, (xf.astype(np.float64) * 2 + yf).sum(axis=axis), rtol=1e-5)
  got = (sp.abs(sp.from_numpy(xf) - sp.from_numpy(yf)) * sp.from_numpy(xf)).sum(axis=axis).optimized().glom()
  np.testing.assert_allclose(got, (np.abs(xf.astype(np.float64) - yf) * xf).sum(axis=axis), rtol=1e-5)
  np.testing.assert_allclose(sp.from_numpy(xd).sum(axis=axis).glom(), xd.sum(axis=axis), rtol=1e-12)
  all_eq(sp.max(sp.from_numpy(xf), axis).glom(), xf.max(axis=axis))
  all_eq(np.asarray(sp.argmin(sp.from_numpy(xf), axis).glom()), np.asarray(xf.argmin(axis=axis)))
  all_eq(np.asarray(sp.argmax(sp.from_numpy(xi), axis).glom()), np.asarray(xi.argmax(axis=axis)))
  if len(shape) == 2:      # row tiles: one launch per contiguous block of rows
    got = sp.from_numpy(xf, tile_hint=(max(1, shape[0] // 3), shape[1])).sum(axis=1).glom()
    np.testing.assert_allclose(got, xf.astype(np.float64).sum(axis=1), rtol=1e-5)


def test_gemm_split_form_matches_one_call():
  """sp_gemm_prepare_a/_b + sp_gemm_prepared (the multi-GPU form) == sp_gemm_f32 (bit-identical)."""
  import torch
  from spartan_b200 import device_ops, blob_ctx
  ctx = blob_ctx.get()
  torch.manual_seed(0)
  M, K, N = 384, 1000, 520
  A = torch.randn(M, K, device=ctx.device); B = torch.randn(K, N, device=ctx.device)
  for prec in ('tf32x1', 'tf32x3', 'bf16x3'):
    C1 = torch.empty(M, N, device=ctx.device); C2 = torch.empty(M, N, device=ctx.device)
    device_ops.gemm([(A, B)], C1, precision=prec)
    Kp = device_ops.gemm_kpad(K, prec)
    pa = torch.zeros(device_ops.gemm_prepared_bytes(M, Kp, prec), dtype=torch.uint8, device=ctx.device)
    pb = torch.zeros(device_ops.gemm_prepared_bytes(N, Kp, prec), dtype=torch.uint8, device=ctx.device)
    # two K strips written at their depth offsets, like the strips of two source ranks
    device_ops.gemm_prepare_a(A[:, :600].contiguous(), pa, Kp, 0, prec); device_ops.gemm_prepare_a(A[:, 600:].contiguous(), pa, Kp, 600, prec)
    device_ops.gemm_prepare_b(B[:600], pb, Kp, 0, prec); device_ops.gemm_prepare_b(B[600:], pb, Kp, 600, prec)
    device_ops.gemm_prepared([(pa, pb, Kp)], C2, accumulate=False, precision=prec)
    assert torch.equal(C1, C2), prec


# ------------------------------------------------------------------ tests/test_reduce.py:38-90 argmin / argmax
def test_argmin_argmax_reference_cases():
  nx1 = np.arange(TEST_SIZE, dtype=np.int64)
  all_eq(sp.arange((TEST_SIZE,), dtype=np.int64).argmin().glom(), nx1.argmin())
  all_eq(sp.arange((TEST_SIZE,), dtype=np.int64).argmax().glom(), nx1.argmax())
  nx2 = np.arange(TEST_SIZE * TEST_SIZE, dtype=np.int64).reshape((TEST_SIZE, TEST_SIZE))
  x2 = sp.arange((TEST_SIZE, TEST_SIZE), dtype=np.int64)
  all_eq(x2.argmin(axis=1).glom(), nx2.argmin(axis=1))
  all_eq(x2.argmax(axis=1).glom(), nx2.argmax(axis=1))
  nx3 = np.arange(TEST_SIZE ** 3, dtype=np.int64).reshape((TEST_SIZE,) * 3)
  for axis in [None, 0, 1, 2]:
    x3 = sp.arange((TEST_SIZE,) * 3, dtype=np.int64)
    all_eq(x3.argmin(axis).glom(), nx3.argmin(axis))
    all_eq(x3.argmax(axis).glom(), nx3.argmax(axis))


@pytest.mark.parametrize('shape,hint', [((257, 129), None), ((64, 4100), (16, 4100)), ((300, 700), (100, 128)), ((5, 6, 7), None)])
def test_argmin_argmax_random(shape, hint):
  """Index results are bit-exact, including ties (smallest index wins, like np.argmin)."""
  rng = np.random.RandomState(13)
  for x in (rng.randn(*shape).astype(np.float32), rng.randint(-5, 5, size=shape).astype(np.int64)):
    for axis in [None] + list(range(len(shape))):
      all_eq(np.asarray(sp.argmin(sp.from_numpy(x, tile_hint=hint), axis).glom()), np.asarray(x.argmin(axis)))
      all_eq(np.asarray(sp.argmax(sp.from_numpy(x, tile_hint=hint), axis).glom()), np.asarray(x.argmax(axis)))


# ------------------------------------------------------------------ k-means and SpMV (reference parity unpinned)
@pytest.mark.parametrize('n,d,k,tile_rows', [(20000, 32, 16, 5000), (3000, 50, 7, 1000), (70000, 256, 64, 70000)])
def test_kmeans_vs_oracle(n, d, k, tile_rows):
  """Labels come from fp32 tensor-core distances vs SciPy's float64 cdist: points within ~1e-5 relative distance of
  a cluster boundary may flip.  Tolerance: >= 99.9 % identical labels after the first pass, centres within 1e-3."""
  from spartan_oracle import apps
  rng = np.random.default_rng(4)
  X = rng.random((n, d), dtype=np.float32)
  c0 = X[:k].copy()
  km = sp.KMeans(n_clusters=k, n_iter=1)
  centers, labels = km.fit(sp.from_numpy(X, tile_hint=(tile_rows, d)), centers=c0)
  ref_c, ref_l = apps.kmeans_fit(X, c0, 1, tile_rows)
  got_l = labels.glom()
  assert got_l.dtype == np.int32 and centers.dtype == np.float32
  assert (got_l == ref_l).mean() >= 0.999
  # every label that differs must be a near tie: the two candidate centres are equally close to 1e-4 relative
  bad = np.nonzero(got_l != ref_l)[0]
  if len(bad):
    xb = X[bad].astype(np.float64); c64 = c0.astype(np.float64)
    d_got = ((xb - c64[got_l[bad]]) ** 2).sum(1); d_ref = ((xb - c64[ref_l[bad]]) ** 2).sum(1)
    assert np.all(np.abs(d_got - d_ref) <= 1e-4 * d_ref)
  # centres: identical up to the points that flipped (each moves a centre by ~1/count) and fp32 accumulation
  np.testing.assert_allclose(centers, ref_c, rtol=1e-2, atol=1e-3)
  counts = np.bincount(got_l, minlength=k)
  sums = np.zeros((k, d)); np.add.at(sums, got_l, X.astype(np.float64))
  np.testing.assert_allclose(centers, sums / np.maximum(counts, 1)[:, None], rtol=1e-5, atol=1e-6)   # own labels: tight
  centers3, _ = sp.KMeans(n_clusters=k, n_iter=3).fit(sp.from_numpy(X, tile_hint=(tile_rows, d)), centers=c0)
  ref_c3, _ = apps.kmeans_fit(X, c0, 3, tile_rows)
  np.testing.assert_allclose(centers3, ref_c3, rtol=5e-2, atol=5e-3)


@pytest.mark.parametrize('n,outlinks,strip', [(5000, 10, 625), (12345, 3, 5000), (40000, 10, 40000)])
def test_spmv_vs_oracle(n, outlinks, strip):
  """PageRank step y = W p (benchmark_pagerank.py shape).  fp32 sums in a different order: 1e-5 relative."""
  from spartan_oracle import apps
  W = apps.make_weights(n, outlinks, seed=5)
  rng = np.random.default_rng(6)
  p = rng.random((n, 1), dtype=np.float32)
  wts = sp.sparse.from_scipy(W, strip_width=strip)
  got = sp.dot(wts, sp.from_numpy(p, tile_hint=(strip, 1))).glom()
  ref = apps.spmv_strips(W, p, strip)
  exact = (W.tocsr().astype(np.float64) @ p.astype(np.float64)).reshape(-1)
  assert got.shape == (n, 1) and got.dtype == np.float32
  np.testing.assert_allclose(got.reshape(-1), ref, rtol=1e-5, atol=1e-6)
  np.testing.assert_allclose(got.reshape(-1), exact, rtol=1e-5, atol=1e-6)
  ones = sp.dot(wts, sp.ones((n, 1), tile_hint=(strip, 1))).glom().reshape(-1)       # p = ones, as in the benchmark
  np.testing.assert_allclose(ones, np.asarray(W.tocsr().sum(axis=1)).reshape(-1), rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------ edge cases: 0-d, empty, ragged, tiny
@pytest.mark.parametrize('n,d,k', [(5000, 32, 16), (33333, 256, 1024), (777, 64, 300), (4096, 128, 257)])
def test_kmeans_fused_epilogue_matches_two_kernel_form(n, d, k):
  """The assignment fused into the GEMM (running arg min across the column tiles in registers, accumulation by the
  epilogue warps) against the candidates + second-kernel form: identical labels and counts, sums equal up to the order of
  the float atomics; both against float64 distances."""
  from spartan_b200._lib import lib
  rng = np.random.default_rng(n + d + k)
  x = rng.random((n, d), dtype=np.float32)
  c0 = x[rng.choice(n, k, replace=False)].copy()
  X = sp.from_numpy(x, tile_hint=(n, d)).evaluate()
  out = {}
  try:
    for fused in (1, 0):
      lib.sp_kmeans_set_fused(fused)
      centers, labels = sp.KMeans(n_clusters=k, n_iter=1).fit(X, centers=c0)
      out[fused] = (centers, labels.glom())
  finally:
    lib.sp_kmeans_set_fused(1)
  all_eq(out[1][1], out[0][1])
  np.testing.assert_allclose(out[1][0], out[0][0], rtol=2e-6, atol=1e-7)
  x64, c64 = x.astype(np.float64), c0.astype(np.float64)
  d2 = (x64 * x64).sum(1)[:, None] - 2.0 * x64 @ c64.T + (c64 * c64).sum(1)[None, :]
  ref = d2.argmin(1)
  bad = np.nonzero(out[1][1] != ref)[0]
  gap = d2[bad, out[1][1][bad]] - d2[bad, ref[bad]]
  assert np.all(gap <= 1e-5 * np.abs(d2[bad, ref[bad]])), 'labels differ away from a tie'


def test_kmeans_and_spmv_on_reference_generated_vectors():
  """The device k-means iteration and the device SpMV on the inputs of tests/golden/app_vectors.json, whose outputs
  were produced by the reference's own mapper functions (oracle/ref_apps/make_app_vectors.py): labels must be
  identical (the points are cast to float32 first; a label may only differ where float64 distances are a near tie),
  counts / centres and the strip products within float32 accuracy."""
  import json, os
  import scipy.sparse
  with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'app_vectors.json')) as f:
    vec = json.load(f)
  for case in vec['kmeans']:
    pts = np.array(case['points']); centers = np.array(case['centers']); k, n = case['k'], case['n']
    X = sp.from_numpy(pts.astype(np.float32), tile_hint=(case['tile_rows'], case['d']))
    new_centers, labels = sp.KMeans(n_clusters=k, n_iter=1).fit(X, centers=centers.astype(np.float32))
    got = labels.glom()
    want = np.array([v for t in case['tiles'] for v in t['labels']])
    if not np.array_equal(got, want):
      d2 = ((pts[:, None, :] - centers[None, :, :]) ** 2).sum(-1)
      bad = np.nonzero(got != want)[0]
      gap = d2[bad, got[bad]] - d2[bad, want[bad]]
      assert np.all(gap <= 1e-5 * d2[bad, want[bad]]), 'labels differ away from a tie'
    else:
      counts = np.sum([t['counts'] for t in case['tiles']], axis=0).astype(np.float64)
      sums = np.sum([t['sums'] for t in case['tiles']], axis=0)
      ok = counts > 0
      np.testing.assert_allclose(new_centers[ok], sums[ok] / counts[ok, None], rtol=2e-6)
  for case in vec['spmv']:
    n, strip = case['n'], case['strip']
    m = scipy.sparse.coo_matrix((np.array(case['vals'], np.float32), (case['rows'], case['cols'])), shape=(n, n))
    x = np.array(case['x'], np.float32).reshape(n, 1)
    want = np.zeros(n, np.float64)
    for s in case['strips']:
      want += np.array(s['partial'], np.float64)
    got = sp.dot(sp.sparse.from_scipy(m, strip_width=strip), sp.from_numpy(x, tile_hint=(strip, 1))).glom().reshape(-1)
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-6)


def test_c_abi_housekeeping_and_flag_words():
  """The self-sufficient part of the boundary: device selection, stream-ordered tile buffers, events, and the flag words
  consumers wait on (sp_write_u32 / sp_wait_u32 incl. its bounded time-out)."""
  import ctypes, torch
  from spartan_b200._lib import lib, check
  check(lib.sp_init(0), 'sp_init')
  stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
  p = ctypes.c_void_p()
  check(lib.sp_tile_alloc(1 << 20, ctypes.byref(p), stream), 'sp_tile_alloc')
  assert p.value
  e0, e1 = ctypes.c_void_p(), ctypes.c_void_p()
  check(lib.sp_event_create(ctypes.byref(e0)), 'ev'); check(lib.sp_event_create(ctypes.byref(e1)), 'ev')
  check(lib.sp_event_record(e0, stream), 'rec')
  check(lib.sp_fill(p, 0, (1 << 20) // 4, 0, 3.0, 0.0, 0, 0, stream), 'sp_fill')          # float32 constant fill
  check(lib.sp_write_u32(p, 7, stream), 'sp_write_u32')
  status = torch.zeros(1, dtype=torch.int32, device='cuda')
  check(lib.sp_wait_u32(p, 7, 1000, ctypes.c_void_p(status.data_ptr()), stream), 'sp_wait_u32')      # already satisfied
  check(lib.sp_wait_u32(p, 8, 20, ctypes.c_void_p(status.data_ptr()), stream), 'sp_wait_u32')        # never: times out
  check(lib.sp_event_record(e1, stream), 'rec')
  ms = ctypes.c_float()
  check(lib.sp_event_elapsed(e0, e1, ctypes.byref(ms)), 'elapsed')
  check(lib.sp_sync(stream), 'sp_sync')
  assert 15.0 <= ms.value < 500.0, ms.value            # the 20 ms time-out is inside the timed region
  assert int(status.item()) == 1
  host = torch.empty(4, dtype=torch.float32)
  torch.cuda.synchronize()
  view = (ctypes.c_float * 4).from_address(host.data_ptr())
  check(lib.sp_download_2d(ctypes.c_void_p(host.data_ptr()), 16, p, 16, 16, 1, stream), 'download')
  check(lib.sp_sync(stream), 'sp_sync')
  assert host[1].item() == 3.0 and np.frombuffer(host.numpy().tobytes(), np.uint32)[0] == 7
  check(lib.sp_tile_free(p, stream), 'free')
  check(lib.sp_event_destroy(e0), 'd'); check(lib.sp_event_destroy(e1), 'd')
  check(lib.sp_shutdown(), 'sp_shutdown')


def test_zero_dim_and_empty_arrays():
  s = sp.from_numpy(np.array(3.0, dtype=np.float32))
  assert (s * 2 + 1).glom() == 7.0
  assert (s * 2 + 1).glom().shape == ()
  for shape in [(0,), (0, 5), (3, 0)]:
    z = sp.zeros(shape)
    assert z.glom().shape == shape
    assert (z + 1).glom().shape == shape
    assert sp.sum(z).glom() == 0.0
  assert sp.sum(sp.zeros((0, 5)), axis=0).glom().shape == (5,)
  all_eq(sp.sum(sp.zeros((0, 5)), axis=0).glom(), np.zeros((5,), np.float32))


def test_general_expression_outside_the_static_catalogue():
  """A chain that needs a temporary (both operands of '+' are sub-trees) and reuses an operand: interpreter path,
  streaming and non-streaming shapes; +,-,*,abs,max are exact so the result is bit-identical to NumPy."""
  rng = np.random.RandomState(21)
  for shape in [(40, 33), (600, 2048), (3, 1 << 18)]:
    x = rng.randn(*shape).astype(np.float32); y = rng.randn(*shape).astype(np.float32)
    X, Y = sp.from_numpy(x), sp.from_numpy(y)
    e = (sp.abs(X - Y) * X + sp.maximum(Y, 0.5))
    all_eq(e.optimized().glom(), np.abs(x - y) * x + np.maximum(y, np.float32(0.5)))
    got = e.sum(axis=0).optimized().glom()
    np.testing.assert_allclose(got, (np.abs(x - y).astype(np.float64) * x + np.maximum(y, 0.5)).sum(axis=0), rtol=2e-5, atol=1e-4)
    all_eq(((X - Y) / (sp.abs(Y) + 1) - (X * X - 3)).optimized().glom(), (x - y) / (np.abs(y) + 1) - (x * x - 3))


@pytest.mark.parametrize('M,N,K', [(256, 256, 64), (129, 70, 200), (384, 520, 1000), (1000, 1300, 2500), (2048, 1024, 4096)])
def test_gemm_cta_pair_matches_single_cta(M, N, K):
  """The cta_group::2 kernel (256 x 256 tile per CTA pair) and the one-CTA kernel issue the same MMAs on the same
  chunks in the same order: results must agree bit for bit -- ragged edges, several segments, accumulate --
  and both must meet the fp32 bar against float64."""
  import torch
  from spartan_b200 import device_ops, blob_ctx
  from spartan_b200._lib import lib, check
  ctx = blob_ctx.get()
  g = torch.Generator(device='cpu'); g.manual_seed(M * 7 + N)
  A = torch.randn(M, K, generator=g).to(ctx.device); B = torch.randn(K, N, generator=g).to(ctx.device)
  C0 = torch.randn(M, N, generator=g).to(ctx.device)
  ref = A.double().cpu().numpy() @ B.double().cpu().numpy()
  k1 = (K // 3 + 3) // 4 * 4
  segs = [(A[:, :k1].contiguous(), B[:k1].contiguous()), (A[:, k1:].contiguous(), B[k1:].contiguous())]
  try:
    for prec in ('bf16x3', 'tf32x3', 'tf32x1'):
      out = {}
      for variant in (1, 2):
        check(lib.sp_gemm_set_variant(variant), 'sp_gemm_set_variant')
        C = torch.empty(M, N, device=ctx.device)
        device_ops.gemm([(A, B)], C, precision=prec)
        Cs = C0.clone()
        device_ops.gemm(segs, Cs, accumulate=True, precision=prec)
        out[variant] = (C.cpu().numpy(), Cs.cpu().numpy())
      all_eq(out[1][0], out[2][0])
      all_eq(out[1][1], out[2][1])
      if prec != 'tf32x1':
        assert np.abs(out[2][0] - ref).max() <= 1e-5 * np.abs(ref).max()
        assert np.abs(out[2][1] - (ref + C0.cpu().numpy())).max() <= 2e-5 * np.abs(ref).max()
  finally:
    check(lib.sp_gemm_set_variant(0), 'sp_gemm_set_variant')


@pytest.mark.parametrize('M,N,K,strip', [(1024, 768, 512, 256), (700, 900, 333, 256), (512, 2048, 640, 512),
                                         (2048, 300, 1000, 512), (2304, 4400, 300, 1024)])
@pytest.mark.parametrize('prec', ['bf16x3', 'tf32x3'])
def test_streamed_dot_matches_resident(M, N, K, strip, prec):
  _streamed_dot_case(M, N, K, strip, prec)


def _streamed_dot_case(M, N, K, strip, prec):
  """dot(from_numpy(a), from_numpy(b)) with the PCIe upload pipelined against the contraction (strips of A rows /
  B columns, L-shaped frontier) must give the bits of the resident path -- ragged strips, unequal strip counts,
  unaligned K -- read back both block by block (read_local_into, event driven) and through glom; the operand
  arrays must be left resident and cached like from_numpy(...).evaluate() leaves them."""
  import torch
  from spartan_b200.expr.base import eval_cache
  rng = np.random.default_rng(M + N + K)
  a = rng.standard_normal((M, K), dtype=np.float32); b = rng.standard_normal((K, N), dtype=np.float32)
  old = (sp.FLAGS.dot_stream_host_operands, sp.FLAGS.dot_stream_strip, sp.FLAGS.dot_stream_min_bytes, sp.FLAGS.dot_precision)
  try:
    sp.FLAGS.dot_precision = prec
    sp.FLAGS.dot_stream_host_operands = False
    want = sp.dot(sp.from_numpy(a), sp.from_numpy(b)).glom()
    sp.FLAGS.dot_stream_host_operands = True
    sp.FLAGS.dot_stream_strip = strip
    sp.FLAGS.dot_stream_min_bytes = 0
    ea, eb = sp.from_numpy(a, tile_hint=(256, 256)), sp.from_numpy(b)
    e = sp.dot(ea, eb, tile_hint=(256, 256))
    c = e.evaluate()
    assert c.block_events, 'the streamed path did not run'
    out = torch.empty((M, N), dtype=torch.float32, pin_memory=True)
    nbytes = c.read_local_into(out.numpy())
    torch.cuda.current_stream().synchronize()
    assert nbytes == M * N * 4 and c.block_events is None
    all_eq(out.numpy(), want)
    all_eq(c.glom(), want)
    all_eq(ea.evaluate().glom(), a)          # cached by the streamed evaluation, resident and complete
    all_eq(eb.evaluate().glom(), b)
    assert eval_cache.get(ea.expr_id) is not None
    ref = a.astype(np.float64) @ b.astype(np.float64)
    assert np.abs(want - ref).max() <= 1e-5 * np.abs(ref).max()
  finally:
    (sp.FLAGS.dot_stream_host_operands, sp.FLAGS.dot_stream_strip, sp.FLAGS.dot_stream_min_bytes, sp.FLAGS.dot_precision) = old


# ------------------------------------------------------------------ views: tests/test_slice.py, test_transpose.py, test_reshape.py
@pytest.mark.parametrize('hint', [None, (4, 4), (3, 10)])
def test_slice_reference_cases(hint):
  """tests/test_slice.py:25-80 on the device: slices are zero-copy strided operands of the fused kernels."""
  T = 10
  # the reference's _arange_mapper (creation.py:135-141) numbers a tile as one contiguous run, which is only right for
  # full-row tiles; with a grid hint the device (np.arange(n).reshape(shape), the documented result) is checked alone
  for m in ((sp, oexpr) if hint != (4, 4) else (sp,)):
    x = m.arange((T, T), tile_hint=hint)
    nx = np.arange(T * T).reshape(T, T)
    all_eq(x[5:8, 5:8].evaluate().glom(), nx[5:8, 5:8])                          # test_slice_get
    all_eq((x[5:8, 5:8] + 1).glom(), nx[5:8, 5:8] + 1)                           # test_slice_map
    all_eq((x[2:9, 1:3] * x[1:8, 5:7]).optimized().glom(), nx[2:9, 1:3] * nx[1:8, 5:7])
    x3 = m.arange((10, 10, 10), dtype=np.int64); n3 = np.arange(1000).reshape((10, 10, 10))
    all_eq((x3[:, :, 0] + 13).glom().reshape(10, 10), n3[:, :, 0] + 13)          # test_slice_map2
    all_eq(x3[:, :, 0].sum().glom(), n3[:, :, 0].sum())                          # test_slice_reduce
    all_eq(x3[2:7, :, 3:9].sum(axis=1).glom(), n3[2:7, :, 3:9].sum(axis=1))
    a = m.arange((T,), dtype=np.int64); na = np.arange(T)
    all_eq((a[1:] - a[:-1]).glom(), na[1:] - na[:-1])                            # test_slice_sub
    all_eq((a[1:] - a[:-1]).optimized().glom(), na[1:] - na[:-1])
  assert sp.extent.from_slice((slice(None), slice(None), 0), [100, 100, 100]).shape == (100, 100, 1)


def test_getitem_integer_and_newaxis():
  """base.py:401-448: integer indices drop their dimension (NumPy semantics; the reference's sequential pops mis-handle
  several integers and a bare integer keeps a unit dimension -- the evident intent is implemented)."""
  n3 = np.arange(4 * 5 * 6, dtype=np.float32).reshape(4, 5, 6)
  x = sp.from_numpy(n3, tile_hint=(2, 5, 3))
  all_eq(x[1].glom(), n3[1])
  all_eq(x[:, 2].glom(), n3[:, 2])
  all_eq(x[1, :, 4].glom(), n3[1, :, 4])
  all_eq(x[-1].glom(), n3[-1])
  all_eq(x[:, sp.newaxis, 2:4].glom(), n3[:, np.newaxis, 2:4])
  all_eq((x[1] * 2 + x[3]).optimized().glom(), n3[1] * 2 + n3[3])
  all_eq(x[1:3].sum(axis=0).glom(), n3[1:3].sum(axis=0))


@pytest.mark.parametrize('hint', [None, (500, 200), (64, 1347)])
def test_transpose_reference_cases(hint):
  """tests/test_transpose.py:9-37."""
  t2 = np.transpose(np.reshape(np.arange(3721 * 1347), (3721, 1347)))
  all_eq(sp.transpose(sp.arange((3721, 1347), tile_hint=hint)).glom(), t2)                        # transpose1
  t3 = np.transpose(np.reshape(np.arange(101 * 102 * 103), (101, 102, 103)))
  all_eq(sp.transpose(sp.arange((101, 102, 103))).glom(), t3)                                     # transpose2
  rng = np.random.RandomState(0)
  n1 = rng.random_sample((401, 97)); n2 = rng.random_sample((401, 97))
  got = sp.dot(sp.from_numpy(n1), sp.transpose(sp.from_numpy(n2))).glom()                                # transpose_dot (f64: exact path)
  assert np.all(np.isclose(np.dot(n1, np.transpose(n2)), got))
  f1, f2 = n1.astype(np.float32), n2.astype(np.float32)
  for a, b, ref in ((sp.from_numpy(f1), sp.transpose(sp.from_numpy(f2)), f1.astype(np.float64) @ f2.T.astype(np.float64)),
                    (sp.transpose(sp.from_numpy(f1)), sp.from_numpy(f2), f1.T.astype(np.float64) @ f2.astype(np.float64)),
                    (sp.from_numpy(f1).T, sp.from_numpy(f2.T.copy()).T, f1.T.astype(np.float64) @ f2.astype(np.float64))):
    got = sp.dot(a, b).glom()            # fp32: tensor-core path reads the transposed views in place (sp_gemm_f32_ex)
    assert got.dtype == np.float32 and np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max()
  x = rng.random_sample((300, 200)).astype(np.float32); y = rng.random_sample((200, 300)).astype(np.float32)
  got, want = both(lambda m: (m.transpose(m.from_numpy(x)) * 2 + m.from_numpy(y)).sum(axis=0).optimized())
  np.testing.assert_allclose(got, want, rtol=1e-5)
  all_eq((sp.from_numpy(x).T - sp.from_numpy(y)).glom(), x.T - y)


def test_reshape_reference_cases():
  """tests/test_reshape.py:9-88,98-121 (dense)."""
  all_eq(sp.reshape(sp.arange((10, 10)), (100,)).glom(), sp.arange((100,)).glom())                # reshape1
  b = sp.reshape(sp.arange((1000,), tile_hint=[100]), (10, 100)).evaluate()                              # reshape2
  sp.reshape(b, (1000,)).evaluate()
  d = sp.reshape(sp.reshape(sp.reshape(sp.arange((100, 100)), (10000,)), (10000, 1)), (1, 10000))
  all_eq(d.glom(), sp.arange((1, 10000)).glom())                                                   # reshape3
  f = sp.arange((10000,))
  for shp in ((10, 1000), (1000, 10), (20, 500), (500, 20), (1, 10000)):
    f = sp.reshape(f, shp)
  all_eq(f.glom(), sp.arange((1, 10000)).glom())                                                   # reshape4
  for n, s1, s2 in ((35511, (133, 267), (267, 133)), (12319, (127, 97), (97, 127))):                      # reshape5, 6
    d = sp.reshape(sp.reshape(sp.reshape(sp.arange((n,)), s1), s2), (1, n))
    all_eq(d.glom(), sp.arange((1, n)).glom())
  targets = [(23, 120, 100), (12, 230, 100), (276000, 1), (1, 276000)]                                    # reshape7
  for src in ((100, 23, 120), (12, 23, 1000), (1, 276000), (276000, 1), (276000,)):
    a = sp.arange(src)
    for shp in targets:
      all_eq(sp.reshape(a, shp).glom(), np.arange(276000).reshape(shp))
  rng = np.random.RandomState(1)                                                                          # reshape_dot
  n1 = rng.random_sample((357, 93)); n2 = rng.random_sample((31, 357))
  all_eq(np.dot(np.reshape(n1, (1071, 31)), n2),
                sp.dot(sp.reshape(sp.from_numpy(n1), (1071, 31)), sp.from_numpy(n2)).glom(), 10e-9)
  n1 = rng.random_sample((357, 718)); n2 = rng.random_sample((718,))
  all_eq(np.dot(n1, np.reshape(n2, (718, 1))),
                sp.dot(sp.from_numpy(n1), sp.reshape(sp.from_numpy(n2), (718, 1))).glom(), 10e-9)
  n1 = rng.random_sample((718,)); n2 = rng.random_sample((1, 357))
  all_eq(np.dot(np.reshape(n1, (718, 1)), n2),
                sp.dot(sp.reshape(sp.from_numpy(n1), (718, 1)), sp.from_numpy(n2)).glom(), 10e-9)
  # maps and reductions over reshaped / ravelled operands, tiled bases
  x = rng.random_sample((60, 70)).astype(np.float32)
  sx = sp.from_numpy(x, tile_hint=(16, 32))
  all_eq((sp.reshape(sx, (70, 60)) * 3).glom(), x.reshape(70, 60) * 3)
  all_eq((sp.ravel(sx) + 1).glom(), x.ravel() + 1)
  np.testing.assert_allclose(sp.reshape(sx, (4, 15, 70)).sum(axis=1).glom(), x.reshape(4, 15, 70).sum(axis=1), rtol=1e-5)


def _jit_stats():
  import ctypes
  from spartan_b200._lib import lib
  c, l, f = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
  lib.sp_jit_stats(ctypes.byref(c), ctypes.byref(l), ctypes.byref(f))
  return c.value, l.value, f.value


@pytest.mark.parametrize('dtype', [np.float32, np.float64, np.int64])
def test_runtime_specialised_kernels_match_interpreter(dtype):
  """A fused chain outside the static catalogue on a large array is compiled once by NVRTC into a straight-line
  instance of the map / reduce kernels (csrc/jit.cu; the reference's per-expression codegen, local.py:58-152).  It runs
  the same exec_op code as the interpreter, so maps and order-independent reductions must be bit-identical to the
  interpreted run and to NumPy (+,-,*,abs,max round once per operation under -fmad=false)."""
  from spartan_b200._lib import lib
  rng = np.random.RandomState(5)
  shape = (2048, 1024)
  if dtype == np.int64:
    x = rng.randint(-1000, 1000, size=shape).astype(dtype); y = rng.randint(-1000, 1000, size=shape).astype(dtype)
    z = rng.randint(-1000, 1000, size=shape).astype(dtype)
    c = 3
  else:
    x = rng.randn(*shape).astype(dtype); y = rng.randn(*shape).astype(dtype); z = rng.randn(*shape).astype(dtype)
    c = dtype(0.5)
  X, Y, Z = sp.from_numpy(x), sp.from_numpy(y), sp.from_numpy(z)

  def build():
    return [(sp.abs(X - Y) * X + sp.maximum(Y, c)).optimized(),               # 2 operands, one temporary
            ((X + Y) * Z - X * c + Y * Y - Z).optimized(),                    # 3 operands (NI = 8 kernel)
            (sp.abs(X - Y) * X + sp.maximum(Y, c)).sum(axis=0).optimized(),   # map+reduce
            ((X - Z) * (Y + Z)).max(axis=0).optimized()]
  want = [np.abs(x - y) * x + np.maximum(y, c), (x + y) * z - x * c + y * y - z]
  try:
    lib.sp_jit_enable(0)
    interp = [e.glom() for e in build()]
    lib.sp_jit_enable(1)
    c0, l0, f0 = _jit_stats()
    jit = [e.glom() for e in build()]
    jit2 = [e.glom() for e in build()]          # second run: cache hits, no new compiles
    c1, l1, f1 = _jit_stats()
  finally:
    lib.sp_jit_enable(1)
  assert f1 == f0, 'run-time specialisation failed: %s' % lib.sp_jit_last_log().decode()
  assert l1 - l0 == 8 and c1 - c0 <= 4, (c0, c1, l0, l1)
  for i, (a, b, c_) in enumerate(zip(interp, jit, jit2)):
    all_eq(b, c_)                              # deterministic: two compiled runs agree bit for bit
    if i == 2 and np.dtype(dtype).kind == 'f':
      # a floating-point sum: the compiled chain runs on the direct reduce kernel, the interpreted one on the ring
      # kernel -- same values, another (fixed) summation order
      ref = (np.abs(x.astype(np.float64) - y) * x + np.maximum(y, c)).sum(axis=0)
      tol = 1e-5 if dtype == np.float32 else 1e-12
      assert np.abs(b - ref).max() <= tol * np.abs(ref).max() and np.abs(a - ref).max() <= tol * np.abs(ref).max()
    else:
      all_eq(a, b)
  all_eq(jit[0], want[0])
  all_eq(jit[1], want[1])
  all_eq(jit[3], ((x - z) * (y + z)).max(axis=0))


def test_replayable_evaluation_tracks_input_updates():
  """sp.replayable: one captured evaluate() (fused map+reduce, finalise, flat sum, a 3-operand map) replayed as a CUDA
  graph gives the eager result bit for bit, and replays see in-place updates of the input arrays."""
  from spartan_b200.expr.base import lazify
  rng = np.random.RandomState(11)
  x = rng.rand(2048, 1024).astype(np.float32); y = rng.rand(2048, 1024).astype(np.float32)
  X = sp.from_numpy(x, tile_hint=(256, 1024)).evaluate(); Y = sp.from_numpy(y, tile_hint=(256, 1024)).evaluate()

  def step():
    a, b = lazify(X), lazify(Y)
    return [(a * 2 + b).sum(axis=0).optimized().evaluate(), (sp.abs(a - b) * a + b * b).optimized().evaluate(),
            (a * b).sum().optimized().evaluate()]
  eager = [r.glom() for r in step()]
  rep = sp.replayable(step)
  assert rep.kernel_launches >= 3
  for _ in range(3):
    got = [r.glom() for r in rep()]
  for e, g in zip(eager, got):
    all_eq(e, g)
  x2 = rng.rand(2048, 1024).astype(np.float32)
  X.update(sp.extent.from_shape(X.shape), x2)          # new data in the same device array
  got = [r.glom() for r in rep()]
  np.testing.assert_allclose(got[0], (x2.astype(np.float64) * 2 + y).sum(axis=0), rtol=1e-5)
  all_eq(got[1], np.abs(x2 - y) * x2 + y * y)
  np.testing.assert_allclose(got[2], (x2.astype(np.float64) * y).sum(), rtol=1e-5)


@pytest.mark.parametrize('iszip', [False, True])
def test_save_load_checkpoint_roundtrip(tmp_path, iszip):
  """tests/test_fio.py:24-31: save / load round trips straight out of and into HBM, the files readable by the
  reference's loader (oracle restatement); checkpoint() writes once and load_data() restores."""
  rng = np.random.RandomState(8)
  x = rng.rand(300, 200).astype(np.float32)
  t1 = sp.from_numpy(x, tile_hint=(64, 50))
  assert sp.save(t1, 'fiotest1', str(tmp_path), iszip) is True
  all_eq(t1.glom(), sp.load('fiotest1', str(tmp_path), iszip).glom())
  all_eq(spartan_oracle.fio.load('fiotest1', str(tmp_path), iszip).glom(), x)
  np.testing.assert_allclose((sp.load('fiotest1', str(tmp_path), iszip) * 2 + t1).sum(axis=0).optimized().glom(),
                             (x.astype(np.float64) * 2 + x).sum(axis=0), rtol=1e-5)
  old = sp.FLAGS.checkpoint_path
  try:
    sp.FLAGS.checkpoint_path = str(tmp_path / 'ckpt')
    c = sp.expr.checkpoint(t1 + 1)
    all_eq(c.glom(), x + 1)
    all_eq(c.load_data().glom(), x + 1)
  finally:
    sp.FLAGS.checkpoint_path = old


def test_gemm_round_sync_is_bit_identical():
  """The soft round barrier of the CTA-pair GEMM (clusters check in at every tile round so shared operand tiles stay in
  L2) only changes WHEN tiles start: more tiles than clusters, ragged edges, accumulate -- same bits with it on and off."""
  import torch
  from spartan_b200 import device_ops, blob_ctx
  from spartan_b200._lib import lib, check
  ctx = blob_ctx.get()
  g = torch.Generator(device='cpu'); g.manual_seed(5)
  M, N, K = 3000, 3500, 700
  A = torch.randn(M, K, generator=g).to(ctx.device); B = torch.randn(K, N, generator=g).to(ctx.device)
  C0 = torch.randn(M, N, generator=g).to(ctx.device)
  out = {}
  try:
    for on in (0, 1):
      check(lib.sp_gemm_set_round_sync(on), 'sp_gemm_set_round_sync')
      C = C0.clone()
      device_ops.gemm([(A, B)], C, accumulate=True, precision='bf16x3')
      out[on] = C.cpu().numpy()
  finally:
    check(lib.sp_gemm_set_round_sync(1), 'sp_gemm_set_round_sync')
  all_eq(out[0], out[1])
  ref = C0.double().cpu().numpy() + A.double().cpu().numpy() @ B.double().cpu().numpy()
  assert np.abs(out[1] - ref).max() <= 1e-5 * np.abs(ref).max()


# ------------------------------------------------------------------ tests/test_creation.py:10-16,65-92, test_statistics.py:32-70
@pytest.mark.parametrize('hint', [None, (16, 4), (7, 10)])
def test_eye_identity_diag(hint):
  """eye / identity as an extent-aware fill (INDEX leaf), diagonal / diagflat / diag as strided rectangle copies."""
  all_eq(sp.eye(100, 10, tile_hint=hint).glom(), np.eye(100, 10))
  all_eq(sp.eye(40, 60, k=3, tile_hint=hint).glom(), np.eye(40, 60, k=3))
  all_eq(sp.eye(40, 60, k=-5, dtype=np.float64, tile_hint=hint).glom(), np.eye(40, 60, k=-5))
  all_eq(sp.identity(100).glom(), np.identity(100))
  got, want = both(lambda m: m.eye(100, 10))
  all_eq(got, want)
  rng = np.random.RandomState(2)
  for shp in ((2, 2), (15, 10), (16, 16), (10, 33)):
    x = rng.randn(*shp)
    all_eq(sp.diagonal(sp.from_numpy(x, tile_hint=hint if hint and shp[0] > 2 else None)).glom(), np.diagonal(x))
    all_eq(oexpr.diagonal(oexpr.from_numpy(x)).glom(), np.diagonal(x))
  x = rng.randn(57, 57)
  all_eq(sp.diag(sp.from_numpy(x)).glom(), np.diag(x))
  all_eq(sp.diag(sp.diag(sp.from_numpy(x))).glom(), np.diag(np.diag(x)))
  xf = rng.randn(300).astype(np.float32)
  all_eq(sp.diagflat(sp.from_numpy(xf, tile_hint=(64,)), tile_hint=(128, 100)).glom(), np.diagflat(xf))


def test_std_reference_cases():
  rng = np.random.RandomState(4)
  for shp in ((10,), (10, 10), (17, 17)):
    x = rng.randn(*shp)
    got, want = both(lambda m: m.std(m.from_numpy(x)))
    assert abs(got - np.std(x)) < 1e-6 and abs(got - want) < 1e-6       # Assert.float_close
  for shp in ((10, 10), (15, 13), (13, 15), (17, 17)):
    x = rng.randn(*shp)
    for axis in (0, 1):
      got, want = both(lambda m: m.std(m.from_numpy(x), axis))
      assert np.allclose(got, np.std(x, axis)) and np.allclose(got, want)
  xb = rng.rand(600, 4096).astype(np.float32)
  np.testing.assert_allclose(sp.std(sp.from_numpy(xb, tile_hint=(128, 4096)), 0).optimized().glom(), np.std(xb.astype(np.float64), 0), rtol=1e-6)
