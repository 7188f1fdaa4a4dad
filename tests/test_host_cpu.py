"""CPU-side checks of the product host: C-ABI exports, native extent/tiling vs the oracle, the
bytecode compiler's dtype rules vs NumPy (through the oracle), fusion structure.  No GPU needed."""
import ctypes
import os
import re

import numpy as np
import pytest

import spartan_oracle
from spartan_oracle import distarray as odist, expr as oexpr, extent as oex

import spartan_b200 as sp
from spartan_b200 import _lib
from spartan_b200.array import distarray as pdist, extent as pex
from spartan_b200.expr import program
from spartan_b200.expr.reduce import _value_tree

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_exports_every_declared_symbol():
  header = open(os.path.join(ROOT, 'include', 'spartan_b200.h')).read()
  declared = set(re.findall(r'\b(sp_[a-z0-9_]+)\s*\(', header))
  declared -= {'sp_status', 'sp_dtype'}
  lib = ctypes.CDLL(_lib.LIB_PATH)
  missing = [name for name in sorted(declared) if not hasattr(lib, name)]
  assert not missing, 'header declares symbols the library does not export: %s' % missing
  assert set(_lib.EXPORTS) <= declared, 'python binding uses undeclared symbols: %s' % (set(_lib.EXPORTS) - declared)
  assert lib.sp_version() >= 100


def test_compute_calls_fail_loudly_without_gpu():
  import torch
  if torch.cuda.is_available():
    pytest.skip('GPU present')
  with pytest.raises(sp.SpartanError):
    sp.ones((4, 4)).glom()


@pytest.mark.parametrize('shape,hint,shards', [
  ((4096, 4096), None, 1), ((4096, 4096), None, 3), ((4096, 4096), None, 8), ((32768, 32768), None, 8),
  ((50, 50, 50), None, 3), ((10,), None, 3), ((10 ** 7, 256), None, 8), ((32768, 32768), (4096, 4096), 8),
  ((132, 100), (50, 33), 3), ((7,), (2,), -1), ((), None, 3), ((5, 1), None, 4)])
def test_tiling_matches_oracle(shape, hint, shards):
  a = pdist.compute_extents(shape, hint, shards)
  b = odist.compute_extents(shape, hint, shards)
  assert [(ex.ul, ex.lr, w) for ex, w in a.items()] == [(ex.ul, ex.lr, w) for ex, w in b.items()]
  if hint is None and shape:
    assert pdist.good_tile_shape(shape, shards) == odist.good_tile_shape(shape, shards)


def test_native_unravel_to_global_match_reference_values():
  # tests/test_extent.py:24-38 through the C ABI
  a = pex.create((2, 2), (7, 7), (10, 10))
  assert [a.to_global(i, axis=None) for i in (0, 10, 11, 20)] == [22, 42, 43, 62]
  rnd = np.random.RandomState(0)
  for _ in range(200):
    shp = tuple(int(x) for x in rnd.randint(1, 30, size=rnd.randint(1, 5)))
    idx = int(rnd.randint(0, int(np.prod(shp))))
    assert pex.unravelled_pos(idx, shp) == oex.unravelled_pos(idx, shp) == tuple(int(i) for i in np.unravel_index(idx, shp))


def test_grid_change_partition_axis_is_refused():
  ex = pex.create((0, 0), (4, 4), (8, 8))
  with pytest.raises(sp.SpartanError):
    pex.change_partition_axis(ex, 1)


def _compile(e, dtypes, scalars=()):
  ops = {}
  for var, child in zip(e.child_to_var, e.children):
    if var in scalars:
      ops[var] = program.Operand('scalar', np.asarray(scalars[var]).dtype, value=np.asarray(scalars[var])[()])
    else:
      ops[var] = program.Operand('array', dtypes[var] if var in dtypes else child.npa.dtype)
  tree = _value_tree(e.op) if isinstance(e, sp.ReduceExpr) else e.op
  return program.compile_tree(tree, ops)


@pytest.mark.parametrize('dt', [np.float32, np.float64, np.int32, np.int64])
@pytest.mark.parametrize('scalar', [2, 2.5, np.float64(0.5), np.int64(3)])
def test_scalar_casting_rule_matches_numpy1(dt, scalar):
  """x * scalar keeps the array dtype unless the scalar is of a higher kind;
  the expected dtype comes from the oracle's legacy_result_type, itself pinned by test_ln."""
  x = sp.from_numpy(np.zeros((4,), dt))
  e = x * scalar
  var_s = e.child_to_var[1]
  c = _compile(e, {e.child_to_var[0]: dt}, {var_s: scalar})
  expect = oexpr.legacy_result_type([np.zeros((4,), dt), np.asarray(scalar)])
  assert c.out_dtype == expect


def test_fused_program_shape():
  x = sp.from_numpy(np.zeros((4, 4), np.float32)); y = sp.from_numpy(np.zeros((4, 4), np.float32))
  e = (x * 2 + y).sum(axis=0).optimized()
  assert isinstance(e, sp.ReduceExpr) and len(e.children) == 3
  assert all(not isinstance(c, sp.MapExpr) for c in e.children)
  c = _compile(e, {v: np.float32 for v in e.child_to_var}, {e.child_to_var[1]: 2})
  ops = [(c.program.op[i], c.program.arg[i]) for i in range(c.program.n_ops)]
  O = _lib.OP
  assert ops == [(O['IN'], 0), (O['CONST'], 0), (O['MUL'], 0), (O['IN'], 1), (O['ADD'], 0)]
  assert c.compute_dtype == _lib.SP_F32 and c.out_dtype == np.float32 and c.program.consts[0] == 2.0


def test_unfused_evaluate_keeps_three_nodes():
  # Expr.evaluate() does not optimise (base.py:300)
  x = sp.from_numpy(np.zeros((4, 4), np.float32))
  e = (x * 2 + x).sum()
  assert isinstance(e.children[0], sp.MapExpr) and isinstance(e.children[0].children[0], sp.MapExpr)


def test_python_lambda_is_rejected():
  x = sp.from_numpy(np.zeros((4,), np.float32))
  e = sp.map(x, lambda t: t + 1)
  with pytest.raises(sp.NotDeviceMappable):
    _compile(e, {e.child_to_var[0]: np.float32})


def test_dtype_of_mixed_chain():
  a = sp.from_numpy(np.zeros((4,), np.float32)); b = sp.from_numpy(np.zeros((4,), np.float64))
  i = sp.from_numpy(np.zeros((4,), np.int32))
  e = ((a * a) + b).optimized()
  c = _compile(e, {})
  assert c.out_dtype == np.float64 and c.compute_dtype == _lib.SP_F64
  O = _lib.OP
  ops = [c.program.op[k] for k in range(c.program.n_ops)]
  assert O['CAST_F32'] in ops          # the float32 product rounds to float32 before widening, like NumPy
  e = (a + i).optimized()
  c = _compile(e, {})
  assert c.out_dtype == np.result_type(np.float32, np.int32) == np.float64
  e = (i / i)
  c = _compile(e, {})
  assert c.out_dtype == np.int32       # Python-2 era np.divide on ints floors
  e = (a > 0)
  c = _compile(e, {e.child_to_var[0]: np.float32}, {e.child_to_var[1]: 0})
  assert c.out_dtype == np.bool_ and c.compute_dtype == _lib.SP_F32


# ------------------------------------------------------------------ C-ABI argument validation (runs before any launch)
def _prog(ops, compute=None):
  from spartan_b200 import device_ops
  return device_ops.make_program(ops, _lib.SP_F32 if compute is None else compute, [1.0])


def test_cabi_rejects_malformed_programs():
  lib = _lib.lib
  dims = _lib.i64arr([1, 1, 4])
  buf = (ctypes.c_float * 4)()
  o = _lib.sp_operand(); o.ptr = ctypes.addressof(buf); o.dtype = _lib.SP_F32; o.stride[2] = 1
  ins = (_lib.sp_operand * 1)(); ins[0] = o
  cases = [
    ([('ADD', 0)], 'stack underflow'),                                   # binary op on an empty stack
    ([('IN', 0), ('IN', 0)], 'leaves 2 values'),                         # two results
    ([('IN', 3)], 'reads operand 3'),                                    # operand index out of range
    ([('IN', 0)] * 5 + [('ADD', 0)] * 4, 'deeper than 4'),               # exceeds the register stack
    ([(7, 0)], 'unknown opcode'),
  ]
  for ops, msg in cases:
    p = _prog(ops)
    rc = lib.sp_map(ctypes.byref(p), 1, ins, ctypes.byref(o), dims, None)
    assert rc in (_lib.lib.sp_map.restype(-1), -1, -3), (ops, rc)
    assert msg in _lib.last_error(), (ops, _lib.last_error())
  p = _prog([('IN', 0)], compute=_lib.SP_I32)                            # int32 is not a register type
  assert lib.sp_map(ctypes.byref(p), 1, ins, ctypes.byref(o), dims, None) == -3
  p = _prog([('IN', 0)])
  assert lib.sp_map_reduce(ctypes.byref(p), 1, ins, ctypes.byref(o), _lib.i64arr([1, 0, 4]), 0, 0, None, 0, None) == -1
  assert 'empty axis' in _lib.last_error()
  assert lib.sp_map_reduce(ctypes.byref(p), 1, ins, ctypes.byref(o), dims, 99, 0, None, 0, None) == -1


def test_cabi_gemm_and_fill_argument_checks():
  lib = _lib.lib
  assert lib.sp_gemm_f32_workspace_bytes(128, 128, 1, _lib.i64arr([64]), 99) == -1
  segs = (_lib.sp_gemm_segment * 1)()
  assert lib.sp_gemm_f32_segments(0, segs, None, 0, 128, 128, 0, _lib.SP_GEMM_BF16X3, None, 0, None) == -1
  assert lib.sp_gemm_f32_segments(1, segs, None, 0, 128, 128, 0, _lib.SP_GEMM_BF16X3, None, 0, None) == -1
  assert 'workspace' in _lib.last_error()
  assert lib.sp_gemm_kpad(100, _lib.SP_GEMM_BF16X3) == 128 and lib.sp_gemm_kpad(100, _lib.SP_GEMM_TF32X1) == 128
  assert lib.sp_gemm_kpad(33, _lib.SP_GEMM_TF32X3) == 64
  assert lib.sp_gemm_prepared_bytes(10, 64, _lib.SP_GEMM_BF16X3) == 10 * 64 * 2 * 2
  assert lib.sp_fill(None, _lib.SP_F32, -1, 0, 0.0, 0.0, 0, 0, None) == -1
  assert lib.sp_fill(None, _lib.SP_F32, 0, 0, 0.0, 0.0, 0, 0, None) == 0          # empty fill is a no-op
  assert lib.sp_fill(None, _lib.SP_F32, 4, 0, 0.0, 0.0, 0, 0, None) == -1
  assert lib.sp_spmv_csr(None, 1, None, None, 0, None, None, 0, 0, None) == 0
  assert lib.sp_kmeans_assign(None, 0, 4, 0, None, 1, None, None, None, None, 0, None) == -1


def test_broadcast_shapes_including_empty():
  z = sp.zeros((0, 5)); o = sp.ones((1, 5)); c = sp.ones((3, 1))
  assert (z + 1).shape == (0, 5) and (z + o).shape == (0, 5) and (o + c).shape == (3, 5)
  assert (sp.zeros((0,)) + 1).shape == (0,)
  assert sp.sum(z, axis=0).shape == (5,) and sp.sum(z).shape == ()
  from spartan_b200.array import distarray as d
  class A(d.DistArray):
    def __init__(self, shape): self.shape = shape; self.dtype = np.dtype(np.float32); self.tiles = {}
  got = d.broadcast([A((0,)), A(())])
  assert [g.shape for g in got] == [(0,), (0,)]
  got = d.broadcast([A((3, 1)), A((1, 5))])
  assert [g.shape for g in got] == [(3, 5), (3, 5)]


def test_optimized_expressions_leave_no_reference_cycles():
  """A cached result holds device memory: dropping the last reference to an expression must free it at once,
  not whenever the cycle collector next runs (the reference's `opt.optimized_expr = opt` is a self-cycle)."""
  import gc
  a = sp.from_numpy(np.zeros((4, 4), np.float32))
  gc.collect()
  gc.disable()
  try:
    for make in (lambda: (a * 2 + a).optimized(), lambda: (a + a).sum(axis=0).optimized(),
                 lambda: sp.dot(a, a).optimized(), lambda: a.optimized()):
      e = make()
      assert e.optimized() is e
      del e
      assert gc.collect() == 0
  finally:
    gc.enable()


@pytest.mark.parametrize('W', [1, 3, 8])
def test_view_tile_tables_match_oracle(W):
  """Slice / Transpose / Reshape report their tiles in view coordinates; the extents must be exactly the ones the
  reference's mappers hand to a kernel (slice.py:9-39, transpose.py:19-24, reshape.py:32-44), restated by the oracle."""
  from spartan_b200 import blob_ctx
  from spartan_b200.array import views
  from spartan_oracle import views as oviews
  old = blob_ctx._global_ctx[0]
  blob_ctx.set(blob_ctx.BlobCtx(0, W, 'cpu'))
  spartan_oracle.initialize(W)
  try:
    def table(v):
      return sorted((ex.ul, ex.lr, tuple(ex.array_shape)) for ex in v.tiles)

    def otable(v):
      out = v.foreach_tile(lambda ex: [(ex.ul, ex.lr, tuple(ex.array_shape))], {})
      return sorted(t for r in out for t in r)

    cases = [((10, 10), None, np.index_exp[5:8, 5:8]), ((10, 10), (4, 4), np.index_exp[5:8, 5:8]),
             ((10, 10, 10), None, np.index_exp[:, :, 0]), ((100, 37), (16, 10), np.index_exp[3:77, 9:30]),
             ((1000,), (100,), np.index_exp[1:]), ((1000,), (100,), np.index_exp[:-1])]
    for shape, hint, idx in cases:
      base = pdist.create(shape, np.float32, tile_hint=hint)
      obase = odist.create(shape, np.float32, tile_hint=hint)
      v, ov = views.Slice(base, idx), oviews.Slice(obase, idx)
      assert v.shape == tuple(ov.shape)
      assert table(v) == otable(ov), (shape, hint, idx)
      owners = dict(((ex.ul, ex.lr), tid.worker) for ex, tid in base.tiles.items())
      for ex, tid in v.tiles.items():                     # a view tile lives where the base tile that backs it lives
        b = pex.compute_slice(v.slice, ex.to_slice())
        hit = [w for (ul, lr), w in owners.items() if all(u <= bu and bl <= l for u, bu, bl, l in zip(ul, b.ul, b.lr, lr))]
        assert hit == [tid.worker]
    for shape, hint in [((372, 134), None), ((31, 32, 33), None), ((50, 60), (16, 16))]:
      base = pdist.create(shape, np.float32, tile_hint=hint)
      obase = odist.create(shape, np.float32, tile_hint=hint)
      v, ov = views.Transpose(base), oviews.Transpose(obase)
      assert v.shape == tuple(ov.shape) and tuple(v.tile_shape()) == tuple(ov.tile_shape())
      # the reference hands the kernel the BASE extent reversed (transpose.py:19-24): same rectangles, view shape
      assert [(ul, lr) for ul, lr, _ in table(v)] == sorted((ex.ul[::-1], ex.lr[::-1]) for ex in obase.tiles)
    for shape, new in [((10, 10), (100,)), ((1000,), (10, 100)), ((100, 23, 120), (12, 230, 100)), ((60, 70), (60, 70, 1)),
                       ((276000,), (1, 276000))]:
      base = pdist.create(shape, np.float32)
      obase = odist.create(shape, np.float32)
      v, ov = views.Reshape(base, new), oviews.Reshape(obase, new)
      assert v._same_tiles == ov._same_tiles and tuple(v.tile_shape()) == tuple(ov.tile_shape())
      assert table(v) == otable(ov), (shape, new)
  finally:
    blob_ctx._global_ctx[0] = old
    blob_ctx._local.ctx = old


def test_runtime_specialisation_compiles_without_gpu():
  """csrc/jit.cu: the fused chain |x-y|*x + max(y, c) is lowered and compiled by NVRTC for sm_100a from the kernel
  headers embedded in the library (no GPU needed to compile; loading and running it is covered by the -m gpu tests)."""
  from spartan_b200 import device_ops
  ops = [('IN', 0), ('IN', 1), ('SUB', 0), ('ABS', 0), ('IN', 0), ('MUL', 0), ('IN', 1), ('CONST', 0), ('MAX', 0), ('ADD', 0)]
  p = device_ops.make_program(ops, _lib.SP_F32, [0.5])
  n = _lib.lib.sp_jit_compile_check(ctypes.byref(p), 2, 1)
  if n < 0 and 'not found' in _lib.lib.sp_jit_last_log().decode():
    pytest.skip('libnvrtc not present on this machine')
  assert n > 10000, _lib.last_error()


@pytest.mark.parametrize('W', [1, 3])
@pytest.mark.parametrize('iszip', [False, True])
def test_tile_files_interchange_with_the_reference_format(tmp_path, W, iszip):
  """fio.py:46-232: arrays saved by the product load through the oracle's restatement of the reference loader and the
  other way round, and the tile files are the same bytes (tests/test_fio.py:24-31 round trips)."""
  from spartan_b200 import blob_ctx
  old = blob_ctx._global_ctx[0]
  blob_ctx.set(blob_ctx.BlobCtx(0, 1, 'cpu'))         # one product rank owns every tile here; W tiles the oracle side
  spartan_oracle.initialize(W)
  try:
    rng = np.random.RandomState(3)
    for case, (shape, hint, dtype) in enumerate([((100, 37), (16, 10), np.float32), ((50,), None, np.int64),
                                                ((12, 8, 6), (5, 8, 3), np.float64)]):
      x = (rng.rand(*shape) * 100).astype(dtype)
      d1, d2 = str(tmp_path / ('product%d' % case)), str(tmp_path / ('oracle%d' % case))
      arr = pdist.create(shape, dtype, tile_hint=hint)
      arr.update(pex.from_shape(shape), x)
      assert sp.save(arr, 'a', d1, iszip) is True
      assert np.array_equal(spartan_oracle.fio.load('a', d1, iszip).glom(), x)            # product -> reference loader
      oarr = odist.create(shape, dtype, tile_hint=hint)
      oarr.update(oex.from_shape(shape), x)
      assert spartan_oracle.fio.save(oarr, 'a', d2, iszip) is True
      back = sp.load('a', d2, iszip).evaluate()                                           # reference writer -> product
      assert back.dtype == np.dtype(dtype) and np.array_equal(back.glom(), x)
      if hint is not None or W == 1:
        names = sorted(os.listdir(os.path.join(d2, 'a')))
        assert names == sorted(os.listdir(os.path.join(d1, 'a')))
        for n in names:
          opener = bz2_open if (iszip and n.endswith('bz2')) else open
          assert opener(os.path.join(d1, 'a', n), 'rb').read() == opener(os.path.join(d2, 'a', n), 'rb').read(), n
  finally:
    blob_ctx._global_ctx[0] = old
    blob_ctx._local.ctx = old


def bz2_open(path, mode):
  import bz2
  return bz2.BZ2File(path, mode)


def test_view_fetch_semantics_match_numpy_on_host_tensors():
  """The extent translation of Slice / Transpose / Reshape (and Expr.__getitem__ shapes) against NumPy on random cases.
  With one rank every fetch is a zero-copy torch view of the slab, so this runs on CPU tensors without any kernel."""
  from spartan_b200 import blob_ctx
  from spartan_b200.array import views
  old = blob_ctx._global_ctx[0]
  blob_ctx.set(blob_ctx.BlobCtx(0, 1, 'cpu'))
  try:
    rng = np.random.RandomState(12)
    for _ in range(60):
      nd = int(rng.randint(1, 4))
      shape = tuple(int(s) for s in rng.randint(1, 9, size=nd))
      hint = tuple(int(rng.randint(1, s + 1)) for s in shape)
      x = rng.rand(*shape).astype(np.float32)
      arr = pdist.create(shape, np.float32, tile_hint=hint)
      arr.update(pex.from_shape(shape), x)
      # random basic slice
      idx = tuple(slice(int(a), int(a + rng.randint(1, s - a + 1))) for s, a in ((s, rng.randint(0, s)) for s in shape))
      v = views.Slice(arr, idx)
      assert np.array_equal(v.glom(), x[idx]), (shape, hint, idx)
      t = views.Transpose(arr)
      assert np.array_equal(t.glom(), x.T)
      assert np.array_equal(views.Slice(t, tuple(reversed(idx))).glom(), x.T[tuple(reversed(idx))])
      # reshape of the (contiguous) base to a random factorisation
      n = int(np.prod(shape))
      divs = [d for d in range(1, n + 1) if n % d == 0]
      a = int(divs[rng.randint(len(divs))])
      for new in ((n,), (a, n // a), (n // a, a, 1)):
        r = views.Reshape(arr, new)
        assert np.array_equal(r.glom(), x.reshape(new)), (shape, new)
        for ex in r.tiles:                                  # every view tile on its own, incl. partial-row extents
          assert np.array_equal(r.fetch(ex).numpy(), x.reshape(new)[ex.to_slice()])
    # shapes produced by Expr.__getitem__ (NumPy semantics for integers and newaxis)
    e = sp.ndarray((4, 5, 6), dtype=np.float32)
    ref = np.zeros((4, 5, 6), np.float32)
    for idx in (1, -1, np.index_exp[:, 2], np.index_exp[1, :, 4], np.index_exp[1:3], np.index_exp[:, sp.newaxis, 2:4],
                np.index_exp[0, 1], np.index_exp[:, :, -1]):
      np_idx = tuple(np.newaxis if i is sp.newaxis else i for i in idx) if isinstance(idx, tuple) else idx
      assert tuple(e[idx].shape) == ref[np_idx].shape, idx
  finally:
    blob_ctx._global_ctx[0] = old
    blob_ctx._local.ctx = old


def test_fusion_counts_distinct_operands():
  """A chain that uses three arrays ten times is ONE kernel (an array used several times is one operand, Python scalars
  are immediates); nine distinct arrays are not (SP_MAX_OPERANDS = 8)."""
  from spartan_b200 import blob_ctx
  old = blob_ctx._global_ctx[0]
  blob_ctx.set(blob_ctx.BlobCtx(0, 1, 'cpu'))
  try:
    x, y, z = (sp.ndarray((64, 64), dtype=np.float32) for _ in range(3))
    e = (((x + y) * z - x) * 0.5 + y * y - z * 2 + 1).optimized()
    assert isinstance(e, sp.MapExpr) and not any(isinstance(c, sp.MapExpr) for c in e.children)
    ids = set(c.expr_id for c in e.children if isinstance(c, sp.NdArrayExpr))
    assert ids == {x.expr_id, y.expr_id, z.expr_id}
    many = [sp.ndarray((8, 8), dtype=np.float32) for _ in range(9)]
    s = many[0]
    for m in many[1:]:
      s = s + m
    opt = s.optimized()
    assert any(isinstance(c, sp.MapExpr) for c in opt.children)      # split: the ninth operand does not fit
  finally:
    blob_ctx._global_ctx[0] = old
    blob_ctx._local.ctx = old


def test_tile_file_bytes_match_golden(tmp_path):
  """tests/golden/fio_vectors.json (made by tests/golden/make_fio_vectors.py): the product writes exactly these bytes and
  both loaders read them back."""
  import base64, json
  from spartan_b200 import blob_ctx
  gold = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'fio_vectors.json')))
  old = blob_ctx._global_ctx[0]
  blob_ctx.set(blob_ctx.BlobCtx(0, 1, 'cpu'))
  spartan_oracle.initialize(1)
  try:
    for case in gold['cases']:
      shape, hint, dtype = tuple(case['shape']), tuple(case['tile_hint']), np.dtype(case['dtype'])
      x = np.array(case['data'], dtype=dtype).reshape(shape)
      arr = pdist.create(shape, dtype, tile_hint=hint)
      arr.update(pex.from_shape(shape), x)
      assert sp.save(arr, case['prefix'], str(tmp_path), False) is True
      d = os.path.join(str(tmp_path), case['prefix'])
      assert sorted(os.listdir(d)) == sorted(case['files'])
      for fn, b64 in case['files'].items():
        assert open(os.path.join(d, fn), 'rb').read() == base64.b64decode(b64), fn
      # a directory holding only the golden bytes loads through both loaders
      g = os.path.join(str(tmp_path), 'golden_' + case['prefix'], case['prefix'])
      os.makedirs(g)
      for fn, b64 in case['files'].items():
        open(os.path.join(g, fn), 'wb').write(base64.b64decode(b64))
      assert np.array_equal(sp.load(case['prefix'], os.path.dirname(g), False).glom(), x)
      assert np.array_equal(spartan_oracle.fio.load(case['prefix'], os.path.dirname(g), False).glom(), x)
  finally:
    blob_ctx._global_ctx[0] = old
    blob_ctx._local.ctx = old


def test_map2_join_extents_match_oracle_join_mapper():
  """join_mapper's strip arithmetic (map.py:248-272, extent.pyx:501-570) is integer work: the extents the device map2
  fetches must equal the ones the oracle's join_mapper hands to the tile function, for every tile."""
  import spartan_oracle
  from spartan_oracle import expr as oexpr, distarray as odist
  from spartan_b200.expr.map2 import join_extents
  from spartan_b200.array import distarray, extent
  for workers in (1, 3, 8):
    spartan_oracle.initialize(workers)
    for shapes, hints, axes in [(((100, 60), (60, 77)), ((25, 60), (60, 20)), (1, 0)),
                                (((64, 30), (30, 50)), ((64, 10), (6, 50)), (1, 0)),
                                (((50, 8), (50,)), ((7, 8), (9,)), (0, 0)),
                                (((33, 20), (33, 20)), ((10, 20), (33, 5)), ())]:
      oarrs = [odist.create(s, np.float32, tile_hint=h) for s, h in zip(shapes, hints)]
      seen = []
      oexpr_fn = lambda extents, tiles, **kw: seen.append(extents if isinstance(extents, list) else [extents] * len(shapes)) or []
      for ex in oarrs[0].tiles:
        oexpr.join_mapper(ex, oarrs, axes, oexpr_fn, None, None)
      got = []
      for ex in distarray.compute_extents(shapes[0], hints[0], workers):
        j = join_extents(extent.create(ex.ul, ex.lr, shapes[0]), list(shapes), axes)
        if j is not None:
          got.append(j)
      assert len(got) == len(seen)
      for a, b in zip(got, seen):
        assert [(e.ul, e.lr, tuple(e.array_shape)) for e in a] == [(e.ul, e.lr, tuple(e.array_shape)) for e in b]


def test_automatic_tiling_minimises_nvlink_bytes():
  """AutomaticTiling (optimize.py:459-1054) with this backend's cost -- bytes over NVLink: the chosen tilings must be a
  global minimum of the cost model (checked against brute force), must remove avoidable traffic in the textbook cases,
  and the rewritten DAG must carry the matching tile_hints."""
  import itertools
  from spartan_b200.expr import tiling
  from spartan_b200.expr.base import Expr
  from spartan_b200.expr.ndarray import NdArrayExpr
  from spartan_b200.expr.dot import DotExpr
  W = 8
  n = 4096
  # 1. a reduction over axis 0 of a fused chain: row tiling needs an all-reduce, column tiling does not
  x = sp.ndarray((n, n), dtype=np.float32); y = sp.ndarray((n, n), dtype=np.float32)
  e = (x * 2 + y).sum(axis=0)
  p = tiling.AutomaticTiling(W)
  p._build(e)
  choice, cost, _ = p.solve()
  assert cost == 0 and set(choice.values()) == {tiling.COL}
  # ... and over axis 1 the other way round
  p = tiling.AutomaticTiling(W); p._build((x * 2 + y).sum(axis=1))
  choice, cost, _ = p.solve()
  assert cost == 0 and set(choice.values()) == {tiling.ROW}
  # 2. operands of one map end up tiled alike (an existing array fixes the choice of the free one)
  fixed = sp.ndarray((n, n), dtype=np.float32, tile_hint=(n, n // W))
  p = tiling.AutomaticTiling(W); p._build(fixed + x)
  choice, cost, _ = p.solve()
  assert cost == 0 and list(choice.values()) == [tiling.COL]
  # 3. dot: a column-tiled C with B tiled by columns moves (W-1)|A| bytes and nothing else
  a = sp.ndarray((n, n), dtype=np.float32); b = sp.ndarray((n, 2 * n), dtype=np.float32)
  d = sp.dot(a, b)
  p = tiling.AutomaticTiling(W); p._build(d)
  choice, cost, til = p.solve()
  assert cost == (W - 1) * n * n * 4          # |A| < |B|: gather A, keep B's columns in place
  assert til[d.expr_id] != tiling.ROW and til[b.expr_id] != tiling.ROW
  # 4. brute force over every assignment of a mixed DAG: the pass returns a global minimum
  u = sp.ndarray((n, n), dtype=np.float32); v = sp.ndarray((n, n), dtype=np.float32); w = sp.ndarray((n, n), dtype=np.float32)
  dag = (sp.dot(u, v) + sp.transpose(w)).sum(axis=1)
  p = tiling.AutomaticTiling(W); p._build(dag)
  free = p.free_nodes()
  choice, cost, _ = p.solve()
  best = min(p.evaluate(dict(zip(free, combo)))[0] for combo in itertools.product((0, 1, 2), repeat=len(free)))
  assert cost == best
  # coordinate descent (used beyond EXHAUSTIVE_LIMIT free nodes) must not be worse than a fixed default
  old = tiling.EXHAUSTIVE_LIMIT
  try:
    tiling.EXHAUSTIVE_LIMIT = 0
    p2 = tiling.AutomaticTiling(W); p2._build(dag)
    _, cost2, _ = p2.solve()
    assert cost2 <= p2.evaluate({})[0]
  finally:
    tiling.EXHAUSTIVE_LIMIT = old
  # 5. the rewritten DAG carries the hints (and only free nodes change)
  out = tiling.AutomaticTiling(W).visit((x * 2 + y).sum(axis=0))
  hints = []
  def walk(ex):
    if isinstance(ex, NdArrayExpr):
      hints.append(ex.tile_hint)
    elif isinstance(ex, Expr):
      for dep in ex.dependencies().values():
        walk(dep)
      if hasattr(ex, 'children') and ex.children is not None:
        for c in ex.children:
          walk(c)
  walk(out)
  assert hints and all(h == (n, n // W) for h in hints), hints
  assert tiling.hint_for((100, 60), tiling.ROW, 8) == (13, 60) and tiling.hint_for((100, 60), tiling.COL, 8) == (100, 8)
  assert tiling.hint_for((100, 60), tiling.BLOCK, 8) == (13, 8)
  # a single rank has nothing to optimise
  e1 = (x * 2 + y).sum(axis=0)
  assert tiling.AutomaticTiling(1).visit(e1) is e1


def test_dot_arrival_groups_cover_every_segment_once():
  """The passes of the multi-GPU dot (segments grouped by expected arrival) must be a partition of the segment order,
  in order; with free passes the local segment goes alone, with expensive passes everything is one launch."""
  from spartan_b200.expr.dot import _arrival_groups
  for n in (1, 2, 3, 4, 8):
    for t_push, t_seg, t_pass in [(0.87, 2.1, 0.0), (0.87, 2.1, 0.5), (3.6, 36.0, 2.4), (3.0, 1.0, 0.1), (0.0, 1.0, 0.2),
                                  (1.0, 1e-9, 0.0), (0.87, 2.1, 50.0)]:
      g = _arrival_groups(n, t_push, t_seg, t_pass)
      assert [j for grp in g for j in grp] == list(range(n))
      assert 1 <= len(g) <= 4
      if t_pass >= 50.0:
        assert len(g) == 1
  assert _arrival_groups(8, 0.87, 2.1, 0.0)[0] == [0]
  assert _arrival_groups(8, 0.87, 2.1, 0.6) == [[0], [1, 2], [3, 4, 5, 6, 7]]
  assert _arrival_groups(2, 3.6, 36.0, 5.0) == [[0, 1]]          # a pass dearer than the stall it avoids
  assert _arrival_groups(2, 3.6, 36.0, 2.4) == [[0], [1]]


def test_location_fill_uses_global_positions_for_every_tiling(monkeypatch):
  """rand / arange tiles are filled from the element's GLOBAL position: a tile that spans full rows is one flat fill at
  its offset, any other tile (column blocks included) a 2-D fill with the array's row pitch.  (A full-height column tile
  once took the flat path because the module's own `all` shadowed the builtin.)"""
  import torch
  from spartan_b200 import device_ops, blob_ctx
  old = blob_ctx.get()
  try:
    blob_ctx.set(blob_ctx.BlobCtx(1, 2, torch.device('cpu')))      # pose as rank 1 of 2; nothing is launched
    calls = []
    monkeypatch.setattr(device_ops, 'fill', lambda t, kind, a=0.0, b=0.0, seed=0, offset=0: calls.append(('flat', tuple(t.shape), offset)))
    monkeypatch.setattr(device_ops, 'fill2d', lambda t, kind, a=0.0, b=0.0, seed=0, offset=0, pitch=0:
                        calls.append(('2d', tuple(t.shape), offset, pitch)))
    sp.rand(384, 640, seed=1, dtype=np.float32).evaluate()
    assert calls == [('flat', (192 * 640,), 192 * 640)]
    del calls[:]
    sp.rand(384, 640, seed=1, dtype=np.float32, tile_hint=(384, 320)).evaluate()
    assert calls == [('2d', (384, 320), 320, 640)]
    del calls[:]
    sp.arange((384, 640), dtype=np.float32, tile_hint=(96, 160)).evaluate()
    assert calls and all(c[0] == '2d' and c[3] == 640 for c in calls)
  finally:
    blob_ctx.set(old)


def test_prepared_operand_cache_epochs_and_eviction():
  """PreparedCache (device_ops): an entry is fresh only within the data epoch it was filled in; a stale entry hands its
  buffers back for re-filling; least-recently-used entries go when the byte budget is exceeded; dropping an array's
  serial removes all of its entries."""
  from spartan_b200.device_ops import PreparedCache
  old = sp.FLAGS.dot_prepared_cache_bytes
  try:
    sp.FLAGS.dot_prepared_cache_bytes = 100
    c = PreparedCache()
    assert c.lookup((1, 'a'), 0) == (None, False)
    c.store((1, 'a'), 'A1', 40, 0)
    assert c.lookup((1, 'a'), 0) == ('A1', True)
    assert c.lookup((1, 'a'), 1) == ('A1', False)          # the array may have changed: same buffers, re-fill
    c.store((1, 'a'), 'A1', 40, 1)
    c.store((2, 'b'), 'B2', 40, 1)
    c.store((3, 'a'), 'A3', 40, 1)                          # 120 > 100: the least recently used entry goes
    assert (1, 'a') not in c.entries and set(c.entries) == {(2, 'b'), (3, 'a')}
    c.lookup((2, 'b'), 1)                                   # touch: (3, 'a') is now the oldest
    c.store((4, 'a'), 'A4', 40, 1)
    assert set(c.entries) == {(2, 'b'), (4, 'a')}
    c.store((4, 'b'), 'B4', 10, 1)
    c.drop(4)
    assert set(c.entries) == {(2, 'b')}
    c.store((5, 'huge'), 'H', 1000, 1)                      # larger than the budget: kept alone rather than thrashing
    assert set(c.entries) == {(5, 'huge')}
    assert c.hits >= 2 and c.misses >= 2
  finally:
    sp.FLAGS.dot_prepared_cache_bytes = old


def test_tiling_memo_hands_out_independent_tables():
  """compute_extents keeps recent answers (the tiling is a pure function of shape / hint / shards): a caller that edits
  the table it got must not change what the next caller sees, and the memo must not confuse hints or shard counts."""
  a = pdist.compute_extents((64, 48), (16, 48), 4)
  first = list(a.items())
  a.popitem()
  a[pex.create((0, 0), (1, 1), (64, 48))] = 99
  b = pdist.compute_extents((64, 48), (16, 48), 4)
  assert list(b.items()) == first and b is not a
  assert list(pdist.compute_extents((64, 48), (16, 48), 2).values()) != [w for _, w in first] or \
    len(set(w for _, w in first)) <= 2
  assert len(pdist.compute_extents((64, 48), (32, 48), 4)) == 2
  assert dict((e.to_tuple(), w) for e, w in pdist.compute_extents((64, 48), None, 4).items()) == \
    dict((e.to_tuple(), w) for e, w in odist.compute_extents((64, 48), None, 4).items())
  # extents hash by their upper-left corner (extent.pyx:93-94) and compare by both corners
  e1, e2 = pex.create((0, 0), (4, 4), (8, 8)), pex.create((0, 0), (4, 8), (8, 8))
  assert hash(e1) == hash(e2) == hash((0, 0)) and e1 != e2 and len({e1: 1, e2: 2}) == 2


def test_passes_keep_nodes_they_do_not_change():
  """A pass that rewrites nothing below a node returns the node itself (no copy of the DAG per pass); a fused node keeps
  the id of the node it replaces, so cached results stay addressable (base.py:49-68)."""
  x = sp.from_numpy(np.zeros((4, 4), np.float32))
  d = sp.dot(x, x)
  assert d.optimized() is d
  e = (x * 2 + x).sum(axis=0)
  o = e.optimized()
  assert o is not e and o.expr_id == e.expr_id and e.optimized() is o
