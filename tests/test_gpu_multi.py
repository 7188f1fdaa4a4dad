"""Two-rank NCCL tests (run when the box has >= 2 GPUs): the SPMD paths that need an exchange -- the dot
all-gather fast path, the general rectangle-fetch path, and the all-reduce combiner of reductions -- against
NumPy on the same seeded inputs."""
import os
import socket
import sys
import traceback

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
  s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q, use_peer):
  try:
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1',
                      MASTER_PORT=str(port), SPARTAN_PEER='1' if use_peer else '0')
    if ROOT not in sys.path:
      sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import spartan_b200 as sp
    ctx = sp.initialize()
    rng = np.random.default_rng(0)
    # dot, regular placement (all-gather of prepared A slabs) and an irregular one (rectangle fetches)
    for (M, K, N, hint) in [(1024, 1024, 1024, (256, 256)), (512, 768, 640, (128, 128)), (300, 500, 260, (100, 130))]:
      a = rng.standard_normal((M, K), dtype=np.float32); b = rng.standard_normal((K, N), dtype=np.float32)
      ref = np.dot(a.astype(np.float64), b.astype(np.float64))
      for prec, tol in [('bf16x3', 1e-5), ('tf32x3', 1e-5), ('simt', 1e-5)]:
        sp.FLAGS.dot_precision = prec
        got = sp.dot(sp.from_numpy(a, tile_hint=hint), sp.from_numpy(b, tile_hint=hint), tile_hint=hint).glom()
        err = np.abs(got - ref).max() / np.abs(ref).max()
        assert err <= tol, (M, K, N, hint, prec, err)
    # the regular placement goes over peer memory (copy-engine pushes + gated segments) when CUDA IPC works
    peer_ok = ctx.peer.available()
    assert peer_ok or not use_peer or os.environ.get('SPARTAN_ALLOW_NO_IPC'), 'CUDA IPC between the two ranks failed'
    print('rank %d: peer memory %s' % (rank, 'available' if peer_ok else 'UNAVAILABLE (NCCL paths)'), flush=True)
    from spartan_b200.expr.base import lazify as _lz
    sp.FLAGS.dot_precision = 'bf16x3'
    # K/W not a multiple of the k-block (zero-padded prepared operands), repeated evaluations over the same arrays
    # (prepared-operand cache + the two halves of the gather buffer), then an in-place update (cache invalidation)
    for (M, K, N, ha, hb, hc) in [(512, 1000, 512, (256, 500), (500, 256), (256, 256)),
                                  (384, 2048, 768, (128, 256), (256, 128), (128, 128))]:
      a = rng.standard_normal((M, K), dtype=np.float32); b = rng.standard_normal((K, N), dtype=np.float32)
      Ad = sp.from_numpy(a, tile_hint=ha).evaluate(); Bd = sp.from_numpy(b, tile_hint=hb).evaluate()
      ref = a.astype(np.float64) @ b.astype(np.float64)
      outs = []
      for _ in range(4):
        outs.append(sp.dot(_lz(Ad), _lz(Bd), tile_hint=hc).glom())
      for o in outs:
        assert np.abs(o - ref).max() <= 1e-5 * np.abs(ref).max()
        assert np.array_equal(o, outs[0])
      a2 = rng.standard_normal((M, K), dtype=np.float32)
      Ad.update(sp.extent.from_shape(Ad.shape), a2)
      got = sp.dot(_lz(Ad), _lz(Bd), tile_hint=hc).glom()
      ref2 = a2.astype(np.float64) @ b.astype(np.float64)
      assert np.abs(got - ref2).max() <= 1e-5 * np.abs(ref2).max()
      sp.FLAGS.dot_prepared_cache = False
      again = sp.dot(_lz(Ad), _lz(Bd), tile_hint=hc).glom()
      sp.FLAGS.dot_prepared_cache = True
      assert np.array_equal(again, got)
    # exchanges of DIFFERENT shapes queued back to back with no host synchronisation in between: each exchange owns a fixed
    # half of the symmetric buffer, so a fast rank's next push can never land in what a slow peer is still reading
    pairs = []
    for (M, K, N) in [(1024, 2048, 512), (256, 512, 1024), (768, 1024, 256)]:
      a = rng.standard_normal((M, K), dtype=np.float32); b = rng.standard_normal((K, N), dtype=np.float32)
      pairs.append((sp.from_numpy(a, tile_hint=(M, K // 2)).evaluate(), sp.from_numpy(b, tile_hint=(K, N // 2)).evaluate(),
                    a.astype(np.float64) @ b.astype(np.float64), (M, N // 2)))
    results = []
    for rep in range(3):
      for Ad, Bd, ref, hc in pairs:
        results.append((sp.dot(_lz(Ad), _lz(Bd), tile_hint=hc).evaluate(), ref))
    for Cd, ref in results:
      got = Cd.glom()
      assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max()
    del results, pairs
    # fused map + reduce: partials combined by ncclAllReduce
    x = rng.random((1024, 2048), dtype=np.float32); y = rng.random((1024, 2048), dtype=np.float32)
    for hint in [(128, 2048), (256, 512), None]:
      got = (sp.from_numpy(x, tile_hint=hint) * 2 + sp.from_numpy(y, tile_hint=hint)).sum(axis=0).optimized().glom()
      np.testing.assert_allclose(got, (x.astype(np.float64) * 2 + y).sum(axis=0), rtol=1e-5)
      got = sp.from_numpy(x, tile_hint=hint).max(axis=1).glom()
      assert np.array_equal(got, x.max(axis=1))
    xi = rng.integers(-50, 50, size=(333, 77)).astype(np.int64)
    assert sp.from_numpy(xi).sum().glom() == xi.sum()
    # sparse column strips x vector: partial y of every rank combined by reduce-scatter (default result tiling) or
    # all-reduce (any other tiling); k-means: rows sharded, sums / counts all-reduced
    import scipy.sparse
    for n_, strip, hint_y in [(4000, 500, None), (4000, 1000, (1333, 1)), (3001, 700, None)]:
      nnz = 6 * n_
      m = scipy.sparse.coo_matrix((rng.random(nnz, dtype=np.float32), (rng.integers(0, n_, nnz), rng.integers(0, n_, nnz))),
                                  shape=(n_, n_))
      xv = rng.random((n_, 1), dtype=np.float32)
      got = sp.dot(sp.sparse.from_scipy(m, strip_width=strip), sp.from_numpy(xv, tile_hint=(strip, 1)), tile_hint=hint_y).glom()
      want = m.tocsr().astype(np.float64).dot(xv.astype(np.float64))
      np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-6)
    pts = rng.random((6000, 64), dtype=np.float32)
    c_init = pts[:10].copy()
    cen, lab = sp.KMeans(n_clusters=10, n_iter=2).fit(sp.from_numpy(pts, tile_hint=(1500, 64)), centers=c_init)
    c64 = c_init.astype(np.float64)
    for _ in range(2):
      d2 = ((pts[:, None, :].astype(np.float64) - c64[None]) ** 2).sum(-1)
      l64 = d2.argmin(1)
      c64 = np.stack([pts[l64 == j].astype(np.float64).mean(0) for j in range(10)])
    assert (lab.glom() == l64).mean() > 0.999
    np.testing.assert_allclose(cen, c64, rtol=1e-4, atol=1e-5)
    # AutomaticTiling (tile_hints chosen to minimise NVLink bytes) must not change any value: the device generator is
    # indexed by global element position, so the arrays are the same under any tiling
    outs = []
    for auto in (False, True):
      sp.FLAGS.opt_auto_tiling = auto
      e1 = (sp.rand(384, 640, seed=21, dtype=np.float32) * 2 + sp.rand(384, 640, seed=22, dtype=np.float32)).sum(axis=0)
      e2 = sp.dot(sp.rand(256, 384, seed=23, dtype=np.float32), sp.rand(384, 512, seed=24, dtype=np.float32))
      outs.append((e1.optimized().glom(), e2.optimized().glom()))
    sp.FLAGS.opt_auto_tiling = False
    np.testing.assert_allclose(outs[0][0], outs[1][0], rtol=1e-6)
    np.testing.assert_allclose(outs[0][1], outs[1][1], rtol=1e-5, atol=1e-5)
    # operands with different tilings: pieces travel point-to-point to the owner of each output tile
    got = (sp.from_numpy(x, tile_hint=(128, 2048)) + sp.from_numpy(y, tile_hint=(1024, 256))).glom()
    assert np.array_equal(got, x + y)
    assert sp.ones((4096, 4096)).sum().glom() == 16777216.0
    # dot over host operands with the upload pipelined against the contraction (per-strip prepare + all-gather + GEMM
    # on three streams); checked against float64 and the resident multi-GPU path, through glom and block-wise read-back
    sp.FLAGS.dot_precision = 'bf16x3'
    for (M, K, N, hint, strip) in [(1024, 1024, 1024, (256, 256), 256), (768, 512, 1024, (128, 128), 200),
                                   (640, 1000, 512, None, 256), (1024, 1024, 1024, (256, 256), 256)]:
      a = rng.standard_normal((M, K), dtype=np.float32); b = rng.standard_normal((K, N), dtype=np.float32)
      ha, hb = (hint, hint) if hint is not None else ((128, 500), (500, 256))     # K/W = 500: padded k-blocks
      hint = hint or (128, 256)
      sp.FLAGS.dot_stream_host_operands = False
      want = sp.dot(sp.from_numpy(a, tile_hint=ha), sp.from_numpy(b, tile_hint=hb), tile_hint=hint).glom()
      sp.FLAGS.dot_stream_host_operands = True; sp.FLAGS.dot_stream_min_bytes = 0; sp.FLAGS.dot_stream_strip = strip
      c = sp.dot(sp.from_numpy(a, tile_hint=ha), sp.from_numpy(b, tile_hint=hb), tile_hint=hint).evaluate()
      assert c.block_events, 'the streamed multi-rank path did not run'
      out = torch.zeros((M, N), dtype=torch.float32, pin_memory=True)
      nbytes = c.read_local_into(out.numpy())
      torch.cuda.current_stream().synchronize()
      assert nbytes == M * N * 4 // world
      full = c.glom()
      # (the resident path contracts its own K segment first, the streamed one walks the source ranks in order:
      #  same products, different fp32 summation order)
      ref = a.astype(np.float64) @ b.astype(np.float64)
      assert np.abs(full - ref).max() <= 1e-5 * np.abs(ref).max()
      assert np.abs(full - want).max() <= 2e-6 * np.abs(ref).max()
      mine = np.zeros((M, N), bool)
      for ex, tid in c.tiles.items():
        if tid.worker == rank:
          mine[ex.to_slice()] = True
      assert np.array_equal(out.numpy()[mine], full[mine])
    sp.FLAGS.dot_stream_min_bytes = 256 << 20; sp.FLAGS.dot_stream_strip = 4096
    # a captured evaluation (kernels + ncclAllReduce in one CUDA graph) replayed on updated inputs
    from spartan_b200.expr.base import lazify
    X = sp.from_numpy(x, tile_hint=(128, 2048)).evaluate(); Y = sp.from_numpy(y, tile_hint=(128, 2048)).evaluate()
    rep = sp.replayable(lambda: (lazify(X) * 2 + lazify(Y)).sum(axis=0).optimized().evaluate())
    for _ in range(3):
      got = rep().glom()
    np.testing.assert_allclose(got, (x.astype(np.float64) * 2 + y).sum(axis=0), rtol=1e-5)
    x2 = rng.random((1024, 2048), dtype=np.float32)
    X.update(sp.extent.from_shape(X.shape), x2)
    np.testing.assert_allclose(rep().glom(), (x2.astype(np.float64) * 2 + y).sum(axis=0), rtol=1e-5)
    if peer_ok:
      assert ctx.peer.gate_timeouts() == 0, 'a gated GEMM segment timed out waiting for its operand strip'
    dist.barrier()
    q.put((rank, 'ok'))
  except Exception:
    q.put((rank, traceback.format_exc()))


@pytest.mark.parametrize('use_peer', [True, False], ids=['peer-memory', 'nccl-only'])
def test_two_ranks_nccl(use_peer):
  """use_peer=False forces the NCCL all-gather implementations of the dot exchange (the fallback when CUDA IPC is not
  available between the ranks)."""
  if torch.cuda.device_count() < 2:
    pytest.skip('needs 2 GPUs')
  world = 2
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, world, port, q, use_peer)) for r in range(world)]
  for p in procs: p.start()
  results = [q.get(timeout=600) for _ in range(world)]
  for p in procs: p.join(timeout=60)
  for rank, msg in results:
    assert msg == 'ok', 'rank %d failed:\n%s' % (rank, msg)
