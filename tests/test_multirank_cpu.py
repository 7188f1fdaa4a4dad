"""world_size-2 checks of the SPMD host logic over gloo on CPU (no kernels run): every rank derives the same
extent -> TileId tables, owns the tiles round-robin placement says it owns, from_numpy/glom round-trip
through the slab upload + broadcast path, and the cross-rank combiner wrapper reduces with the right op."""
import os
import socket
import sys
import traceback

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
  s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
  try:
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1',
                      MASTER_PORT=str(port))
    for p in (ROOT, os.path.join(ROOT, 'oracle')):
      if p not in sys.path:
        sys.path.insert(0, p)
    import torch.distributed as dist
    import spartan_b200 as sp
    from spartan_b200 import comm, _lib
    from spartan_b200.array import distarray
    from spartan_oracle import distarray as odist
    ctx = sp.initialize(device=torch.device('cpu'))
    assert (ctx.worker_id, ctx.num_workers) == (rank, world)

    # 1. placement: identical tables on every rank, equal to the oracle's round-robin assignment
    for shape, hint in [((64, 48), (16, 16)), ((50, 50, 50), None), ((10,), None), ((32768, 32768), (4096, 4096))]:
      arr = distarray.create(shape, np.float32, tile_hint=hint) if np.prod(shape) < 10 ** 6 else None
      ext = distarray.compute_extents(shape, hint, world)
      ref = odist.compute_extents(shape, hint, world)
      assert [(e.ul, e.lr, w) for e, w in ext.items()] == [(e.ul, e.lr, w) for e, w in ref.items()]
      if arr is not None:
        table = [(e.ul, e.lr, t.worker, t.id) for e, t in arr.tiles.items()]
        gathered = [None] * world
        dist.all_gather_object(gathered, table)
        assert all(g == gathered[0] for g in gathered), 'ranks disagree on the tile table'
        for e, t in arr.tiles.items():
          assert t.worker == ext[e] % world
          assert (t in ctx._blobs) == (t.worker == rank)

    # 2. from_numpy -> glom across ranks (slab upload on the owner, broadcast on glom)
    rng = np.random.RandomState(0)
    for shape, hint in [((37, 53), (10, 53)), ((37, 53), (37, 7)), ((64, 64), (16, 16)), ((9,), (2,))]:
      x = rng.randn(*shape).astype(np.float32)
      got = sp.from_numpy(x, tile_hint=hint).glom()
      assert np.array_equal(got, x), (shape, hint)

    # 3. the cross-rank combiner = all-reduce with the matching op
    for op, fn in [(_lib.SP_RED_SUM, np.add), (_lib.SP_RED_MIN, np.minimum), (_lib.SP_RED_MAX, np.maximum),
                   (_lib.SP_RED_PROD, np.multiply)]:
      mine = torch.tensor([1.0 + rank, 5.0 - rank, 2.0], dtype=torch.float64)
      comm.allreduce(mine, op)
      ref = np.array([1.0, 5.0, 2.0])
      for r in range(1, world):
        ref = fn(ref, np.array([1.0 + r, 5.0 - r, 2.0]))
      assert np.array_equal(mine.numpy(), ref), (op, mine, ref)
    flags = torch.tensor([rank == 0, True, False])
    comm.allreduce(flags, _lib.SP_RED_ALL); assert flags.tolist() == [False, True, False]
    flags = torch.tensor([rank == 0, True, False])
    comm.allreduce(flags, _lib.SP_RED_ANY); assert flags.tolist() == [True, True, False]

    # 3b. tile files: every rank writes / reads only the tiles it owns, rank 0 writes the array-wide file; the result
    # loads identically on both ranks and through the oracle's restatement of the reference loader
    import tempfile
    from spartan_oracle import fio as ofio
    import spartan_oracle
    spartan_oracle.initialize(world)
    base = [tempfile.mkdtemp(prefix='sp_fio_') if rank == 0 else None]
    dist.broadcast_object_list(base, src=0)
    for iszip in (False, True):
      x = rng.randn(70, 33).astype(np.float32)
      arr = sp.from_numpy(x, tile_hint=(16, 33)).evaluate()
      assert sp.save(arr, 'w2', base[0], iszip) is True
      names = sorted(n for n in os.listdir(os.path.join(base[0], 'w2')) if n.endswith('bz2') == iszip and 'dist' not in n)
      assert len(names) == len(arr.tiles), names              # one file per tile, written by its owner
      assert np.array_equal(sp.load('w2', base[0], iszip).glom(), x)
      assert np.array_equal(ofio.load('w2', base[0], iszip).glom(), x)
      dist.barrier()

    # 3c. views: both ranks derive the same view tile tables, and every view tile lives on the rank of the base tile
    from spartan_b200.array import views
    arr = distarray.create((64, 48), np.float32, tile_hint=(16, 16))
    for v in (views.Slice(arr, np.index_exp[5:40, 10:33]), views.Transpose(arr), views.Reshape(arr, (48, 64)),
              views.Reshape(arr, (64, 48, 1))):
      table = [(e.ul, e.lr, tuple(e.array_shape), t.worker) for e, t in v.tiles.items()]
      gathered = [None] * world
      dist.all_gather_object(gathered, table)
      assert all(g == gathered[0] for g in gathered), 'ranks disagree on a view tile table'
      assert all(0 <= w < world for _, _, _, w in table)

    # 4. compute without a GPU must fail loudly on every rank, not fall back
    try:
      (sp.from_numpy(np.ones((4, 4), np.float32)) + 1).glom()
      raise AssertionError('compute on CPU did not fail')
    except sp.SpartanError:
      pass
    dist.barrier()
    q.put((rank, 'ok'))
  except Exception:
    q.put((rank, traceback.format_exc()))


def test_world_size_2_gloo():
  world = 2
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
  for p in procs: p.start()
  results = [q.get(timeout=240) for _ in range(world)]
  for p in procs: p.join(timeout=60)
  for rank, msg in results:
    assert msg == 'ok', 'rank %d failed:\n%s' % (rank, msg)
