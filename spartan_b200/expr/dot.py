"""``dot`` -- dense contraction of tiled arrays (reference: spartan/expr/dot.py:95-299).

The reference routes 2-D x 2-D products through map2 / outer joins: every A tile is re-partitioned
into a K strip, the matching B strip is fetched, ``tiles[0].dot(tiles[1])`` produces a full-size
rank-k partial of C and the partials are np.add-merged on the owners of the C tiles
(dot.py:195-238, map.py:243-286).  That moves #tiles x M x N output bytes and, on grid tilings,
contracts the wrong K indices.

B200-native plan -- owner computes: each rank produces exactly the C tiles it owns,
C[R, Cc] = A[R, :] . B[:, Cc], in as few launches as the placement allows (one per contiguous run of
owned rows x contiguous run of owned columns; ONE launch for a single GPU or for column-block
placement), with the K loop walking "segments" -- one per (A strip, B strip) pair -- inside a single
tcgen05 kernel so the accumulator never leaves the SM between strips.  The only data exchange is the
operand strips a rank does not own: an all-gather of the A slabs over NVLink when the placement is
regular, point-to-point rectangle fetches otherwise.  No output partial ever crosses a link and the
summation order is fixed.
"""
import numpy as np
import torch

from .. import blob_ctx, device_ops
from ..array import distarray, extent
from ..config import FLAGS
from .._lib import SpartanError, SP_GEMM_MAX_SEGMENTS
from .base import Expr, lazify, eval_cache
from .write_array import WriteArrayExpr


def _runs(intervals):
  """Merges sorted disjoint [lo, hi) intervals that touch."""
  out = []
  for lo, hi in intervals:
    if out and out[-1][1] == lo:
      out[-1] = (out[-1][0], hi)
    else:
      out.append((lo, hi))
  return out


def _owned_runs(array, worker):
  """Per-axis contiguous runs of the tiles of ``array`` owned by ``worker`` (2-D arrays), or None when
  the owned tiles are not a cartesian product of row and column intervals."""
  local = [ex for ex, tid in array.tiles.items() if tid.worker == worker]
  if not local:
    return [], []
  axes = distarray._product_layout(local, 2)
  if axes is None:
    return None
  return _runs(axes[0]), _runs(axes[1])


def _regular_layout(av, bv, target, W, M, K):
  """The placement the multi-GPU fast paths need: every rank owns whole column blocks of A, B and C (full height),
  equally wide, and the B columns a rank needs for its C columns are its own -- e.g. round-robin over an evenly
  divisible tile grid.  Returns per-rank (A axes, B axes, C runs) or None."""
  a_axes, c_runs, b_axes = [], [], []
  for w in range(W):
    la = [ex for ex, tid in av.tiles.items() if tid.worker == w]
    lb = [ex for ex, tid in bv.tiles.items() if tid.worker == w]
    pa = distarray._product_layout(la, 2)
    pb = distarray._product_layout(lb, 2)
    rc = _owned_runs(target, w)
    if pa is None or pb is None or rc is None or not rc[0]:
      return None
    if _runs(pa[0]) != [(0, M)] or _runs(pb[0]) != [(0, K)] or rc[0] != [(0, M)]:
      return None                       # every rank must hold full columns of A and B and full-height C blocks
    if a_axes and sum(b - a for a, b in pa[1]) != sum(b - a for a, b in a_axes[0][1]):
      return None                       # equal slab widths (all-gather needs equal contributions)
    # the B columns this rank needs (= its C columns) must be local to it
    bcols = pb[1]
    for c0, c1 in rc[1]:
      if not any(a <= c0 and c1 <= b for a, b in _runs(bcols)):
        return None
    a_axes.append(pa); b_axes.append(pb); c_runs.append(rc)
  return a_axes, b_axes, c_runs


_PUSH_BYTES_PER_S = 600e9       # copy-engine push over NVLink 5, one stream (measured: 3.76 GB in 6.1 ms)
_DOT_FLOPS_PER_S = 480e12       # algorithmic rate of the bf16x3 contraction on one B200


_PASS_BYTES_PER_S = 2.0e12      # rate at which an accumulating pass re-reads and re-writes its C block
_PASS_FIXED_S = 0.3e-3          # ramp / tail of one more launch


def _arrival_groups(n_segments, t_push, t_segment, t_pass=0.0, max_groups=4):
  """Positions 0..n-1 of the segment order (0 = local, j = the j-th push to land, at j * t_push) split into contiguous
  passes.  A pass starts when its last segment has landed and the previous pass is done, takes len * t_segment, and
  every pass after the first pays t_pass (it accumulates into C).  Returns the split with the earliest finish; ties go
  to fewer passes."""
  import itertools
  best, best_t = None, None
  for k in range(1, min(max_groups, n_segments) + 1):
    for cuts in itertools.combinations(range(1, n_segments), k - 1):
      bounds = (0,) + cuts + (n_segments,)
      t = 0.0
      for g in range(k):
        lo, hi = bounds[g], bounds[g + 1]
        t = max(t, (hi - 1) * t_push) + (hi - lo) * t_segment + (t_pass if g > 0 else 0.0)
      if best_t is None or t < best_t - 1e-12:
        best, best_t = bounds, t
  return [list(range(best[g], best[g + 1])) for g in range(len(best) - 1)]


class _Trace(object):
  """Timeline of one streamed evaluation (FLAGS.dot_trace): named CUDA events on the streams involved, reported as
  milliseconds since the evaluation began.  A diagnostic; costs a few event records."""

  def __init__(self, ctx, main):
    self.t0 = torch.cuda.Event(enable_timing=True)
    self.t0.record(main)
    self.marks = []

  def mark(self, name, stream):
    ev = torch.cuda.Event(enable_timing=True)
    ev.record(stream)
    self.marks.append((name, ev))

  def report(self):
    torch.cuda.synchronize()
    return [(name, round(self.t0.elapsed_time(ev), 3)) for name, ev in self.marks]


class DotExpr(Expr):
  """dot.py:95-158 (the node the reference defines but no longer constructs -- Q11 -- is the natural
  home of the GEMM evaluator)."""
  members = ('matrix_a', 'matrix_b', 'tile_hint')

  def __str__(self):
    return 'Dot[%s, %s, %s]' % (self.matrix_a, self.matrix_b, self.tile_hint)

  def compute_shape(self):
    a, b = self.matrix_a.shape, self.matrix_b.shape
    if len(a) == 1 and len(b) == 1:
      return (1,)
    if len(a) > 1 and len(b) == 1:
      return (a[0],)
    if len(a) > 1 and len(b) > 1:
      return (a[0], b[1])
    raise ValueError('vector x matrix dot is not defined by the reference (tests/test_dot.py:45-53)')

  # ------------------------------------------------------------------ host operands: upload || contract || read back
  def evaluate(self):
    cache = self.cache()
    if cache is not None:
      return cache
    ctx = blob_ctx.get()
    if self._streamable(ctx):
      value = self._evaluate_streamed(ctx)
      if self.needs_cache:
        eval_cache.set(self.expr_id, value)
      return value
    return Expr.evaluate(self)

  def _streamable(self, ctx):
    """Both operands are host arrays that have not been uploaded yet (``from_numpy`` nodes, write_array.py:424-445),
    large enough for PCIe time to matter, on one GPU: the upload is then pipelined with the contraction instead of
    running in front of it."""
    if not FLAGS.dot_stream_host_operands or ctx.device.type != 'cuda':
      return False
    if FLAGS.dot_precision == 'simt':
      return False
    a, b = self.matrix_a, self.matrix_b
    for e in (a, b):
      if not isinstance(e, WriteArrayExpr) or e.cache() is not None:
        return False
      npa = e.npa
      if npa.ndim != 2 or npa.dtype != np.float32 or npa.shape[0] == 0 or npa.strides[1] != npa.itemsize:
        return False
    if a.npa.shape[1] != b.npa.shape[0] or a.npa.shape[1] == 0 or b.npa.shape[1] == 0:
      return False
    return a.npa.nbytes + b.npa.nbytes >= FLAGS.dot_stream_min_bytes

  def _evaluate_streamed(self, ctx):
    """C = A . B with A and B still in host memory.  A is cut into row strips and B into column strips; strip s of
    each is uploaded on a copy stream while the compute stream contracts what has already arrived: after A_s lands,
    C[rows_s, columns of strips < s]; after B_s lands, C[rows of strips <= s, columns_s] (an L-shaped frontier, the
    order that makes the most output computable per byte uploaded).  Every C element is still
    produced by one contraction over the full K in the order of the resident path: bit-identical to it.  Each launch
    leaves an event behind; ``DistArrayImpl.read_local_into`` uses them to start the D2H copy of a finished block
    while later blocks are still being computed.  The operand arrays end up resident and cached exactly as
    ``from_numpy(...).evaluate()`` would leave them."""
    ea, eb = self.matrix_a, self.matrix_b
    a_np, b_np = ea.npa, eb.npa
    M, K = a_np.shape
    N = b_np.shape[1]
    precision = FLAGS.dot_precision
    av = distarray.create(a_np.shape, np.float32, tile_hint=ea.tile_hint)
    bv = distarray.create(b_np.shape, np.float32, tile_hint=eb.tile_hint)
    if ctx.num_workers > 1:
      return self._evaluate_streamed_ranks(ctx, av, bv, M, N, K, precision)
    target = distarray.create((M, N), np.float32, reducer=np.add, tile_hint=self.tile_hint or (M, N))
    if av.slab is None or bv.slab is None or target.slab is None:
      raise SpartanError('streamed dot expects slab-backed arrays on a single rank')
    strip = int(FLAGS.dot_stream_strip)
    pa = device_ops.PreparedOperand(M, K, precision, 'dot_stream_a')
    pb = device_ops.PreparedOperand(N, K, precision, 'dot_stream_b')

    def strips(n):
      # equal strips, the last one halved: what can only start after the final byte has arrived (and must be read back
      # after the final launch) is half as large
      out = [(r, min(n, r + strip)) for r in range(0, n, strip)]
      lo, hi = out[-1]
      if len(out) > 1 and hi - lo >= 1024:
        mid = lo + ((hi - lo) // 2 + 255) // 256 * 256
        out[-1:] = [(lo, mid), (mid, hi)]
      return out
    ra, cb = strips(M), strips(N)
    main = torch.cuda.current_stream(ctx.device)
    copy = ctx.side_stream('h2d')
    copy.wait_stream(main)              # recycled allocations may still be in use by work queued on the main stream
    C = target.slab
    done = []

    def contract(r0, r1, c0, c1):
      device_ops.gemm_prepared_rows(pa, r0, r1, pb, c0, c1, C[r0:r1, c0:c1], accumulate=False)
      done.append((extent.create((r0, c0), (r1, c1), (M, N)), main.record_event()))

    for s in range(max(len(ra), len(cb))):                      # L-shaped frontier
      ev_a = ev_b = None
      with torch.cuda.stream(copy):
        if s < len(ra):
          r0, r1 = ra[s]
          device_ops.upload_rect(av.slab[r0:r1, :], a_np[r0:r1, :])
          ev_a = copy.record_event()
        if s < len(cb):
          c0, c1 = cb[s]
          device_ops.upload_rect(bv.slab[:, c0:c1], b_np[:, c0:c1])
          ev_b = copy.record_event()
      if ev_a is not None:
        main.wait_event(ev_a)
        pa.prepare_a(av.slab[r0:r1, :], r0)
        ncols = cb[min(s, len(cb)) - 1][1] if s > 0 else 0      # columns whose strips (< s) are prepared already
        if ncols:
          contract(r0, r1, 0, ncols)
      if ev_b is not None:
        main.wait_event(ev_b)
        pb.prepare_b(bv.slab[:, c0:c1], c0)
        rows = ra[min(s, len(ra) - 1)][1]                       # rows whose strips (<= s) are prepared
        if s == max(len(ra), len(cb)) - 1 and rows >= 2048:
          half = (rows // 2 + 255) // 256 * 256                 # final launch in two: read-back of the first half overlaps
          contract(0, half, c0, c1)
          contract(half, rows, c0, c1)
        else:
          contract(0, rows, c0, c1)
    for arr in (av, bv, target):
      for tid in arr.tiles.values():
        ctx.tile(tid).valid = True
    target.block_events = done
    for e, v in ((ea, av), (eb, bv)):
      if e.needs_cache:
        eval_cache.set(e.expr_id, v)
    return target

  def _evaluate_streamed_ranks(self, ctx, av, bv, M, N, K, precision):
    """The streamed dot on several GPUs (regular placement, see _regular_layout): every rank uploads only its own
    column blocks.  Its B block goes first; A then arrives in row strips, and for each strip
        copy stream:  H2D of the strip's owned columns
        comm stream:  split/round the strip (sp_gemm_prepare_a_rows) and ncclAllGather the prepared strip
        main stream:  C[strip, mine] = sum over source ranks p of  A_p[strip, :] . B[k_p, mine]   (one launch, W segments)
    run concurrently for different strips; a finished strip of C can be read back (read_local_into) while later
    strips are in flight.  Any other placement uploads the operands whole and takes the resident path."""
    import torch.distributed as dist
    ea, eb = self.matrix_a, self.matrix_b
    a_np, b_np = ea.npa, eb.npa
    W, me = ctx.num_workers, ctx.worker_id
    target = distarray.create((M, N), np.float32, reducer=np.add, tile_hint=self.tile_hint or (M, N))
    layout = _regular_layout(av, bv, target, W, M, K) \
      if (av.slab is not None and bv.slab is not None and target.slab is not None) else None
    if layout is not None and any(_runs(layout[1][w][1]) != layout[2][w][1] for w in range(W)):
      layout = None                     # a rank's B columns must be exactly its C columns (the slabs are used whole)
    if layout is None:
      for e, v, npa in ((ea, av, a_np), (eb, bv, b_np)):
        v.update(extent.from_shape(npa.shape), npa)
        if e.needs_cache:
          eval_cache.set(e.expr_id, v)
      return Expr.evaluate(self)
    a_axes, b_axes, c_runs = layout
    ka, nb = av.slab.shape[1], bv.slab.shape[1]
    strip = int(FLAGS.dot_stream_strip)
    n_strips = -(-M // strip)
    if W <= SP_GEMM_MAX_SEGMENTS and 2 * n_strips * W <= 2048 and ctx.peer.available():
      return self._evaluate_streamed_peer(ctx, av, bv, target, a_axes, b_axes, c_runs, M, N, K, precision)
    main = torch.cuda.current_stream(ctx.device)
    copy, commst = ctx.side_stream('h2d'), ctx.side_stream('dot_comm')
    copy.wait_stream(main); commst.wait_stream(main)

    def upload_cols(slab, host, r0, r1, col_intervals):
      off = 0
      for c0, c1 in col_intervals:
        device_ops.upload_rect(slab[r0:r1, off:off + (c1 - c0)], host[r0:r1, c0:c1])
        off += c1 - c0

    # B[:, mine]: up first, prepared once per source rank p as the rows k_p of the transposed operand
    with torch.cuda.stream(copy):
      upload_cols(bv.slab, b_np, 0, K, b_axes[me][1])
      ev_b = copy.record_event()
    main.wait_event(ev_b)
    pbs = []
    for p in range(W):
      pb = device_ops.PreparedOperand(nb, ka, precision, 'dot_ms_b%d' % p)
      off = 0
      for a, b in a_axes[p][1]:
        pb.prepare_b(bv.slab[a:b, :], 0, k_offset=off)
        off += b - a
      pbs.append(pb)
    ev_pb = main.record_event()
    strip = int(FLAGS.dot_stream_strip)
    done = []
    for i, r0 in enumerate(range(0, M, strip)):
      r1 = min(M, r0 + strip)
      with torch.cuda.stream(copy):
        upload_cols(av.slab, a_np, r0, r1, a_axes[me][1])
        ev_a = copy.record_event()
      with torch.cuda.stream(commst):
        commst.wait_event(ev_a)
        if i == 0:
          commst.wait_event(ev_pb)        # scratch buffers recycled from an earlier evaluation: main-stream readers first
        # constructed on the stream that fills it: the zero-fill of a padded operand (K/W not a multiple of the k-block)
        # must be ordered before prepare_a and the gather, not behind the previous strip's GEMM on the main stream
        mine = device_ops.PreparedOperand(r1 - r0, ka, precision, 'dot_ms_mine%d' % i)
        nbytes = mine.buf.numel()
        gathered = ctx.scratch(W * nbytes, 'dot_ms_gath%d' % i)[:W * nbytes]
        mine.prepare_a(av.slab[r0:r1, :], 0)
        dist.all_gather_into_tensor(gathered, mine.buf)
        ev_g = commst.record_event()
      main.wait_event(ev_g)
      views = [(gathered.data_ptr() + p * nbytes, mine.copy_stride, pbs[p].row_ptr(0), pbs[p].copy_stride, mine.Kp)
               for p in range(W)]
      device_ops.gemm_prepared_views(views, target.slab[r0:r1, :], False, precision)
      ev_c = main.record_event()
      for c0, c1 in c_runs[me][1]:
        done.append((extent.create((r0, c0), (r1, c1), (M, N)), ev_c))
    for arr in (av, bv, target):
      for tid in arr.tiles.values():
        if ctx.is_local(tid):
          ctx.tile(tid).valid = True
    target.block_events = done
    for e, v in ((ea, av), (eb, bv)):
      if e.needs_cache:
        eval_cache.set(e.expr_id, v)
    return target

  def _evaluate_streamed_peer(self, ctx, av, bv, target, a_axes, b_axes, c_runs, M, N, K, precision):
    """The streamed multi-GPU dot over peer memory.  Per A row strip i, every rank
        copy stream:  H2D of the strip's owned columns (all strips are queued up front; PCIe never waits for compute)
        main stream:  split/round the strip straight into slot [i][me] of its symmetric gather buffer
        push stream:  copy-engine push of that slot into slot [i][me] of every peer + the epoch word behind it
        main stream:  ONE gated launch  C[strip, mine] = sum_p A_p[strip, :] . B[k_p, mine]  that starts on the local
                      segment and picks each peer's segment up when its flag arrives
    with the strip loop software-pipelined on the main stream (prepare i+1 is queued before contract i), so the only SM
    work between two contractions is one short preparation kernel and nothing on the SMs ever waits for a collective."""
    ea, eb = self.matrix_a, self.matrix_b
    a_np, b_np = ea.npa, eb.npa
    W, me = ctx.num_workers, ctx.worker_id
    peer = ctx.peer
    ka, nb = av.slab.shape[1], bv.slab.shape[1]
    strip = int(FLAGS.dot_stream_strip)
    strips = [(r0, min(M, r0 + strip)) for r0 in range(0, M, strip)]
    S = len(strips)
    Kp = device_ops.gemm_kpad(ka, precision)
    part = device_ops.gemm_prepared_bytes(strip, Kp, precision)         # bytes of one (strip, source rank) slot
    rb = device_ops.gemm_row_bytes(Kp, precision)
    main = torch.cuda.current_stream(ctx.device)
    copy, push = ctx.side_stream('h2d'), ctx.side_stream('peer_push')
    copy.wait_stream(main)
    if ctx.push_done is not None:
      main.wait_event(ctx.push_done)
    gather = peer.buffer('dot_stream_gather', 2 * (S * W * part + 1024))
    f0 = peer.flag_range('dot_stream_gather', 2048)
    epoch, esrc = peer.next_epoch('dot_stream_gather')
    half = epoch & 1

    # The two halves of the buffer (and of the flag range) are FIXED regions, whatever the strip count of this evaluation:
    # exchange e only ever touches half e & 1, so an evaluation with other shapes cannot land in memory a slower peer is
    # still reading for the previous exchange.
    half_bytes = gather.nbytes // 2 // 1024 * 1024

    def slot_off(i, p):                 # byte offset of slot (strip i, source p) in this exchange's half
      return half * half_bytes + (i * W + p) * part

    def flag_idx(i, p):
      return f0 + half * 1024 + i * W + p

    def upload_cols(slab, host, r0, r1, col_intervals):
      off = 0
      for c0, c1 in col_intervals:
        device_ops.upload_rect(slab[r0:r1, off:off + (c1 - c0)], host[r0:r1, c0:c1])
        off += c1 - c0

    trace = _Trace(ctx, main) if FLAGS.dot_trace else None
    with torch.cuda.stream(copy):
      if trace: trace.mark('h2d begin', copy)
      upload_cols(bv.slab, b_np, 0, K, b_axes[me][1])
      ev_b = copy.record_event()
      if trace: trace.mark('h2d B done', copy)
      ev_a = []
      for i, (r0, r1) in enumerate(strips):
        upload_cols(av.slab, a_np, r0, r1, a_axes[me][1])
        ev_a.append(copy.record_event())
        if trace: trace.mark('h2d A%d done' % i, copy)

    order = [(me + j) % W for j in range(W)]

    def prepare_and_push(i):
      r0, r1 = strips[i]
      main.wait_event(ev_a[i])
      mine = device_ops.PreparedOperand(r1 - r0, ka, precision, None,
                                        buf=gather.tensor[slot_off(i, me):slot_off(i, me) + part])
      mine.prepare_a(av.slab[r0:r1, :], 0)
      ev = main.record_event()
      if trace: trace.mark('prep A%d done' % i, main)
      with torch.cuda.stream(push):
        push.wait_event(ev)
        dsts = [(me - j) % W for j in range(1, W)]           # ring order: the peer that needs this slot first goes first
        peer.push([gather.ptrs[d] + slot_off(i, me) for d in dsts], mine.buf.data_ptr(), mine.nbytes,
                  [peer.flag_ptr(d, flag_idx(i, me)) for d in dsts], esrc)
        ctx.push_done = push.record_event()
        if trace: trace.mark('push A%d done' % i, push)
      return mine

    # B[:, mine]: prepared once per source rank p as the rows k_p of the transposed operand
    main.wait_event(ev_b)
    pbs = []
    for p in range(W):
      pb = device_ops.PreparedOperand(nb, ka, precision, 'dot_ms_b%d' % p)
      off = 0
      for a, b in a_axes[p][1]:
        pb.prepare_b(bv.slab[a:b, :], 0, k_offset=off)
        off += b - a
      pbs.append(pb)
    done = []
    mines = {0: prepare_and_push(0)}
    for i, (r0, r1) in enumerate(strips):
      if i + 1 < S:
        mines[i + 1] = prepare_and_push(i + 1)
      mine = mines.pop(i)
      views, flags = [], []
      for p in order:
        a_ptr = gather.local_ptr + slot_off(i, p)
        views.append((a_ptr, (r1 - r0) * rb, pbs[p].row_ptr(0), pbs[p].copy_stride, Kp))
        flags.append(0 if p == me else peer.flag_ptr(me, flag_idx(i, p)))
      device_ops.gemm_prepared_views_gated(views, flags, [epoch] * W, peer.status.data_ptr(), target.slab[r0:r1, :],
                                           False, precision)
      ev_c = main.record_event()
      if trace: trace.mark('gemm %d done' % i, main)
      for c0, c1 in c_runs[me][1]:
        done.append((extent.create((r0, c0), (r1, c1), (M, N)), ev_c))
    for arr in (av, bv, target):
      for tid in arr.tiles.values():
        if ctx.is_local(tid):
          ctx.tile(tid).valid = True
    target.block_events = done
    target.trace = trace
    for e, v in ((ea, av), (eb, bv)):
      if e.needs_cache:
        eval_cache.set(e.expr_id, v)
    return target

  def _peer_gather_path(self, ctx, av, bv, target, shape, M, N, K, dtype, precision):
    """Multi-GPU dot over peer memory for the regular placement (every rank owns whole column blocks of A, B and C):
    every rank prepares its own A slab once (cached while A is unchanged), its copy engines push the prepared slab into
    the symmetric gather buffer of every peer -- ring order, an epoch word behind each copy -- and ONE gated tcgen05
    launch per C block contracts  C[:, mine] = sum_p A_p . B[k_p, mine]  starting with the local segment and taking
    each peer's segment as its flag arrives.  The exchange costs no SM and is hidden behind the segments already present;
    the gather buffer is double-buffered by epoch so a fast rank never overwrites a slab a slow peer is still reading.
    Returns False when the placement is not of this form or CUDA IPC is unavailable (NCCL path follows)."""
    W, me = ctx.num_workers, ctx.worker_id
    if W == 1 or W > SP_GEMM_MAX_SEGMENTS or len(shape) != 2 or dtype != np.float32 or precision == 'simt':
      return False
    if not (isinstance(av, distarray.DistArrayImpl) and isinstance(bv, distarray.DistArrayImpl)):
      return False
    if av.dtype != np.float32 or bv.dtype != np.float32:
      return False
    if av.slab is None or bv.slab is None or target.slab is None:
      return False
    layout = _regular_layout(av, bv, target, W, M, K)
    if layout is None or not ctx.peer.available():
      return False
    a_axes, b_axes, c_runs = layout
    peer = ctx.peer
    width = av.slab.shape[1]
    Kp = device_ops.gemm_kpad(width, precision)
    a_bytes = device_ops.gemm_prepared_bytes(M, Kp, precision)
    rb = device_ops.gemm_row_bytes(Kp, precision)
    main = torch.cuda.current_stream(ctx.device)
    push = ctx.side_stream('peer_push')
    if ctx.push_done is not None:
      main.wait_event(ctx.push_done)        # an earlier push may still be reading the buffer about to be re-prepared
    mine = device_ops.cached_operand(av, ('a_slab',), M, width, precision, lambda op: op.prepare_a(av.slab, 0))
    gather = peer.buffer('dot_gather', 2 * (W * a_bytes + 1024))
    half_bytes = gather.nbytes // 2 // 1024 * 1024      # the halves are fixed regions (see _evaluate_streamed_peer)
    f0 = peer.flag_range('dot_gather', 2 * W)
    epoch, esrc = peer.next_epoch('dot_gather')
    half = epoch & 1
    ev = main.record_event()
    with torch.cuda.stream(push):
      push.wait_event(ev)
      dsts = [(me - j) % W for j in range(1, W)]
      peer.push([gather.ptrs[d] + half * half_bytes + me * a_bytes for d in dsts], mine.buf.data_ptr(), mine.nbytes,
                [peer.flag_ptr(d, f0 + half * W + me) for d in dsts], esrc)
      ctx.push_done = push.record_event()
    order = [(me + j) % W for j in range(W)]
    # The kernel walks K innermost: a launch cannot finish its first tile before the last of ITS segments has landed.
    # So the segments are contracted in a few passes, grouped by when they arrive (local first; then what the copy
    # engines deliver while the previous pass computes), each pass accumulating into C.  The gates keep any grouping
    # correct; the grouping only decides how much of the exchange hides behind the math (measured on 2 x B200 with the
    # 8-rank traffic pattern: one pass over all segments 20-22 ms, the GEMM alone 15.8-17 ms, profiles/r02_peer_probe*).
    n_cols = sum(c1 - c0 for c0, c1 in c_runs[me][1])
    if FLAGS.dot_passes == 'single':
      groups = [list(range(W))]
    else:
      groups = _arrival_groups(W, mine.nbytes / _PUSH_BYTES_PER_S, 2.0 * M * n_cols * width / _DOT_FLOPS_PER_S,
                               2.0 * M * n_cols * 4 / _PASS_BYTES_PER_S + _PASS_FIXED_S)
    for (r0, r1) in c_runs[me][0]:
      for (c0, c1) in c_runs[me][1]:
        Cv = target.fetch(extent.create((r0, c0), (r1, c1), shape))
        for gi, grp in enumerate(groups):
          views, flags = [], []
          for p in [order[j] for j in grp]:
            def fill(op, p=p):
              off = 0
              for a, b in a_axes[p][1]:
                Bv = bv.fetch(extent.create((a, c0), (b, c1), bv.shape))       # zero-copy view of this rank's B slab
                op.prepare_b(Bv, 0, k_offset=off)
                off += b - a
            pb = device_ops.cached_operand(bv, ('b_cols', c0, c1, p), c1 - c0, width, precision, fill)
            a_ptr = mine.row_ptr(r0) if p == me else gather.local_ptr + half * half_bytes + p * a_bytes + r0 * rb
            views.append((a_ptr, M * rb, pb.row_ptr(0), pb.copy_stride, Kp))
            flags.append(0 if p == me else peer.flag_ptr(me, f0 + half * W + p))
          device_ops.gemm_prepared_views_gated(views, flags, [epoch] * len(views), peer.status.data_ptr(), Cv, gi > 0,
                                               precision)
    return True

  def _allgather_path(self, ctx, av, bv, target, shape, M, N, K, dtype, precision):
    """Multi-GPU fast path for the regular placement (every rank owns whole column blocks of A, B and C, e.g.
    round-robin over an evenly divisible tile grid): ONE NCCL all-gather moves the A slabs, then every rank
    contracts C[:, mine] = sum over source ranks p, column intervals (a,b) of p:
        A_p[:, a:b] (view into the gathered buffer, no copy)  .  B[a:b, mine] (view into this rank's B slab)
    as K-segments of a single tcgen05 launch per block.  The local segment runs while the gather is in
    flight.  Returns False when the placement is not of this form (the general rectangle-fetch path follows)."""
    import torch.distributed as dist
    W, me = ctx.num_workers, ctx.worker_id
    if W == 1 or len(shape) != 2 or dtype != np.float32:
      return False
    if not (isinstance(av, distarray.DistArrayImpl) and isinstance(bv, distarray.DistArrayImpl)):
      return False
    if av.dtype != np.float32 or bv.dtype != np.float32:
      return False
    layout = _regular_layout(av, bv, target, W, M, K)
    if layout is None:
      return False
    a_axes, b_axes, c_runs = layout
    if av.slab is None or bv.slab is None or target.slab is None:
      return False

    # Each rank prepares (rounds / splits) ONLY its own A slab; the prepared slabs are what travels.
    if precision == 'simt':
      return False
    width = av.slab.shape[1]
    Kp = device_ops.gemm_kpad(width, precision)
    a_bytes = device_ops.gemm_prepared_bytes(M, Kp, precision)
    mine = ctx.scratch(a_bytes, 'dot_a_prepared')[:a_bytes]
    if Kp != (width + 3) // 4 * 4:
      mine.zero_()
    device_ops.gemm_prepare_a(av.slab, mine, Kp, 0, precision)
    gathered = ctx.scratch(W * a_bytes, 'dot_allgather')[:W * a_bytes]
    work = dist.all_gather_into_tensor(gathered, mine, async_op=True)

    blocks = [(r, c) for r in c_runs[me][0] for c in c_runs[me][1]]
    # B strips for every source rank p: rows = p's k-intervals (in slab order), columns = this block
    b_prepared = {}
    for (r0, r1), (c0, c1) in blocks:
      b_bytes = device_ops.gemm_prepared_bytes(c1 - c0, Kp, precision)
      buf = ctx.scratch(W * b_bytes, 'dot_b_prepared_%d_%d' % (c0, c1))[:W * b_bytes]
      if Kp != width:
        buf.zero_()
      for p in range(W):
        out = buf[p * b_bytes:(p + 1) * b_bytes]
        off = 0
        for a, b in a_axes[p][1]:
          Bv = bv.fetch(extent.create((a, c0), (b, c1), bv.shape))     # zero-copy view of this rank's B slab
          device_ops.gemm_prepare_b(Bv, out, Kp, off, precision)
          off += b - a
        b_prepared[(c0, c1, p)] = out
    views = {}
    for (r0, r1), (c0, c1) in blocks:       # local contribution first: overlaps the all-gather
      Cv = target.fetch(extent.create((r0, c0), (r1, c1), shape))
      views[(c0, c1)] = Cv
      device_ops.gemm_prepared([(mine, b_prepared[(c0, c1, me)], Kp)], Cv, accumulate=False, precision=precision)
    work.wait()
    for (r0, r1), (c0, c1) in blocks:
      segs = [(gathered[p * a_bytes:(p + 1) * a_bytes], b_prepared[(c0, c1, p)], Kp) for p in range(W) if p != me]
      device_ops.gemm_prepared(segs, views[(c0, c1)], accumulate=True, precision=precision)
    return True

  def _evaluate(self, ctx, deps):
    av = deps['matrix_a']
    bv = deps['matrix_b']
    if isinstance(bv, np.ndarray):
      bv = distarray.LocalWrapper(bv)      # dot.py:254-262: a NumPy right operand is held by every rank
    a_nd, b_nd = len(av.shape), len(bv.shape)
    if a_nd == 1 and b_nd == 1:
      if av.shape[0] != bv.shape[0]:
        raise ValueError('objects are not aligned')
      shape, M, N, K = (1,), 1, 1, av.shape[0]
    elif a_nd > 1 and b_nd == 1:
      if av.shape[1] != bv.shape[0]:
        raise ValueError('objects are not aligned')
      shape, M, N, K = (av.shape[0],), av.shape[0], 1, av.shape[1]
    elif a_nd > 1 and b_nd > 1:
      if av.shape[1] != bv.shape[0]:
        raise ValueError('objects are not aligned')
      shape, M, N, K = (av.shape[0], bv.shape[1]), av.shape[0], bv.shape[1], av.shape[1]
    else:
      raise ValueError
    tile_hint = self.tile_hint
    if tile_hint is None and len(shape) == 2:
      tile_hint = shape                     # dot.py:281-282
    dtype = np.result_type(av.dtype, bv.dtype)
    if dtype.kind == 'b':
      dtype = np.dtype(np.int64)
    target = distarray.create(shape, dtype, reducer=np.add, tile_hint=tile_hint)

    precision = FLAGS.dot_precision
    W = ctx.num_workers
    me = ctx.worker_id
    # 2-D views of everything: vectors become [K,1] / [1,K] / [M,1]
    def a_region(r0, r1):
      if a_nd == 1:
        return extent.create((0,), (K,), av.shape)
      return extent.create((r0, 0), (r1, K), av.shape)

    def b_region(c0, c1):
      if b_nd == 1:
        return extent.create((0,), (K,), bv.shape)
      return extent.create((0, c0), (K, c1), bv.shape)

    def as2d(t, rows, cols):
      return t.reshape(rows, cols)

    def cast(t):
      return t if t.dtype == blob_ctx.torch_dtype(dtype) else t.to(blob_ctx.torch_dtype(dtype))

    if (self._peer_gather_path(ctx, av, bv, target, shape, M, N, K, dtype, precision) or
        self._allgather_path(ctx, av, bv, target, shape, M, N, K, dtype, precision)):
      for tid in target.tiles.values():
        if ctx.is_local(tid):
          ctx.tile(tid).valid = True
      return target

    for w in range(W):
      if len(shape) == 2:
        runs = _owned_runs(target, w)
        if runs is None:     # scattered placement: one block per owned tile
          blocks = [((ex.ul[0], ex.lr[0]), (ex.ul[1], ex.lr[1])) for ex, tid in target.tiles.items()
                    if tid.worker == w]
        else:
          blocks = [(r, c) for r in runs[0] for c in runs[1]]
      else:
        blocks = [((ex.ul[0], ex.lr[0]), (0, 1)) for ex, tid in target.tiles.items() if tid.worker == w]
        if a_nd == 1:
          blocks = [((0, 1), (0, 1))] if any(tid.worker == w for tid in target.tiles.values()) else []
      a_cache, b_cache = {}, {}
      for (r0, r1), (c0, c1) in blocks:
        if (r0, r1) not in a_cache:
          a_cache[(r0, r1)] = av.fetch(a_region(r0, r1), dst=w)
        if (c0, c1) not in b_cache:
          b_cache[(c0, c1)] = bv.fetch(b_region(c0, c1), dst=w)
        if w != me:
          continue
        A = cast(as2d(a_cache[(r0, r1)], r1 - r0 if a_nd > 1 else 1, K))
        B = cast(as2d(b_cache[(c0, c1)], K, c1 - c0 if b_nd > 1 else 1))
        if len(shape) == 2:
          creg = extent.create((r0, c0), (r1, c1), shape)
        else:
          creg = extent.create((r0,), (r1,), shape) if a_nd > 1 else extent.create((0,), (1,), shape)
        Cv = target.fetch(creg)            # zero-copy view of this rank's slab / tile
        C2 = Cv.reshape(A.shape[0], B.shape[1])
        if C2.data_ptr() != Cv.data_ptr() or C2.stride(-1) != 1:
          raise SpartanError('dot target block is not addressable as a row-major view')
        if (W == 1 and len(shape) == 2 and dtype == np.float32 and precision != 'simt' and K > 0
            and isinstance(av, distarray.DistArrayImpl) and isinstance(bv, distarray.DistArrayImpl)
            and A.stride(1) == 1 and B.stride(1) == 1 and FLAGS.dot_prepared_cache):
          # operands prepared once while the arrays are unchanged (same kernels, same bits as the uncached call)
          pa = device_ops.cached_operand(av, ('a_rows', r0, r1), r1 - r0, K, precision, lambda op, A=A: op.prepare_a(A, 0))
          pb = device_ops.cached_operand(bv, ('b_cols', c0, c1), c1 - c0, K, precision, lambda op, B=B: op.prepare_b(B, 0))
          device_ops.gemm_prepared_rows(pa, 0, r1 - r0, pb, 0, c1 - c0, C2, accumulate=False)
          continue
        device_ops.gemm_views(A, B, C2, accumulate=False, precision=precision)   # transposed views: no copy
    for tid in target.tiles.values():
      if ctx.is_local(tid):
        ctx.tile(tid).valid = True
    return target


def dot(a, b, tile_hint=None):
  """Compute the dot product (matrix multiplication) of 2 arrays (dot.py:243-299).

  :param a: `Expr`
  :param b: `Expr` or `numpy.ndarray`
  :param tile_hint: tiling of the result (default: one tile, like the reference)
  :rtype: `DotExpr`
  """
  from ..sparse import SparseStripsExpr, spmv
  if isinstance(a, SparseStripsExpr):          # dot.py:212-217 sparse branch: sparse strips x dense vector
    return spmv(a, b, tile_hint)
  a = lazify(a)
  if not isinstance(b, np.ndarray):
    b = lazify(b)
  e = DotExpr(matrix_a=a, matrix_b=b, tile_hint=tile_hint)
  e.compute_shape()       # raises ValueError early for undefined shapes, like dot.py:264-299
  a_shape, b_shape = a.shape, b.shape
  if (len(a_shape) == 1 and a_shape[0] != b_shape[0]) or (len(a_shape) > 1 and a_shape[1] != b_shape[0]):
    raise ValueError('objects are not aligned')
  return e
