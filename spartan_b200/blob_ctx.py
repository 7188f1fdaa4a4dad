"""BlobCtx: the tile store and kernel dispatcher of one GPU rank.

Keeps the reference's name and call surface (spartan/blob_ctx.py:18-301: create / get / update /
destroy_all / map / tile_op / num_workers / is_master, module-level get()/set()) but none of its
mechanism: there is no master, no RPC and no pickled closures.  The job is SPMD -- one process per
GPU, every process builds the same expression DAG and walks the same tiles in the same order:

  * ``num_workers``  = number of GPU ranks (torch.distributed world size), ``worker_id`` = this rank;
  * a tile lives in the HBM of exactly one rank; TileId = (worker, per-worker sequence number) is
    assigned deterministically, so every rank knows the whole extent -> TileId table of every array
    without exchanging a byte (the reference learns it from RPC replies, distarray.py:202-208);
  * ``map`` runs the per-tile kernel function for every tile on every rank; the function launches CUDA
    work only where the tile is local (``ctx.is_local``) and takes part in collectives otherwise.

Device memory comes from torch (caching allocator, streams); compute goes through the C ABI.
"""
import os
import threading

import numpy as np
import torch

from . import comm
from ._lib import SpartanError
from .core import TileId

_TORCH_DTYPES = {
  np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
  np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64,
  np.dtype(np.uint8): torch.uint8, np.dtype(np.bool_): torch.bool,
}


def torch_dtype(dtype):
  dtype = np.dtype(dtype)
  if dtype not in _TORCH_DTYPES:
    raise SpartanError('dtype %s is not supported by the device evaluator (float32/float64/int32/int64/uint8/bool)'
                       % dtype)
  return _TORCH_DTYPES[dtype]


class BlobCtx(object):
  def __init__(self, worker_id=0, num_workers=1, device=None):
    self.worker_id = int(worker_id)
    self.num_workers = int(num_workers)
    if device is None:
      device = torch.device('cuda', int(os.environ.get('LOCAL_RANK', 0))) if torch.cuda.is_available() \
        else torch.device('cpu')
    self.device = torch.device(device)
    self._blobs = {}                          # TileId -> DeviceTile, local tiles only (worker.py:70 _blobs)
    self._next_id = [0] * self.num_workers    # per-worker id sequence, identical on every rank
    self._rr = 0
    self._scratch = {}
    self._retired_scratch = []                # outgrown scratch buffers a captured graph may still address
    self.graph_captures = 0                   # evaluations captured as CUDA graphs so far (replay.py)
    self._side_streams = {}
    self.kernel_launches = 0                  # launches of this library's kernels (bench.py "gpu_launches")
    self.data_epoch = 0                       # bumped whenever an existing array may have changed in place
    self._peer = None
    self.push_done = None                     # event after the last copy-engine push of this rank
    self.numa_cpus = None                     # CPUs this rank was pinned to (multi-rank jobs)

  @property
  def peer(self):
    """Peer memory of the job (spartan_b200/peer.py), created on first use."""
    if self._peer is None:
      from .peer import PeerMemory
      self._peer = PeerMemory(self)
    return self._peer

  def touch(self):
    """Some array was modified in place: derived copies (prepared GEMM operands) are stale."""
    self.data_epoch += 1

  # ------------------------------------------------------------------ reference surface
  def is_master(self):
    return True      # SPMD: every rank holds the DAG; blob_ctx.py:35-40

  def is_local(self, tile_id):
    return tile_id.worker == self.worker_id or tile_id.worker < 0

  def new_tile_id(self, hint=-1):
    """blob_ctx.py:221-254: hint >= 0 picks worker hint % num_workers, otherwise round-robin."""
    if hint is None or hint < 0:
      worker = self._rr % self.num_workers
      self._rr += 1
    else:
      worker = int(hint) % self.num_workers
    tid = TileId(worker, self._next_id[worker])
    self._next_id[worker] += 1
    return tid

  def create(self, tile, hint=-1):
    """Registers ``tile`` (a DeviceTile, or a zero-argument factory producing one, evaluated only on
    the owning rank) and returns its TileId.  Synchronous: there is no Future to wait on."""
    tid = self.new_tile_id(hint)
    if tid.worker == self.worker_id:
      self._blobs[tid] = tile() if callable(tile) else tile
    return tid

  def get(self, tile_id, subslice=None):
    """Local tile data (a device tensor view); blob_ctx.py:127-143."""
    return self._blobs[tile_id].get(subslice)

  def tile(self, tile_id):
    return self._blobs[tile_id]

  def update(self, tile_id, subslice, data, reducer, wait=True):
    """blob_ctx.py:163-179 -> Tile.merge; local tiles only."""
    self.data_epoch += 1
    return self._blobs[tile_id].update(subslice, data, reducer)

  def destroy_all(self, tile_ids):
    for tid in tile_ids:
      self._blobs.pop(tid, None)

  def tile_op(self, tile_id, fn):
    return fn(self._blobs[tile_id])

  def map(self, tile_ids, mapper_fn, kw):
    """blob_ctx.py:256-275 + Worker._run_kernel (worker.py:232-304): every rank calls ``mapper_fn`` for
    every tile, in a canonical order (worker, id); results keyed by tile id."""
    out = {}
    for tid in sorted(tile_ids, key=lambda t: (t.worker, t.id)):
      out[tid] = mapper_fn(tid, self._blobs.get(tid), **kw)
    return out

  # ------------------------------------------------------------------ device helpers
  def stream_ptr(self):
    if self.device.type != 'cuda':
      raise SpartanError('no CUDA device: spartan_b200 has no CPU fallback for compute kernels')
    return torch.cuda.current_stream(self.device).cuda_stream

  def side_stream(self, name):
    """A second CUDA stream of this rank (H2D / D2H copies that overlap kernels on the main stream)."""
    st = self._side_streams.get(name)
    if st is None:
      st = self._side_streams[name] = torch.cuda.Stream(self.device)
    return st

  def empty(self, shape, dtype):
    return torch.empty(tuple(int(s) for s in shape), dtype=torch_dtype(dtype), device=self.device)

  def scratch(self, nbytes, key='default'):
    """Grow-only device scratch buffers (reduction partials, GEMM operand copies)."""
    buf = self._scratch.get(key)
    if buf is None or buf.numel() < nbytes:
      if buf is not None and self.graph_captures > 0:
        # a captured CUDA graph (replay.py) may have baked this buffer's address in: keep it alive instead of handing
        # its memory back to the allocator
        self._retired_scratch.append(buf)
      self._scratch[key] = None
      buf = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
      self._scratch[key] = buf
    return buf

  def synchronize(self):
    if self.device.type == 'cuda':
      torch.cuda.synchronize(self.device)


_local = threading.local()
_global_ctx = [None]


def get():
  """blob_ctx.py:290-296."""
  ctx = getattr(_local, 'ctx', None) or _global_ctx[0]
  if ctx is None:
    ctx = initialize()
  return ctx


def set(ctx):
  """blob_ctx.py:298-301."""
  _local.ctx = ctx
  _global_ctx[0] = ctx


def _bind_to_gpu_numa_node(index):
  """Pins this process to the CPUs next to its GPU (sysfs local_cpulist of the GPU's PCI function), so that the pinned
  host buffers it allocates from now on are first-touched on that NUMA node and H2D / D2H traffic of the ranks does not
  all cross one socket link.  Multi-rank jobs only; SPARTAN_NUMA_BIND=0 disables.  Returns the CPU set or None."""
  if os.environ.get('SPARTAN_NUMA_BIND', '1') == '0':
    return None
  try:
    import pynvml
    pynvml.nvmlInit()
    visible = os.environ.get('CUDA_VISIBLE_DEVICES')
    phys = index
    if visible:
      ids = [v for v in visible.split(',') if v.strip()]
      if index < len(ids) and ids[index].strip().isdigit():
        phys = int(ids[index])
    bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(phys)).busId
    bus = bus.decode() if isinstance(bus, bytes) else bus
    dom, rest = bus.split(':', 1)
    path = '/sys/bus/pci/devices/%s:%s/local_cpulist' % (dom[-4:].lower(), rest.lower())
    with open(path) as f:
      text = f.read().strip()
    cpus = set()
    for part in text.split(','):
      if '-' in part:
        a, b = part.split('-')
        cpus.update(range(int(a), int(b) + 1))
      elif part:
        cpus.add(int(part))
    cpus &= os.sched_getaffinity(0)
    if not cpus:
      return None
    os.sched_setaffinity(0, cpus)
    return cpus
  except Exception:
    return None


def initialize(device=None):
  """One-time setup of this rank: picks cuda:LOCAL_RANK, joins the torch.distributed job when
  WORLD_SIZE > 1 (NCCL on GPUs, gloo on CPU) and installs the context.  Replaces
  spartan.initialize() -> cluster.start_cluster (spartan/__init__.py:42-56, cluster.py:125-169)."""
  rank, world = comm.init_from_env(device)
  if device is None and torch.cuda.is_available():
    device = torch.device('cuda', int(os.environ.get('LOCAL_RANK', 0)))
    torch.cuda.set_device(device)
  ctx = BlobCtx(rank, world, device)
  if world > 1 and ctx.device.type == 'cuda':
    ctx.numa_cpus = _bind_to_gpu_numa_node(ctx.device.index or 0)
  set(ctx)
  return ctx


def shutdown():
  ctx = _global_ctx[0]
  if ctx is not None and ctx._peer is not None:
    ctx._peer.close()
    ctx._peer = None
  _global_ctx[0] = None
  _local.ctx = None
