"""Peer memory of a multi-GPU job: symmetric HBM buffers every rank can write into with its copy engines.

The reference moves tiles between workers through its RPC layer (Worker.get / Worker.update, spartan/worker.py:126-217;
join_mapper's strip fetches, spartan/expr/operator/map.py:243-286).  Here a rank that produced a strip a peer needs
*pushes* it: a copy-engine transfer over NVLink/NVSwitch into a buffer the peer allocated (CUDA IPC), followed in
stream order by a 4-byte epoch word into the peer's flag slot.  No SM takes part, so the persistent tensor-core
kernel that fills every SM keeps running while operands arrive; it polls the flag itself and starts a K segment
the moment the segment is there (sp_gemm_prepared_views_gated).

All calls that allocate are collective (every rank makes them in the same order with the same sizes -- the SPMD
contract of blob_ctx); they only involve the host when a buffer is created or grown.
"""
import ctypes
import os

import torch
import torch.distributed as dist

from . import comm
from ._lib import lib, check, SpartanError


class _DevicePtr(object):
  """CUDA array interface over raw device memory (lets torch view a buffer this library allocated)."""

  def __init__(self, ptr, nbytes):
    self.__cuda_array_interface__ = {'shape': (int(nbytes),), 'typestr': '|u1', 'data': (int(ptr), False),
                                     'version': 3, 'strides': None}


class SymmetricBuffer(object):
  """One allocation of ``nbytes`` on every rank; ``ptrs[r]`` is rank r's copy as addressable from this process."""

  def __init__(self, nbytes, local_ptr, ptrs, device):
    self.nbytes = int(nbytes)
    self.local_ptr = int(local_ptr)
    self.ptrs = [int(p) for p in ptrs]
    self._keep = _DevicePtr(local_ptr, nbytes)
    self.tensor = torch.as_tensor(self._keep, device=device)       # uint8 view of the local copy


class PeerMemory(object):
  FLAG_WORDS = 4096

  def __init__(self, ctx):
    self.ctx = ctx
    self.world, self.rank = ctx.num_workers, ctx.worker_id
    self._buffers = {}
    self._available = None
    self._epochs = {}              # per buffer key: exchange rounds so far
    self._writes = 0
    self._flag_cursor = 0
    self._flag_ranges = {}
    self.flags = None              # SymmetricBuffer of FLAG_WORDS uint32
    self.epoch_src = None          # local device words the copy engines read the epoch from
    self.status = None             # device uint32: gates that timed out

  # ------------------------------------------------------------------ availability
  def available(self):
    """Collective, decided once: CUDA IPC between the ranks works (same node, peer access).  SPARTAN_PEER=0 forces the
    NCCL paths."""
    if self._available is not None:
      return self._available
    ok = self.world > 1 and self.ctx.device.type == 'cuda' and os.environ.get('SPARTAN_PEER', '1') != '0'
    if ok:
      try:
        self.flags = self._allocate(self.FLAG_WORDS * 4)
      except Exception:
        ok = False
    if self.world > 1 and self.ctx.device.type == 'cuda':
      t = torch.tensor([1 if ok else 0], device=self.ctx.device, dtype=torch.int32)
      dist.all_reduce(t, op=dist.ReduceOp.MIN)
      ok = bool(t.item())
    if ok:
      self.epoch_src = torch.zeros(64, dtype=torch.int32, device=self.ctx.device)
      self.status = torch.zeros(1, dtype=torch.int32, device=self.ctx.device)
    self._available = ok
    return ok

  # ------------------------------------------------------------------ allocation
  def _allocate(self, nbytes):
    nbytes = (int(nbytes) + 1023) // 1024 * 1024
    p = ctypes.c_void_p()
    check(lib.sp_peer_alloc(nbytes, ctypes.byref(p)), 'sp_peer_alloc')
    hb = lib.sp_peer_handle_bytes()
    handle = ctypes.create_string_buffer(hb)
    check(lib.sp_peer_export(p, handle), 'sp_peer_export')
    handles = [None] * self.world
    dist.all_gather_object(handles, bytes(handle.raw))
    ptrs = []
    for r in range(self.world):
      if r == self.rank:
        ptrs.append(p.value)
        continue
      q = ctypes.c_void_p()
      check(lib.sp_peer_import(ctypes.create_string_buffer(handles[r], hb), ctypes.byref(q)), 'sp_peer_import')
      ptrs.append(q.value)
    return SymmetricBuffer(nbytes, p.value, ptrs, self.ctx.device)

  def _release(self, buf):
    for r, q in enumerate(buf.ptrs):
      if r != self.rank:
        lib.sp_peer_close(ctypes.c_void_p(q))
    buf.tensor = None
    lib.sp_peer_free(ctypes.c_void_p(buf.local_ptr))

  def buffer(self, key, nbytes):
    """The symmetric buffer ``key`` with at least ``nbytes`` bytes (collective; grows by re-allocation after a device
    synchronise + barrier so that no peer is still writing into the old one)."""
    buf = self._buffers.get(key)
    if buf is not None and buf.nbytes >= nbytes:
      return buf
    torch.cuda.synchronize(self.ctx.device)
    comm.barrier()
    if buf is not None:
      self._buffers.pop(key)
      self._release(buf)
    buf = self._allocate(nbytes)
    self._buffers[key] = buf
    comm.barrier()
    return buf

  def flag_range(self, key, n):
    """First index of ``n`` consecutive flag words reserved for ``key`` (same on every rank)."""
    r = self._flag_ranges.get(key)
    if r is None:
      if self._flag_cursor + n > self.FLAG_WORDS:
        raise SpartanError('out of peer flag words')
      r = self._flag_ranges[key] = self._flag_cursor
      self._flag_cursor += n
    return r

  def flag_ptr(self, rank, index):
    return self.flags.ptrs[rank] + 4 * int(index)

  # ------------------------------------------------------------------ epochs and pushes
  def next_epoch(self, key):
    """A new exchange round over buffer ``key`` (same number on every rank); returns (epoch, pointer of a local word
    that holds it once the current stream reaches this point -- the word the copy engines send as the flag)."""
    e = self._epochs[key] = self._epochs.get(key, 0) + 1
    self._writes += 1
    src = self.epoch_src.data_ptr() + 4 * (self._writes % 64)
    check(lib.sp_write_u32(ctypes.c_void_p(src), e & 0xffffffff, self.ctx.stream_ptr()), 'sp_write_u32')
    return e, src

  def push(self, dsts, src_ptr, nbytes, flag_dsts, flag_src):
    """Copy-engine push of ``nbytes`` from ``src_ptr`` to every pointer in ``dsts`` on the current stream, each followed
    by the epoch word at ``flag_src`` into the matching ``flag_dsts`` entry."""
    n = len(dsts)
    if n == 0:
      return
    d = (ctypes.c_void_p * n)(*dsts)
    f = (ctypes.c_void_p * n)(*flag_dsts)
    check(lib.sp_peer_push(n, d, ctypes.c_void_p(src_ptr), int(nbytes), f, ctypes.c_void_p(flag_src),
                           self.ctx.stream_ptr()), 'sp_peer_push')

  def gate_timeouts(self):
    """Number of gated segments that gave up waiting (0 in a healthy job).  Synchronises."""
    return int(self.status.item()) if self.status is not None else 0

  def close(self):
    if self.ctx.device.type == 'cuda':
      torch.cuda.synchronize(self.ctx.device)
    for buf in list(self._buffers.values()) + ([self.flags] if self.flags is not None else []):
      try:
        self._release(buf)
      except Exception:
        pass
    self._buffers.clear()
    self.flags = None
    self._available = False
