"""Separates the costs inside the multi-GPU resident dot at the per-rank shape of the 8-GPU run
(C[32768, 4096] = sum over 8 K-segments of A_p[32768, 4096] . B_p[4096, 4096]), on TWO GPUs: each rank pushes the
seven remote segments (7 x 512 MiB) to its single peer, so every GPU sees the in/out copy-engine traffic of the
8-rank ring.  Cases: GEMM alone (operands present) / pushes alone / gated GEMM with the pushes in flight, for
sequential pushes and for pushes spread over several streams.
  torchrun --nproc-per-node 2 tools/peer_gemm_probe.py"""
import ctypes, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch.distributed as dist
import spartan_b200 as sp
from spartan_b200 import device_ops, comm
from spartan_b200._lib import lib, check

ctx = sp.initialize()
me, W = ctx.worker_id, ctx.num_workers
assert W == 2
peer = ctx.peer
assert peer.available()
other = 1 - me
M, Nc, Ks, S = 32768, 4096, 4096, 8
prec = 'bf16x3'
Kp = device_ops.gemm_kpad(Ks, prec)
a_bytes = device_ops.gemm_prepared_bytes(M, Kp, prec)
rb = device_ops.gemm_row_bytes(Kp, prec)
gather = peer.buffer('probe_gather', S * a_bytes)
f0 = peer.flag_range('probe', S)
src = torch.rand(M, Ks, device=ctx.device)
mine = device_ops.PreparedOperand(M, Ks, prec, None)
mine.prepare_a(src, 0)
for p in range(S):      # fill every slot locally so that "GEMM alone" has real operands
  gather.tensor[p * a_bytes:(p + 1) * a_bytes].copy_(mine.buf)
bsrc = torch.rand(Ks, Nc, device=ctx.device)
pbs = []
for p in range(S):
  pb = device_ops.PreparedOperand(Nc, Ks, prec, None)
  pb.prepare_b(bsrc, 0)
  pbs.append(pb)
C = torch.empty(M, Nc, device=ctx.device)
main = torch.cuda.current_stream()
streams = [torch.cuda.Stream() for _ in range(7)]
epoch = [0]


def launch(gated, e):
  views, flags = [], []
  for p in range(S):
    a_ptr = mine.row_ptr(0) if p == 0 else gather.local_ptr + p * a_bytes
    views.append((a_ptr, M * rb, pbs[p].row_ptr(0), pbs[p].copy_stride, Kp))
    flags.append(0 if (p == 0 or not gated) else peer.flag_ptr(me, f0 + p))
  device_ops.gemm_prepared_views_gated(views, flags, [e] * S, peer.status.data_ptr(), C, False, prec)


def pushes(nstreams, e, esrc):
  ev = main.record_event()
  for j in range(1, S):
    st = streams[(j - 1) % nstreams]
    with torch.cuda.stream(st):
      st.wait_event(ev)
      peer.push([gather.ptrs[other] + j * a_bytes], mine.buf.data_ptr(), a_bytes, [peer.flag_ptr(other, f0 + j)], esrc)
  for st in streams[:nstreams]:
    main.wait_stream(st) if False else None


def timed(fn, n=5):
  for _ in range(2): fn()
  comm.barrier(); torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(n): fn()
  for st in streams: main.wait_stream(st)
  e1.record(); comm.barrier(); torch.cuda.synchronize()
  return e0.elapsed_time(e1) / n


def case_gemm_alone():
  launch(False, 0)


def make_case(nstreams, with_gemm):
  def fn():
    e, esrc = peer.next_epoch('probe')
    pushes(nstreams, e, esrc)
    if with_gemm:
      launch(True, e)
    else:
      for st in streams[:nstreams]: main.wait_stream(st)
  return fn

def traced(nstreams, gated):
  """One step with events: when each pushed segment has left this rank, and when the GEMM ends (ms from the start)."""
  comm.barrier(); torch.cuda.synchronize()
  t0 = torch.cuda.Event(enable_timing=True); t0.record()
  e, esrc = peer.next_epoch('probe')
  ev = main.record_event()
  marks = []
  for j in range(1, S):
    st = streams[(j - 1) % nstreams]
    with torch.cuda.stream(st):
      st.wait_event(ev)
      peer.push([gather.ptrs[other] + j * a_bytes], mine.buf.data_ptr(), a_bytes, [peer.flag_ptr(other, f0 + j)], esrc)
      m = torch.cuda.Event(enable_timing=True); m.record(st); marks.append(m)
  launch(gated, e)
  g = torch.cuda.Event(enable_timing=True); g.record(main)
  for st in streams: main.wait_stream(st)
  comm.barrier(); torch.cuda.synchronize()
  return {'push_done_ms': [round(t0.elapsed_time(m), 2) for m in marks], 'gemm_done_ms': round(t0.elapsed_time(g), 2)}


out = {'gemm_alone_ms': timed(case_gemm_alone)}
for ns in (1, 2):
  for gated in (False, True):
    traced(ns, gated)
    out['trace_%dstreams_%s' % (ns, 'gated' if gated else 'ungated')] = traced(ns, gated)
lib.sp_gemm_set_round_sync(0)
out['gemm_alone_no_round_sync_ms'] = timed(case_gemm_alone)
out['gated_with_pushes_1stream_no_round_sync_ms'] = timed(make_case(1, True))
lib.sp_gemm_set_round_sync(1)
for ns in (1, 2, 4, 7):
  out['pushes_alone_%dstreams_ms' % ns] = timed(make_case(ns, False))
  out['gated_gemm_with_pushes_%dstreams_ms' % ns] = timed(make_case(ns, True))
out['push_bytes_per_step'] = 7 * a_bytes
out['gate_timeouts'] = peer.gate_timeouts()
if me == 0:
  print(json.dumps(out), flush=True)
comm.barrier()
sp.shutdown()
dist.destroy_process_group()
