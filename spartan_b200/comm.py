"""torch.distributed plumbing (NCCL over NVLink/NVSwitch on GPUs, gloo on CPU for host-logic tests).

This is the whole replacement for the reference's RPC layer (spartan/rpc/*, ZeroMQ ROUTER/DEALER):
the cross-tile combiner becomes one all-reduce, tile movement becomes broadcast / send-recv.
With a single rank every function is a no-op.
"""
import os

import torch
import torch.distributed as dist

from ._lib import SP_RED_SUM, SP_RED_MIN, SP_RED_MAX, SP_RED_PROD, SP_RED_ALL, SP_RED_ANY


def init_from_env(device=None):
  """Returns (rank, world_size).  Joins the process group described by RANK / WORLD_SIZE /
  MASTER_ADDR / MASTER_PORT (as set by torchrun) if it is not initialised yet."""
  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  if world > 1 and not dist.is_initialized():
    use_cuda = torch.cuda.is_available() and (device is None or torch.device(device).type == 'cuda')
    if use_cuda:
      local = int(os.environ.get('LOCAL_RANK', rank))
      torch.cuda.set_device(local)
      dist.init_process_group(backend='nccl', device_id=torch.device('cuda', local))
    else:
      dist.init_process_group(backend='gloo')
  if dist.is_initialized():
    return dist.get_rank(), dist.get_world_size()
  return 0, 1


def world_size():
  return dist.get_world_size() if dist.is_initialized() else 1


def rank():
  return dist.get_rank() if dist.is_initialized() else 0


_REDUCE_OPS = {
  SP_RED_SUM: dist.ReduceOp.SUM, SP_RED_MIN: dist.ReduceOp.MIN, SP_RED_MAX: dist.ReduceOp.MAX,
  SP_RED_PROD: dist.ReduceOp.PRODUCT, SP_RED_ALL: dist.ReduceOp.MIN, SP_RED_ANY: dist.ReduceOp.MAX,
}


def allreduce(tensor, red_op):
  """Cross-rank combiner: replaces the N point-to-point ``update`` RPCs into an owner tile
  (blob_ctx.py:163-179, tile.pyx:263-283) with ncclAllReduce of the matching op."""
  if world_size() == 1:
    return tensor
  t = tensor
  if t.dtype == torch.bool:
    t = t.view(torch.uint8)
  dist.all_reduce(t, op=_REDUCE_OPS[red_op])
  return tensor


def broadcast(tensor, src):
  if world_size() > 1:
    t = tensor.view(torch.uint8) if tensor.dtype == torch.bool else tensor
    dist.broadcast(t, src=src)
  return tensor


def all_gather(out_list, tensor):
  if world_size() == 1:
    out_list[0].copy_(tensor)
  else:
    dist.all_gather(out_list, tensor)


def batch_p2p(ops):
  """ops: list of ('send'|'recv', tensor, peer).  Executes them as one NCCL group."""
  if not ops:
    return
  reqs = dist.batch_isend_irecv([dist.P2POp(dist.isend if kind == 'send' else dist.irecv, t, peer)
                                 for kind, t, peer in ops])
  for r in reqs:
    r.wait()


def barrier():
  if world_size() > 1:
    dist.barrier()
