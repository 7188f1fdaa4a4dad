"""On-disk tile format and checkpoints (reference: spartan/expr/fio.py:46-232 save / load, operator/checkpoint.py:11-59).

The byte format is the reference's, so arrays written by either side load on the other: a directory
``<path>/<prefix>/`` with ``<prefix>_dist.spf`` (text: shape, tile shape, dtype, DENSITY) and one file per tile,
``<prefix>_<ul>_<lr>_spf`` (``..._spfbz2`` when zipped): the npy magic ``\\x93NUMPY\\x01\\x00``, a little-endian
16-bit header length, a Python-literal dict (ul, lr, shape, dtype, type) padded to a 16-byte boundary, then the tile's
raw C-order bytes.  Every rank writes / reads the tiles it owns, straight from / into its HBM slab (one pitched D2H or
H2D per tile); rank 0 writes the array-wide file.  Sparse tiles and the pickle variants are outside the device path.
"""
import ast
import bz2
import os

import numpy as np
import torch

from .. import blob_ctx, comm, device_ops
from ..array import distarray
from ..config import FLAGS
from .._lib import SpartanError
from .base import Expr, lazify, evaluate

MAGIC = b'\x93NUMPY\x01\x00'


def save_filename(path, prefix, ul, lr, iszip=False, suffix=''):
  """fio.py:46-63 (dense, non-pickle)."""
  fn = path + '/' + prefix + '/' + prefix + '_' + str(tuple(ul)) + '_' + str(tuple(lr))
  if suffix != '':
    fn += '_' + suffix
  fn += '_spf'
  if iszip:
    fn += 'bz2'
  return fn


def _tile_header(ex, shape, dtype):
  """fio.py:70-81."""
  tile_dict = {'ul': tuple(ex.ul), 'lr': tuple(ex.lr), 'shape': tuple(shape), 'dtype': str(np.dtype(dtype)),
               'type': 'DENSITY'}
  dict_cnt = str(tile_dict)
  if (len(MAGIC) + 2 + len(dict_cnt)) % 16 != 0:
    dict_cnt += (16 - (len(MAGIC) + 2 + len(dict_cnt)) % 16) * ' '
  return MAGIC + bytes([len(dict_cnt) % 256, len(dict_cnt) // 256]) + dict_cnt.encode('ascii')


def _write_dist(path, prefix, array):
  """fio.py:114-131."""
  with open(path + '/' + prefix + '/' + prefix + '_dist.spf', 'w') as fp:
    fp.write(''.join(str(d) + ' ' for d in array.shape) + '\n')
    fp.write(''.join(str(d) + ' ' for d in array.tile_shape()) + '\n')
    fp.write(str(array.dtype) + '\n')
    fp.write('DENSITY\n')


def save(array, prefix, path='.', iszip=False):
  """Save ``array`` (an Expr or a DistArray) under ``path/prefix`` (fio.py:134-156).  Not lazy; returns True."""
  ctx = blob_ctx.get()
  array = evaluate(array) if isinstance(array, Expr) else array
  if not isinstance(array, distarray.DistArrayImpl):
    raise SpartanError('save() needs a materialised distributed array (evaluate views with a map first)')
  os.makedirs(path + '/' + prefix, exist_ok=True)
  if ctx.worker_id == 0:
    _write_dist(path, prefix, array)
  for ex, tid in array.tiles.items():
    if not ctx.is_local(tid):
      continue
    t = ctx.get(tid, None)
    host = np.empty(tuple(t.shape), dtype=array.dtype)
    if t.device.type == 'cuda' and t.dim() <= 2:
      device_ops.download_rect(host, t)                     # pitched D2H out of the slab
      torch.cuda.current_stream(ctx.device).synchronize()
    else:
      host[...] = t.cpu().numpy()
    fn = save_filename(path, prefix, ex.ul, ex.lr, iszip)
    fp = bz2.BZ2File(fn, 'w', compresslevel=1) if iszip else open(fn, 'wb')
    try:
      fp.write(_tile_header(ex, host.shape, array.dtype))
      fp.write(host.tobytes())
    finally:
      fp.close()
  comm.barrier()
  return True


def _read_dist(path, prefix):
  """fio.py:190-208."""
  fn = path + '/' + prefix + '/' + prefix + '_dist.spf'
  if not os.path.exists(fn):
    raise IOError(fn)
  with open(fn) as fp:
    shape = [int(i) for i in fp.readline().strip().split()]
    tile_hint = [int(i) for i in fp.readline().strip().split()]
    dtype = np.dtype(''.join(fp.readline().strip()))
    sparse = fp.readline().find('SPARSE') != -1
  return {'shape': shape, 'sparse': sparse, 'dtype': dtype, 'tile_hint': tile_hint}


def read_tile(path, prefix, ex, dtype, iszip=False):
  """fio.py:159-187 _load_mapper (dense): the tile's data as a host ndarray."""
  fn = save_filename(path, prefix, ex.ul, ex.lr, iszip)
  fp = bz2.BZ2File(fn, 'r') if iszip else open(fn, 'rb')
  try:
    fp.read(8)                                     # magic number and version
    dlen = fp.read(2)
    ast.literal_eval(fp.read(dlen[0] + dlen[1] * 256).decode('ascii'))   # redundant with _dist.spf, kept for the format
    data = np.frombuffer(bytearray(fp.read()), dtype=dtype) if iszip else np.fromfile(fp, dtype=dtype)
  finally:
    fp.close()
  return data.reshape(ex.shape)


class LoadExpr(Expr):
  """``load`` as a node: a new array tiled as recorded, every rank filling its own tiles from disk (fio.py:211-232;
  the reference routes the same mapper through ``shuffle``)."""
  members = ('prefix', 'path', 'iszip', 'info')

  def visit(self, visitor):
    return self

  def dependencies(self):
    return {}

  def compute_shape(self):
    return tuple(self.info['shape'])

  def _evaluate(self, ctx, deps):
    info = self.info
    if info['sparse']:
      raise SpartanError('sparse arrays are outside the device path')
    arr = distarray.create(tuple(info['shape']), info['dtype'], tile_hint=info['tile_hint'] or None)
    for ex, tid in arr.tiles.items():
      if not ctx.is_local(tid):
        continue
      t = ctx.tile(tid)
      data = read_tile(self.path, self.prefix, ex, info['dtype'], self.iszip)
      dst = t.get(None)
      if dst.dim() <= 2:
        device_ops.upload_rect(dst, data)
      else:
        dst.copy_(torch.from_numpy(np.ascontiguousarray(data)))
      t.valid = True
    if ctx.device.type == 'cuda':
      torch.cuda.current_stream(ctx.device).synchronize()     # the host buffers go away after this call
    return arr


def load(prefix, path='.', iszip=False):
  """Load ``path/prefix`` into a new array (lazy; fio.py:211-232)."""
  return LoadExpr(prefix=prefix, path=path, iszip=iszip, info=_read_dist(path, prefix))


class CheckpointExpr(Expr):
  """operator/checkpoint.py:11-47, 'disk' mode: evaluating the node writes its source to disk once; ``load_data``
  brings it back (e.g. after the cached result was dropped)."""
  members = ('src', 'path', 'mode', 'ready')

  def dependencies(self):
    return {'src': self.src}

  def compute_shape(self):
    return self.src.shape

  def load_data(self, cached_result=None):
    if not self.ready or self.mode != 'disk':
      return None
    return load('%s' % self.expr_id, path=self.path, iszip=False).evaluate()

  def _evaluate(self, ctx, deps):
    result = deps['src']
    if self.mode == 'disk':
      save(result, '%s' % self.expr_id, path=self.path, iszip=False)
    self.ready = True
    return result


def checkpoint(x, mode='disk'):
  """Make a checkpoint for ``x`` (operator/checkpoint.py:50-59)."""
  return CheckpointExpr(src=lazify(x), path=getattr(FLAGS, 'checkpoint_path', '/tmp/spartan/checkpoint'), mode=mode,
                        ready=False)
