"""ORACLE (test infrastructure only) -- the reference's on-disk tile format, dense arrays.

Restates spartan/expr/fio.py: save_filename :46-63, _save_reducer :66-111 (dense branch), _save :114-131,
save :134-156, _load_mapper :159-187 (dense branch), _load :190-208, load :211-232.  Python 2's ``str`` / ``chr``
byte strings become ``bytes`` here; the bytes on disk are the same.
"""
import ast
import bz2
import os

import numpy as np

from . import distarray, extent
from .expr import Expr, evaluate


def save_filename(**kw):
  # fio.py:46-63
  fn = kw['path'] + '/' + kw['prefix'] + '/' + kw['prefix'] + '_' + str(kw['ul']) + '_' + str(kw['lr'])
  if kw['suffix'] != '':
    fn += '_' + kw['suffix']
  if not kw['isnp']:
    fn += '_' + 'sp'
    fn += 'p' if kw['ispickle'] else 'f'
    if kw['iszip']:
      fn += 'bz2'
  return fn


def _save_reducer(ex, tile, path=None, prefix=None, iszip=None):
  # fio.py:66-111, dense
  if not os.path.exists(path + '/' + prefix):
    os.makedirs(path + '/' + prefix)
  tile_dict = {'ul': ex.ul, 'lr': ex.lr, 'shape': tile.shape, 'dtype': str(tile.dtype), 'type': 'DENSITY'}
  cnt = b'\x93NUMPY\x01\x00'
  dict_cnt = str(tile_dict)
  if (len(cnt) + 2 + len(dict_cnt)) % 16 != 0:
    dict_cnt += str((16 - (len(cnt) + 2 + len(dict_cnt)) % 16) * ' ')
  cnt += bytes([len(dict_cnt) % 256, len(dict_cnt) // 256]) + dict_cnt.encode('ascii')
  kw = {'path': path, 'prefix': prefix, 'suffix': '', 'ul': ex.ul, 'lr': ex.lr, 'ispickle': False, 'isnp': False,
        'iszip': bool(iszip)}
  fn = save_filename(**kw)
  fp = bz2.BZ2File(fn, 'w', compresslevel=1) if iszip else open(fn, 'wb')
  fp.write(cnt)
  fp.write(np.ascontiguousarray(tile).data)
  fp.close()
  return np.asarray(1)


def _save(path, prefix, array, iszip):
  # fio.py:114-131
  path = path + '/' + prefix
  if not os.path.exists(path):
    os.makedirs(path)
  with open(path + '/' + prefix + '_dist.spf', 'w') as fp:
    for dim in array.shape:
      fp.write(str(dim) + ' ')
    fp.write('\n')
    for dim in array.tile_shape():
      fp.write(str(dim) + ' ')
    fp.write('\n')
    fp.write(str(array.dtype) + '\n')
    fp.write('DENSITY\n')


def save(array, prefix, path='.', iszip=False):
  # fio.py:134-156: reduce(array, None, _save_reducer, np.multiply) == 1
  array = evaluate(array) if isinstance(array, Expr) else array
  _save(path, prefix, array, iszip)
  ret = 1
  for ex in array.tiles:
    ret *= int(_save_reducer(ex, array.fetch(ex), path=path, prefix=prefix, iszip=iszip))
  return ret == 1


def _load_mapper(ex, prefix=None, path=None, dtype=None, iszip=None):
  # fio.py:159-187, dense
  kw = {'path': path, 'prefix': prefix, 'suffix': '', 'ul': ex.ul, 'lr': ex.lr, 'ispickle': False, 'isnp': False,
        'iszip': bool(iszip)}
  fn = save_filename(**kw)
  fp = bz2.BZ2File(fn, 'r') if iszip else open(fn, 'rb')
  fp.read(8)
  dlen = fp.read(2)
  dlen = dlen[0] + dlen[1] * 256
  ast.literal_eval(fp.read(dlen).decode('ascii'))
  data = np.frombuffer(fp.read(), dtype=dtype).copy()
  data.shape = ex.shape
  fp.close()
  return data


def _load(path, prefix, iszip):
  # fio.py:190-208
  fn = path + '/' + prefix + '/' + prefix + '_dist.spf'
  if not os.path.exists(fn):
    raise IOError
  with open(fn) as fp:
    shape = [int(i) for i in fp.readline().strip().split()]
    tile_hint = [int(i) for i in fp.readline().strip().split()]
    dtype = np.dtype(''.join(fp.readline().strip()))
    sparse = fp.readline().find('SPARSE') != -1
  return {'shape': shape, 'sparse': sparse, 'dtype': dtype, 'tile_hint': tile_hint}


def load(prefix, path='.', iszip=False):
  # fio.py:211-232 (evaluated eagerly: returns the DistArray)
  info = _load(path, prefix, iszip)
  arr = distarray.create(tuple(info['shape']), info['dtype'], tile_hint=info['tile_hint'])
  for ex in arr.tiles:
    arr.update(ex, _load_mapper(ex, prefix=prefix, path=path, dtype=info['dtype'], iszip=iszip))
  return arr
