"""Generates tests/golden/fio_vectors.json: the bytes of the tile files of two small arrays as written by the ORACLE's
restatement of the reference writer (spartan/expr/fio.py:66-131; the reference itself is Python 2 and cannot run here).
The product's writer and both loaders are checked against these bytes (tests/test_host_cpu.py), so the on-disk format
cannot drift.  Run:  python tests/golden/make_fio_vectors.py"""
import base64
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import spartan_oracle
from spartan_oracle import distarray, extent, fio

spartan_oracle.initialize(1)
cases = []
for name, shape, hint, dtype in [('f32_grid', (10, 7), (4, 5), 'float32'), ('i64_strips', (9,), (4,), 'int64')]:
  x = (np.arange(int(np.prod(shape))).reshape(shape) * 3 - 7).astype(dtype)
  arr = distarray.create(shape, np.dtype(dtype), tile_hint=hint)
  arr.update(extent.from_shape(shape), x)
  d = tempfile.mkdtemp()
  fio.save(arr, name, d, False)
  files = {}
  for fn in sorted(os.listdir(os.path.join(d, name))):
    files[fn] = base64.b64encode(open(os.path.join(d, name, fn), 'rb').read()).decode('ascii')
  cases.append({'prefix': name, 'shape': list(shape), 'tile_hint': list(hint), 'dtype': dtype,
                'data': x.ravel().tolist(), 'files': files})
out = os.path.join(ROOT, 'tests', 'golden', 'fio_vectors.json')
json.dump({'generated_by': 'tests/golden/make_fio_vectors.py (oracle restatement of spartan/expr/fio.py)', 'cases': cases},
          open(out, 'w'), indent=1)
print(out, sum(len(c['files']) for c in cases), 'files')
