"""Builds the REFERENCE's own extent algebra (spartan/array/extent.pyx) in place with Cython,
Python-2 language level, output only into oracle/_ref/.  No reference source is copied into the repo.

  python oracle/ref_extent/build_ref_extent.py          # -> oracle/_ref/ref_extent*.so

Only usable where /root/reference exists (this container); the GPU box uses the committed vectors in
tests/golden/extent_vectors.json instead.
"""
import os
import shutil
import subprocess
import sys
import sysconfig

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '..', '_ref')
SRC = '/root/reference/spartan/array/extent.pyx'


def build():
  if not os.path.exists(SRC):
    print('reference not present; skipping'); return None
  os.makedirs(OUT, exist_ok=True)
  c_file = os.path.join(OUT, 'ref_extent.c')
  # module name must match the PyInit symbol: cythonize under the name ref_extent
  subprocess.check_call([sys.executable, '-m', 'cython', '-2', '--module-name', 'ref_extent', SRC, '-o', c_file])
  ext = sysconfig.get_config_var('EXT_SUFFIX')
  so = os.path.join(OUT, 'ref_extent' + ext)
  inc = [sysconfig.get_paths()['include'], np.get_include()]
  cmd = ['gcc', '-O2', '-shared', '-fPIC', '-w', c_file, '-o', so] + ['-I' + i for i in inc]
  subprocess.check_call(cmd)
  return so


if __name__ == '__main__':
  print(build())
