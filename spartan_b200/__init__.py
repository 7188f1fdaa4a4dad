"""spartan_b200 -- a B200-native execution backend for Spartan-style tiled-array expressions.

Keeps the expression-graph surface of spartan-array/spartan (``ones / zeros / rand / arange /
from_numpy / dot / sum / map / reduce``, ``MapExpr / ReduceExpr / DotExpr``, ``tile_hint``,
``blob_ctx``) and replaces the master/worker runtime and the per-tile NumPy evaluator with tiles
resident in HBM and hand-written sm_100a CUDA kernels behind a C ABI (include/spartan_b200.h).
"""
from . import _lib                       # loads libspartan_b200.so; ImportError if it is missing
from . import blob_ctx, config, util, comm
from .config import FLAGS
from ._lib import SpartanError
from .expr import *                      # noqa: F401,F403  (spartan/__init__.py:33 does the same)
from .expr import map, sum, min, max, abs, all, any   # noqa: F401  names that shadow builtins on purpose
from . import expr
from .array import distarray, extent, tile
from . import sparse
from .examples.kmeans import KMeans
from .replay import replayable, Replayable


def initialize(argv=None, device=None):
  """Set up this rank (spartan/__init__.py:42-56).  Idempotent."""
  return blob_ctx.initialize(device=device)


def shutdown():
  """spartan/__init__.py:59-62."""
  from .expr.base import eval_cache
  eval_cache.clear()
  blob_ctx.shutdown()
