"""Definitions of expressions and optimisations (reference: spartan/expr/__init__.py:26-142)."""
from .builtins import astype, size
from .builtins import empty, empty_like
from .builtins import zeros, zeros_like, ones, ones_like, full, full_like, arange, eye, identity
from .diag import diagonal, diagflat, diag, DiagonalExpr, DiagFlatExpr
from .builtins import all, any, equal, not_equal, greater, greater_equal, less, less_equal
from .builtins import logical_and, logical_or, logical_xor
from .builtins import add, sub, multiply, divide, true_divide, floor_divide
from .builtins import reciprocal, negative, fmod, mod, remainder
from .builtins import power, ln, log, square, sqrt, exp
from .builtins import abs, maximum, minimum, sum, prod
from .builtins import rand, randn
from .builtins import max, min, mean, std
from .builtins import count_nonzero, count_zero, argmin, argmax
from .dot import dot, DotExpr

from .base import Expr, evaluate, optimized_dag
from .base import eager, lazify, as_array, glom
from .base import NotShapeable, newaxis, Val, AsArray, ListExpr, TupleExpr
from ..array.distarray import broadcast
from .map import map, map_tiles, MapExpr, map_with_location, tile_mapper
from .map2 import (map2, outer, Map2Expr, OuterProductExpr, DeviceTileFunction, device_tile_function, dot_map2_mapper,
                   dot_outer_mapper, dot_as_join)
from .ndarray import ndarray, NdArrayExpr
from .optimize import optimize, MapMapFusion, ReduceMapFusion
from .tiling import AutomaticTiling
from .reduce import reduce, ReduceExpr, ArgReduceExpr
from .write_array import from_numpy, WriteArrayExpr
from .slice import SliceExpr
from .transpose import transpose, TransposeExpr
from .reshape import reshape, ravel, ReshapeExpr
from .fio import save, load, checkpoint, LoadExpr, CheckpointExpr
from .program import NotDeviceMappable
from . import local
import sys as _sys
_map_module = _sys.modules[__name__ + '.map']
_reduce_module = _sys.modules[__name__ + '.reduce']

# method-style access (expr/__init__.py:68-100)
Expr.all = all
Expr.any = any
Expr.argmax = argmax
Expr.argmin = argmin
Expr.astype = astype
Expr.dot = dot
Expr.fill = full_like
Expr.max = max
Expr.mean = mean
Expr.min = min
Expr.prod = prod
Expr.std = std
Expr.diagonal = diagonal
Expr.sum = sum
Expr.ravel = ravel
Expr.flatten = ravel
Expr.reshape = reshape
Expr.transpose = transpose
Expr.T = property(transpose)


class operator(object):
  """Namespace shim so ``expr.operator.local.LocalInput`` style paths of the reference resolve."""
  local = local
  map = _map_module
  reduce = _reduce_module
