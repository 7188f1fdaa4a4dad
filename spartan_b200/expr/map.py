"""The ``map`` operation (reference: spartan/expr/operator/map.py:33-205 and
map_with_location.py:22-60).

``a + b`` is ``map((a, b), np.add)``.  Inputs are broadcast to a common shape; the kernel runs once
per tile of the largest input.  Where the reference's ``tile_mapper`` evaluates the LocalExpr tree
with one NumPy call (and one temporary) per node, the tree is compiled once per MapExpr to bytecode
(program.py) and each tile is ONE launch of the fused CUDA map kernel writing straight into the
output tile in HBM.
"""
from .. import blob_ctx, device_ops, util
from ..array import distarray
from ..array.distarray import Broadcast, broadcast, LocalWrapper
from ..core import LocalKernelResult
from . import program
from .base import ListExpr, Expr, as_array
from .local import LocalInput, LocalMapExpr, LocalMapLocationExpr, make_var


def bind_operands(children, child_to_var):
  """LocalInput name -> program.Operand for the evaluated children of a map / reduce."""
  operands = {}
  seen = {}
  for child, var in zip(children, child_to_var):
    base = child.base if isinstance(child, Broadcast) else child
    if isinstance(base, LocalWrapper) and base.shape == ():
      operands[var] = program.Operand('scalar', base.dtype, value=base.host_data()[()])
    else:
      # the same array used twice (x*x, (x-y)*x ...) is one operand: read once per element
      key = (id(base), tuple(child.shape))
      operands[var] = program.Operand('array', child.dtype, alias=seen.get(key))
      seen.setdefault(key, var)
  return operands


def get_local_values(ex, children, child_to_var, used_vars, owner):
  """Device tensors of every *used* operand for output extent ``ex`` (map.py:33-45).  Broadcast
  children hand over their un-expanded base region; the kernel broadcasts through zero strides."""
  values = {}
  for child, var in zip(children, child_to_var):
    if var not in used_vars:
      continue
    if isinstance(child, Broadcast):
      values[var] = child.fetch_base_tile(ex, dst=owner)
    else:
      values[var] = child.fetch(ex, dst=owner)
  return values


def blockwise_ok(largest, children):
  """A map / reduce can run one launch per contiguous block of the rank's slab (instead of one per tile)
  when every operand is either tiled and placed exactly like the largest one or is a value every rank
  holds: no operand region then needs stitching or communication."""
  if largest.slab is None:
    return False
  for child in children:
    base = child.base if isinstance(child, Broadcast) else child
    if base is largest or isinstance(base, LocalWrapper):
      continue
    if isinstance(child, Broadcast) or not largest.same_layout(base):
      return False
  return True


def slabwise_operands(largest, children, child_to_var, used_vars, free_axes=None):
  """When every array operand of a map / reduce is tiled and placed exactly like ``largest`` (or is ``largest``), the
  rank's whole share can be processed as ONE launch over the slabs themselves -- slab coordinates line up across
  operands, and an element-wise result does not care which global rows / columns a slab position stands for.
  ``free_axes`` lists the axes along which the slab may be a non-contiguous selection of the array (None = any axis,
  the case of a map; for a reduction only the reduced axis).  Returns the slab tensors in ``used_vars`` order, or None."""
  if largest.slab is None or not isinstance(largest, distarray.DistArrayImpl) or not largest.shape:
    return None
  for d, ivs in enumerate(largest.slab_axes):
    if free_axes is not None and d not in free_axes and not (len(ivs) >= 1 and ivs[0][0] == 0 and
                                                             ivs[-1][1] == largest.shape[d] and
                                                             all(a[1] == b[0] for a, b in zip(ivs, ivs[1:]))):
      return None
  slabs = {}
  for child, var in zip(children, child_to_var):
    if var not in used_vars:
      continue
    if child is largest:
      slabs[var] = largest.slab
    elif isinstance(child, distarray.DistArrayImpl) and largest.same_layout(child):
      slabs[var] = child.slab
    else:
      return None
  return [slabs[v] for v in used_vars]


def tile_mapper(ex, children, child_to_var, op, compiled, output):
  """Runs for each tile of a map (map.py:48-88): one fused kernel launch on the owning GPU."""
  ctx = blob_ctx.get()
  tile_id = output.tiles[ex]
  owner = tile_id.worker
  values = get_local_values(ex, children, child_to_var, compiled.used_vars, owner)
  if owner == ctx.worker_id:
    out_tile = ctx.tile(tile_id)
    inputs = [values[v] for v in compiled.used_vars]
    device_ops.run_map(compiled.program, inputs, out_tile.get(None))
    out_tile.valid = True
  return LocalKernelResult(result=[(ex, tile_id)])


class MapExpr(Expr):
  """Mapping an operator over one or more inputs (map.py:91-169)."""
  members = ('children', 'child_to_var', 'op')

  def pretty_str(self):
    return 'Map[%d](%s, %s)' % (self.expr_id, self.op.pretty_str(), self.children)

  def compute_shape(self):
    # map.py:104-128
    orig_shapes = [list(x.shape) for x in self.children]
    max_dim = max(len(s) for s in orig_shapes)
    new_shapes = [[1] * (max_dim - len(s)) + s for s in orig_shapes]
    out = []
    for i in range(max_dim):
      sizes = [s[i] for s in new_shapes]
      out.append(next((v for v in sizes if v != 1), 1))     # the size that is not 1 (it may be 0)
    return tuple(out)

  def _evaluate(self, ctx, deps):
    # map.py:149-169
    children = list(deps['children'])
    child_to_var = list(deps['child_to_var'])
    children = broadcast(children)
    largest = distarray.largest_value(children)
    i = children.index(largest)
    children[0], children[i] = children[i], children[0]
    child_to_var[0], child_to_var[i] = child_to_var[i], child_to_var[0]
    if not (isinstance(largest, distarray.DistArrayImpl) or getattr(largest, 'is_view', False)):
      raise program.NotDeviceMappable('a map needs at least one distributed (non-broadcast) input')

    loc_kernel = getattr(self.op.fn, 'device_location_kernel', None) \
      if isinstance(self.op, LocalMapLocationExpr) else None
    if loc_kernel is not None:
      # map_with_location over a freshly created array (arange, eye...): an extent-aware fill kernel
      out_dtype = self.op.fn.result_dtype(largest.dtype, self.op.kw)
      output = distarray.create_like(largest, out_dtype)
      for ex, tid in sorted(output.tiles.items(), key=lambda kv: (kv[1].worker, kv[1].id)):
        if ctx.is_local(tid):
          t = ctx.tile(tid)
          loc_kernel(t.get(None), ex, **(self.op.kw or {}))
          t.valid = True
      return output

    compiled = program.compile_tree(self.op, bind_operands(children, child_to_var))
    output = distarray.create_like(largest, compiled.out_dtype)
    if blockwise_ok(largest, children) and output.slab is not None:
      # one fused launch per rank when all operands share the layout, else one per contiguous block of this rank's
      # share (the tiling is the reference's unit of RPC dispatch, not a unit of work the GPU needs)
      slabs = slabwise_operands(largest, children, child_to_var, compiled.used_vars)
      if slabs is not None and output.slab.shape == largest.slab.shape:
        if output.slab.numel():
          device_ops.run_map(compiled.program, slabs, output.slab)
        for tid in output.tiles.values():
          if ctx.is_local(tid):
            ctx.tile(tid).valid = True
        return output
      for block in largest.local_blocks():
        values = get_local_values(block, children, child_to_var, compiled.used_vars, ctx.worker_id)
        device_ops.run_map(compiled.program, [values[v] for v in compiled.used_vars], output.slab_view(block))
      for tid in output.tiles.values():
        if ctx.is_local(tid):
          ctx.tile(tid).valid = True
      return output
    largest.foreach_tile(tile_mapper, kw={'children': children, 'child_to_var': child_to_var, 'op': self.op,
                                          'compiled': compiled, 'output': output})
    return output


def map(inputs, fn, numpy_expr=None, fn_kw=None):
  """Evaluate ``fn`` over each tile of the input (map.py:172-205)."""
  assert fn is not None
  if not util.is_iterable(inputs):
    inputs = [inputs]
  op_deps, children, child_to_var = [], [], []
  for v in inputs:
    v = as_array(v)
    varname = make_var()
    children.append(v)
    child_to_var.append(varname)
    op_deps.append(LocalInput(idx=varname))
  op = LocalMapExpr(fn=fn, kw=fn_kw, pretty_fn=numpy_expr, deps=op_deps)
  return MapExpr(children=ListExpr(vals=children), child_to_var=child_to_var, op=op)


map_tiles = map   # the name BASELINE.json's north_star uses


def map_with_location(inputs, fn, numpy_expr=None, fn_kw=None):
  """map_with_location.py:22-60.  On the device only location kernels built into the library
  (arange, ...) are supported: ``fn`` must carry a ``device_location_kernel``."""
  assert fn is not None
  if not util.is_iterable(inputs):
    inputs = [inputs]
  op_deps, children, child_to_var = [], [], []
  for v in inputs:
    v = as_array(v)
    varname = make_var()
    children.append(v)
    child_to_var.append(varname)
    op_deps.append(LocalInput(idx=varname))
  op_deps += [LocalInput(idx='extent')]
  op = LocalMapLocationExpr(fn=fn, kw=fn_kw, pretty_fn=numpy_expr, deps=op_deps)
  return MapExpr(children=ListExpr(vals=children), child_to_var=child_to_var, op=op)


from .map2 import map2, outer          # noqa: E402,F401  (map.py:337-375, outer.py:102-120)
