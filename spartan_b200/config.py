"""The handful of flags the hot path reads (reference: spartan/config.py:27-227, optimize.py:1084-1101).
``num_workers`` becomes the number of GPU ranks (one process per GPU)."""


import os


class Flags(object):
  def __init__(self):
    self.optimization = True            # optimize.py:1101
    self.opt_map_fusion = True          # optimize.py:1096
    self.opt_reduce_fusion = True       # optimize.py:1099
    # optimize.py:1094 (default True there); here arrays without a tile_hint keep the reference's default tiling unless
    # this is switched on -- results are identical either way, only placement (and NVLink traffic) changes
    self.opt_auto_tiling = False
    self.opt_expression_cache = True    # base.py:21
    self.tile_assignment_strategy = 'round_robin'   # distarray.py:441-445 (the only strategy on one box)
    # tensor-core mode of dot(): 'bf16x3' (default: 16-bit-mantissa split, 3 bf16 passes), 'tf32x3' (22-bit),
    # 'tf32x1' (fastest, 11-bit operands), 'simt' (CUDA cores, plain IEEE fp32)
    self.dot_precision = 'bf16x3'
    # dot(from_numpy(a), from_numpy(b)) on one GPU: pipeline the PCIe upload of host operands with the contraction
    # (row strips of A / column strips of B of this many rows / columns) when the operands are at least this large
    self.dot_stream_host_operands = True
    self.dot_stream_strip = 2048         # measured: 196.6 ms at 2048 vs 199-202 ms at 1024 / 3072 / 4096 (tools/e2e_strip_probe.py)
    self.dot_stream_min_bytes = 256 << 20
    # prepared (rounded / split / transposed) GEMM operands of unchanged arrays are kept between evaluations, up to this
    # many bytes (0 disables the cache)
    # multi-GPU dot: 'auto' = segments contracted in passes grouped by arrival (dot.py _arrival_groups), 'single' = one
    # gated launch over all segments (tuning aid)
    self.dot_passes = os.environ.get('SPARTAN_DOT_PASSES', 'auto')
    self.dot_trace = False              # record a CUDA-event timeline of streamed multi-GPU dots (diagnostic)
    self.dot_prepared_cache = True
    self.dot_prepared_cache_bytes = 48 << 30
    self.checkpoint_path = '/tmp/spartan/checkpoint'     # config.py:96 default checkpoint directory

  def __repr__(self):
    return 'FLAGS(%s)' % ', '.join('%s=%r' % kv for kv in sorted(self.__dict__.items()))


FLAGS = Flags()
