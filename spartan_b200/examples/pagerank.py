"""PageRank's sparse matrix x vector step on the device (reference: tests/benchmark_pagerank.py:11-45).

The reference builds ``wts`` with a ``shuffle`` whose mapper ``make_weights`` draws, for every column strip,
``strip_width * OUTLINKS_PER_PAGE`` random (dest row, source column, weight) triples on the strip's worker, and then
times ``expr.dot(wts, p).evaluate()``.  Here every rank draws the triples of the strips it owns on its own GPU
(seeded per strip, so the matrix does not depend on the number of ranks) and keeps them as CSR column strips
(spartan_b200/sparse.py); the product is ``sp.dot(wts, p)``."""
import numpy as np
import torch

from .. import blob_ctx, sparse

OUTLINKS_PER_PAGE = 10          # benchmark_pagerank.py:7


def strip_entries(n, c0, c1, outlinks, seed, device):
  """The non-zeros of column strip [c0, c1): (rows, global cols, values) -- make_weights (benchmark_pagerank.py:11-24)
  with a generator seeded by (seed, c0)."""
  g = torch.Generator(device=device)
  g.manual_seed((int(seed) * 1000003 + int(c0)) & 0x7fffffffffffffff)
  num_out = (c1 - c0) * outlinks
  rows = torch.randint(0, n, (num_out,), generator=g, device=device, dtype=torch.int64)
  cols = torch.randint(c0, c1, (num_out,), generator=g, device=device, dtype=torch.int64)
  vals = torch.rand((num_out,), generator=g, device=device, dtype=torch.float32)
  return rows, cols, vals


def make_weights(n, strip_width, outlinks=OUTLINKS_PER_PAGE, seed=0):
  """The (n, n) link matrix as column strips of ``strip_width`` columns, generated on the devices."""
  ctx = blob_ctx.get()
  return sparse.from_device_coo((n, n), strip_width,
                                lambda c0, c1: strip_entries(n, c0, c1, outlinks, seed, ctx.device))
