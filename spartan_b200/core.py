"""TileId / LocalKernelResult PODs (reference: spartan/core.pyx:16-41, :159-169)."""
import collections

TileId = collections.namedtuple('TileId', ['worker', 'id'])


class LocalKernelResult(object):
  """What a per-tile mapper returns: ``result`` = [(extent, TileId), ...] (or [] for reducers that
  wrote into a target array); ``futures`` is kept for signature parity and is always None here --
  launches are asynchronous on the device stream instead of RPC futures."""
  __slots__ = ('result', 'futures')

  def __init__(self, result=None, futures=None):
    self.result = result if result is not None else []
    self.futures = futures
