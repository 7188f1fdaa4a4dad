import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
  if p not in sys.path:
    sys.path.insert(0, p)


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: needs a CUDA device (run with -m gpu on the B200 box)')


@pytest.fixture
def oracle3():
  """Fresh oracle context with the reference's default of 3 workers (spartan/config.py:131)."""
  import spartan_oracle
  from spartan_oracle import expr
  spartan_oracle.initialize(3)
  expr.eval_cache.clear()
  return spartan_oracle
