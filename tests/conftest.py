"""pytest configuration: registers the `gpu` marker and common fixtures."""

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
  config.addinivalue_line(
      'markers', 'gpu: needs a CUDA device (run with -m gpu on the B200 box)')


@pytest.fixture(scope='session')
def mesh_golden():
  return np.load(os.path.join(GOLDEN_DIR, 'mesh_golden.npz'))


@pytest.fixture(scope='session')
def flow_golden():
  return np.load(os.path.join(GOLDEN_DIR, 'flow_golden.npz'))
