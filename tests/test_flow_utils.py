"""sofima_b200.flow_utils (host filters, as in the reference) against golden vectors
from the reference's own module and the reference's KATs (tests/flow_utils_test.py)."""

import os

import numpy as np
import pytest

from sofima_b200 import flow_utils
from sofima_b200.decorators import flow as dflow

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'flow_utils_golden.npz')


@pytest.fixture(scope='module')
def g():
  return np.load(GOLDEN)


def test_apply_mask():  # tests/flow_utils_test.py:26-37
  flow = np.zeros((3, 1, 50, 50))
  mask = np.zeros((1, 50, 50), dtype=bool)
  mask[0, 10, 15] = True
  mask[0, 3, 4] = True
  flow_utils.apply_mask(flow, mask)
  expected = np.zeros((3, 1, 50, 50))
  expected[:, 0, 10, 15] = np.nan
  expected[:, 0, 3, 4] = np.nan
  np.testing.assert_array_equal(flow, expected)


def test_clean_flow_kat():  # tests/flow_utils_test.py:39-66
  flow = np.zeros((4, 1, 50, 40))
  flow[2, ...] = 2.0
  flow[2, 0, 10, 20] = 1.2
  flow[3, 0, 10, 22] = 1.2
  flow[3, 0, 10, 24] = 1.6
  flow[0, 0, 5, 4] = 12
  flow[1, 0, 5, 6] = -14
  flow[:, 0, 5, 10] = 2
  flow[:, 0, 15, 10] = 7
  cleaned = flow_utils.clean_flow(flow, min_peak_ratio=1.4, min_peak_sharpness=1.6,
                                  max_magnitude=10, max_deviation=5)
  expected = np.zeros((2, 1, 50, 40))
  expected[:, 0, 5, 10] = 2
  expected[:, 0, 15, 10] = np.nan  # median filter
  expected[:, 0, 10, 20] = np.nan  # peak sharpness
  expected[:, 0, 10, 22] = np.nan  # peak ratio
  expected[:, 0, 5, 4] = np.nan  # magnitude
  expected[:, 0, 5, 6] = np.nan  # magnitude
  np.testing.assert_array_equal(cleaned, expected)


def test_reconcile_flows_kat():  # tests/flow_utils_test.py:68-96
  flow1 = np.full((3, 1, 50, 40), np.nan)
  flow2 = np.full((3, 1, 50, 40), np.nan)
  flow3 = np.full((3, 1, 50, 40), np.nan)
  flow1[:, 0, 10, 10] = 2.
  flow2[:, 0, 10, 10] = 3.  # ignored, flow1 preferred
  flow3[:, 0, 20, 20] = 4.
  flow2[:, 0, 20, 20] = 1.  # ignored, min_delta_z
  flow2[:, 0, 30:35, 30:35] = 5
  flow2[0, 0, 32, 32] = 15  # ignored, max_deviation
  got = flow_utils.reconcile_flows([flow1, flow2, flow3], max_gradient=0, max_deviation=8,
                                   min_patch_size=0, min_delta_z=2)
  expected = np.full((3, 1, 50, 40), np.nan)
  expected[:, 0, 10, 10] = 2.
  expected[:, 0, 20, 20] = 4.
  expected[:, 0, 30:35, 30:35] = 5
  expected[:, 0, 32, 32] = np.nan
  np.testing.assert_array_equal(got, expected)


def test_clean_flow_golden(g):
  f4 = g['clean_in']
  np.testing.assert_array_equal(
      flow_utils.clean_flow(f4.copy(), 1.4, 0.6, 25.0, 4.0), g['clean_out'])
  np.testing.assert_array_equal(flow_utils.clean_flow(f4.copy(), 1.4, 0.6, 0, 0),
                                g['clean_out_nodev'])
  np.testing.assert_array_equal(flow_utils.clean_flow(f4[:2].copy(), 0, 0, 25.0, 4.0),
                                g['clean2_out'])
  np.testing.assert_array_equal(
      flow_utils.clean_flow(g['clean3d_in'].copy(), 1.4, 0.6, 25.0, 4.0, dim=3),
      g['clean3d_out'])
  assert 0.3 < np.isnan(g['clean_out']).mean() < 0.8


def test_reconcile_flows_golden(g):
  a, b, c = g['rec_a'], g['rec_b'], g['rec_c']
  np.testing.assert_array_equal(
      flow_utils.reconcile_flows([a.copy(), b.copy(), c.copy()], 5.0, 3.0, 12), g['rec_out'])
  np.testing.assert_array_equal(flow_utils.reconcile_flows([a.copy(), b.copy()], 0, 0, 0),
                                g['rec_out_nofilter'])
  np.testing.assert_array_equal(
      flow_utils.reconcile_flows([g['rec3_a'].copy(), g['rec3_b'].copy()], 5.0, 3.0, 12,
                                 min_delta_z=2), g['rec3_out'])


def test_decorator_chunk_functions(g):
  # decorators/flow.py:38-43, :349-354: chunk-shaped wrappers of the two filters
  # (the reference squeezes the chunk first, so for 2-d chunks with a singleton z
  # only the per-vector tests can run, max_deviation = 0: SciPy's 4-d median window
  # does not fit the squeezed array there either)
  chunk = g['clean_in'][:, :1]                      # [4, 1, y, x]
  got = dflow.clean_flow(chunk, min_peak_ratio=1.4, min_peak_sharpness=0.6,
                         max_magnitude=25.0, max_deviation=0)
  assert got.shape == (2, 1) + chunk.shape[2:]
  want = flow_utils.clean_flow(chunk, 1.4, 0.6, 25.0, 0)
  np.testing.assert_array_equal(got, want)
  got3 = dflow.clean_flow(g['clean3d_in'], min_peak_ratio=1.4, min_peak_sharpness=0.6,
                          max_magnitude=25.0, max_deviation=4.0)
  np.testing.assert_array_equal(got3, g['clean3d_out'])
  rec = dflow.reconcile_flow(g['rec_a'], max_gradient=5.0, max_deviation=3.0,
                             min_patch_size=12)
  np.testing.assert_array_equal(
      rec, flow_utils.reconcile_flows([g['rec_a']], 5.0, 3.0, 12))
