"""CPU twin of tests/test_pipeline_gpu.py: the config-4 chain on the oracle recovers the
known drift, and its threshold decisions are far from their boundaries (so that the fp32
rounding differences of the statistics channels between two correct implementations
cannot flip them)."""

import numpy as np

from tests import test_pipeline_gpu as tp


def test_oracle_chain_and_margins():
  stack = tp.make_stack()
  flow, clean, solved, warped = tp.oracle_chain(stack)
  assert flow.shape == (4, 3, 11, 11)
  # section 1 is section 0 shifted by (dy, dx) = (3, -2): flow (x, y) = (dx, dy), the sign
  # convention of tests/flow_field_test.py:28-37
  np.testing.assert_array_equal(flow[0, 0], -2.0)
  np.testing.assert_array_equal(flow[1, 0], 3.0)
  valid = ~np.isnan(flow[0])
  sharp, ratio = np.abs(flow[2][valid]), np.abs(flow[3][valid])
  assert np.all(np.abs(sharp - tp.CLEAN['min_peak_sharpness']) > 0.02 * sharp)
  nz = ratio > 0
  assert np.all(np.abs(ratio[nz] - tp.CLEAN['min_peak_ratio']) > 0.02)
  assert np.isnan(clean).any()                     # the blanked region was rejected
  assert np.isfinite(solved).all() and np.abs(solved).max() > 1.0
  assert warped.shape == stack.shape and warped.dtype == stack.dtype
  # the warp undoes the drift: section 1 warped onto section 0's frame matches it better
  a, b = stack[0, 80:400, 80:400].astype(float), stack[1, 80:400, 80:400].astype(float)
  w = warped[1, 80:400, 80:400].astype(float)
  assert np.abs(w - a).mean() < 0.7 * np.abs(b - a).mean()


def test_oracle_liconn_chain_and_margins(monkeypatch):
  flows, fine, agg, xr, t, warped = tp.oracle_stitch3d(monkeypatch)
  assert t == 300 and np.isfinite(xr).all() and np.abs(xr).max() > 0.5
  for fm in flows:
    for f in fm.values():
      valid = ~np.isnan(f[0])
      sharp, ratio = np.abs(f[3][valid]), np.abs(f[4][valid])
      assert np.all(np.abs(sharp - tp.CLEAN3['min_peak_sharpness']) > 0.02 * sharp)
      nz = ratio > 0
      assert np.all(np.abs(ratio[nz] - tp.CLEAN3['min_peak_ratio']) > 0.02)
  kept = sum(int((~np.isnan(f[0])).sum()) for fm in fine for f in fm.values())
  assert kept > 60
  for k, w in warped.items():
    assert w.shape == (24, 56, 64) and w.dtype == np.uint8
