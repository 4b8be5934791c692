"""Device flow filters (csrc/flowfilt.cu: clean_flow, reconcile_flows, mask_irregular) against
the golden vectors of the reference's own flow_utils.py and against the host filters, which
follow the reference line by line: results must be EQUAL, NaN patterns included."""

import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'flow_utils_golden.npz')


@pytest.fixture(scope='module')
def g():
  return np.load(GOLDEN)


@pytest.fixture(scope='module')
def dev():
  import torch
  if not torch.cuda.is_available():
    pytest.skip('needs a CUDA device')
  return lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_clean_flow_device_equals_golden(g, dev):
  from sofima_b200 import flow_utils
  f4 = g['clean_in']
  cases = [((f4, 1.4, 0.6, 25.0, 4.0), {}, 'clean_out'),
           ((f4, 1.4, 0.6, 0, 0), {}, 'clean_out_nodev'),
           ((f4[:2], 0, 0, 25.0, 4.0), {}, 'clean2_out'),
           ((g['clean3d_in'], 1.4, 0.6, 25.0, 4.0), {'dim': 3}, 'clean3d_out')]
  for (f, *args), kw, key in cases:
    got = flow_utils.clean_flow(dev(f), *args, **kw)
    assert got.is_cuda
    np.testing.assert_array_equal(got.cpu().numpy(), g[key])


def test_reconcile_flows_device_equals_golden(g, dev):
  from sofima_b200 import flow_utils
  a, b, c = g['rec_a'], g['rec_b'], g['rec_c']
  got = flow_utils.reconcile_flows([dev(a), dev(b), dev(c)], 5.0, 3.0, 12)
  np.testing.assert_array_equal(got.cpu().numpy(), g['rec_out'])
  got = flow_utils.reconcile_flows([dev(a), dev(b)], 0, 0, 0)
  np.testing.assert_array_equal(got.cpu().numpy(), g['rec_out_nofilter'])
  got = flow_utils.reconcile_flows([dev(g['rec3_a']), dev(g['rec3_b'])], 5.0, 3.0, 12,
                                   min_delta_z=2)
  np.testing.assert_array_equal(got.cpu().numpy(), g['rec3_out'])


def test_reference_kats_on_device(dev):  # tests/flow_utils_test.py:39-96
  from sofima_b200 import flow_utils
  flow = np.zeros((4, 1, 50, 40), np.float32)
  flow[2, ...] = 2.0
  flow[2, 0, 10, 20] = 1.2
  flow[3, 0, 10, 22] = 1.2
  flow[3, 0, 10, 24] = 1.6
  flow[0, 0, 5, 4] = 12
  flow[1, 0, 5, 6] = -14
  flow[:, 0, 5, 10] = 2
  flow[:, 0, 15, 10] = 7
  cleaned = flow_utils.clean_flow(dev(flow), min_peak_ratio=1.4, min_peak_sharpness=1.6,
                                  max_magnitude=10, max_deviation=5).cpu().numpy()
  expected = np.zeros((2, 1, 50, 40))
  expected[:, 0, 5, 10] = 2
  for y, x in ((15, 10), (10, 20), (10, 22), (5, 4), (5, 6)):
    expected[:, 0, y, x] = np.nan
  np.testing.assert_array_equal(cleaned, expected)
  f1 = np.full((3, 1, 50, 40), np.nan, np.float32)
  f2, f3 = f1.copy(), f1.copy()
  f1[:, 0, 10, 10] = 2.
  f2[:, 0, 10, 10] = 3.
  f3[:, 0, 20, 20] = 4.
  f2[:, 0, 20, 20] = 1.
  f2[:, 0, 30:35, 30:35] = 5
  f2[0, 0, 32, 32] = 15
  got = flow_utils.reconcile_flows([dev(f1), dev(f2), dev(f3)], max_gradient=0, max_deviation=8,
                                   min_patch_size=0, min_delta_z=2).cpu().numpy()
  expected = np.full((3, 1, 50, 40), np.nan)
  expected[:, 0, 10, 10] = 2.
  expected[:, 0, 20, 20] = 4.
  expected[:, 0, 30:35, 30:35] = 5
  expected[:, 0, 32, 32] = np.nan
  np.testing.assert_array_equal(got, expected)


@pytest.mark.parametrize('seed', [0, 1, 2])
def test_random_fields_device_equals_host(dev, seed):
  """Larger random fields with NaN holes, ragged valid regions (connected components) and
  values sitting exactly on thresholds."""
  import scipy.ndimage as ndi
  from sofima_b200 import flow_utils
  rng = np.random.default_rng(seed)
  shape = (4, 3, 97, 113)
  f = np.zeros(shape, np.float32)
  f[:2] = ndi.gaussian_filter(rng.standard_normal((2,) + shape[1:]), (0, 0, 3, 3)) * 30
  f[:2] += (rng.random((2,) + shape[1:]) < 0.05) * rng.standard_normal((2,) + shape[1:]) * 40
  f[2] = rng.random(shape[1:]) * 3
  f[3] = np.where(rng.random(shape[1:]) < 0.3, 0.0, rng.random(shape[1:]) * 3)
  f[:, rng.random(shape[1:]) < 0.2] = np.nan
  f[0, 0, 5, 5], f[1, 0, 5, 5] = 25.0, -25.0          # exactly max_magnitude: kept
  f[2, 0, 7, 7] = np.float32(0.6)                     # exactly min_peak_sharpness: kept
  want = flow_utils.clean_flow(f.copy(), 1.4, 0.6, 25.0, 4.0)
  got = flow_utils.clean_flow(dev(f), 1.4, 0.6, 25.0, 4.0).cpu().numpy()
  np.testing.assert_array_equal(got, want)
  other = flow_utils.clean_flow((f * np.float32(0.9)).copy(), 1.1, 0.2, 40.0, 8.0)
  for args in ((5.0, 3.0, 12), (2.0, 0, 30), (0, 1.5, 5), (0, 0, 200)):
    w = flow_utils.reconcile_flows([want.copy(), other.copy()], *args)
    d = flow_utils.reconcile_flows([dev(want), dev(other)], *args).cpu().numpy()
    np.testing.assert_array_equal(d, w)
  assert 0.05 < np.isnan(w).mean() < 0.999


@pytest.mark.parametrize('iters', [0, 1, 3])
def test_mask_irregular_device_equals_host(dev, iters):
  import scipy.ndimage as ndi
  from sofima_b200.processor import mesh as pmesh
  rng = np.random.default_rng(5 + iters)
  m = (ndi.gaussian_filter(rng.standard_normal((2, 90, 77)), (0, 1.5, 1.5)) * 90).astype(np.float32)
  m[:, rng.random((90, 77)) < 0.02] = np.nan
  for stride, frac, max_frac in (((40, 40), 0.5, 2.0), ((20.0, 30.0), 0.7, None)):
    host = m.copy()
    bad_h = pmesh.mask_irregular(host, stride, frac, max_frac, dilation_iters=iters)
    d = dev(m)
    bad_d = pmesh.mask_irregular(d, stride, frac, max_frac, dilation_iters=iters)
    np.testing.assert_array_equal(bad_d.cpu().numpy(), bad_h)
    np.testing.assert_array_equal(d.cpu().numpy(), host)
    assert bad_h.any()
