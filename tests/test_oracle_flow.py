"""Pins oracle/flow_oracle.py: reference KATs (tests/flow_field_test.py) + golden
vectors generated from the reference's own flow_field.py."""

import numpy as np
import pytest

from oracle import flow_oracle as fo


def _check_flow(got, want, stats_rtol=2e-3):
  assert got.shape == want.shape
  nd = got.shape[0] - 2
  np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
  np.testing.assert_array_equal(got[:nd], want[:nd])  # integer offsets: exact
  ok = ~np.isnan(want[nd])
  np.testing.assert_allclose(got[nd][ok], want[nd][ok], rtol=stats_rtol)
  np.testing.assert_allclose(got[nd + 1][ok], want[nd + 1][ok], rtol=stats_rtol)


# ---- ports of /root/reference/tests/flow_field_test.py ---------------------------


def test_kat_delta_and_mask():  # flow_field_test.py:24-56
  pre = np.zeros((120, 120), np.uint8)
  post = np.zeros((120, 120), np.uint8)
  pre[60, 60] = 255
  post[70, 53] = 255
  calc = fo.MaskedXCorrWithStatsCalculator()
  field = calc.flow_field(pre, post, patch_size=80, step=40, batch_size=4)
  np.testing.assert_array_equal([4, 2, 2], field.shape)
  np.testing.assert_array_equal(7 * np.ones((2, 2)), field[0])
  np.testing.assert_array_equal(-10 * np.ones((2, 2)), field[1])
  np.testing.assert_array_equal(np.zeros((2, 2)), field[3])
  post[54, 68] = 255
  mask = np.zeros((128, 128), bool)
  mask[:55, :70] = 1
  field = calc.flow_field(pre, post, patch_size=80, step=40, post_mask=mask,
                          batch_size=4)
  np.testing.assert_array_equal(7 * np.ones((2, 2)), field[0])
  np.testing.assert_array_equal(-10 * np.ones((2, 2)), field[1])
  np.testing.assert_array_equal(np.zeros((2, 2)), field[3])


def test_kat_3d():  # flow_field_test.py:58-72
  pre = np.zeros((50, 100, 100), np.uint8)
  post = np.zeros((50, 100, 100), np.uint8)
  pre[25, 50, 50] = 255
  post[22, 45, 54] = 255
  flow = fo.MaskedXCorrWithStatsCalculator().flow_field(
      pre, post, patch_size=(40, 80, 80), step=10, batch_size=1)
  np.testing.assert_array_equal([5, 2, 3, 3], flow.shape)
  np.testing.assert_array_equal(np.full([2, 3, 3], -4), flow[0])
  np.testing.assert_array_equal(np.full([2, 3, 3], 5), flow[1])
  np.testing.assert_array_equal(np.full([2, 3, 3], 3), flow[2])


def test_kat_peak():  # flow_field_test.py:74-94
  hy, hx = np.mgrid[:50, :50]
  cy, cx = 20, 28
  r = np.sqrt(2 * (cx - hx) ** 2 + (cy - hy) ** 2)
  xcorr = 10 * np.exp(-r / 4)
  peaks = fo.batched_peaks(xcorr[np.newaxis], (25, 25), min_distance=2,
                           threshold_rel=0.5, peak_radius=(2, 3))
  assert peaks.shape == (1, 4)
  support = np.min(xcorr.astype(np.float32)[cy - 2:cy + 3, cx - 3:cx + 4])
  assert peaks[0, 0] == 3 and peaks[0, 1] == -5
  assert peaks[0, 2] == np.float32(10) / support
  assert peaks[0, 3] == 0


def test_kat_post_targeting():  # flow_field_test.py:96-125
  pre = np.zeros((120, 120), np.uint8)
  post = np.zeros((120, 120), np.uint8)
  pre[50, 55] = 255
  post[100, 100] = 255
  calc = fo.MaskedXCorrWithStatsCalculator()
  field = calc.flow_field(pre, post, patch_size=80, step=40, batch_size=4)
  assert np.all(np.isnan(field[:, 0, 0]))
  tgt = np.full((2, 2, 2), 40.0, np.float32)
  field = calc.flow_field(pre, post, patch_size=80, step=40, batch_size=4,
                          post_targeting_field=tgt, post_targeting_step=40)
  np.testing.assert_array_equal(-45 * np.ones((2, 2)), field[0])
  np.testing.assert_array_equal(-50 * np.ones((2, 2)), field[1])


# ---- golden vectors from the reference source ------------------------------------


def test_golden_textured(flow_golden):
  g = flow_golden
  calc = fo.MaskedXCorrWithStatsCalculator()
  pre, post = g['tex160_b4_pre'], g['tex160_b4_post']
  for bs in (4, 7, 64):
    _check_flow(calc.flow_field(pre, post, 160, 40, batch_size=bs),
                g[f'tex160_b{bs}_flow'])


def test_golden_periodic_second_peak(flow_golden):
  g = flow_golden
  got = fo.MaskedXCorrWithStatsCalculator().flow_field(
      g['periodic_pre'], g['periodic_post'], 120, 40, batch_size=5)
  assert (g['periodic_flow'][3] > 0).all()  # ratio channel is exercised
  _check_flow(got, g['periodic_flow'])


def test_golden_masked(flow_golden):
  g = flow_golden
  calc = fo.MaskedXCorrWithStatsCalculator()
  kw = dict(pre_mask=g['masked_pre_mask'], post_mask=g['masked_post_mask'],
            batch_size=6)
  _check_flow(calc.flow_field(g['masked_pre'], g['masked_post'], (96, 128),
                              (32, 40), **kw), g['masked_flow'])
  _check_flow(calc.flow_field(g['masked_pre'], g['masked_post'], (96, 128),
                              (32, 40), mask_only_for_patch_selection=True,
                              max_masked=0.4, **kw), g['masked_selonly_flow'])


def test_golden_postpatch_and_targeting(flow_golden):
  g = flow_golden
  calc = fo.MaskedXCorrWithStatsCalculator(mean=40.0, peak_radius=(3, 4))
  got = calc.flow_field(g['postpatch_pre'], g['postpatch_post'], 128, 24,
                        post_patch_size=96, selection_mask=g['postpatch_sel'],
                        batch_size=16)
  _check_flow(got, g['postpatch_flow'])
  got = fo.MaskedXCorrWithStatsCalculator().flow_field(
      g['masked_pre'], g['masked_post'], 128, 24,
      pre_targeting_field=g['pretarget_tg'], pre_targeting_step=24, batch_size=32)
  _check_flow(got, g['pretarget_flow'])


def test_golden_kats(flow_golden):
  g = flow_golden
  calc = fo.MaskedXCorrWithStatsCalculator()
  got = calc.flow_field(g['kat_delta_pre'], g['kat_delta_post'], 80, 40,
                        batch_size=4)
  np.testing.assert_array_equal(got[[0, 1, 3]], g['kat_delta_flow'][[0, 1, 3]])
  got = calc.flow_field(g['kat_notarget_pre'], g['kat_notarget_post'], 80, 40,
                        batch_size=4)
  np.testing.assert_array_equal(np.isnan(got), np.isnan(g['kat_notarget_flow']))
  got = fo.batched_peaks(g['peaks_bump_img'][np.newaxis], (25, 25), 2, 0.5, (2, 3))
  np.testing.assert_array_equal(got, g['peaks_bump'])


def test_helpers():
  m = np.random.default_rng(0).random((37, 41)) > 0.5
  s = fo.query_integral_image(fo.integral_image(m), (8, 10), (4, 5))
  want = np.array([[m[y:y + 8, x:x + 10].sum() for x in range(0, 41 - 10 + 1, 5)]
                   for y in range(0, 37 - 8 + 1, 4)])
  np.testing.assert_array_equal(s, want)
  assert [len(b) for b in fo.batch(list(range(10)), 4)] == [4, 4, 2]
  assert [fo.next_fast_len(n) for n in (319, 159, 255, 7, 1)] == [320, 160, 256, 8, 1]
  import scipy.fftpack  # the routine the reference calls (flow_field.py:67)
  assert all(fo.next_fast_len(n) == scipy.fftpack.next_fast_len(n) for n in range(1, 17000))


def test_reference_numpy_branch_golden():
  """Oracle vs the reference's OWN NumPy branch, flow_field.masked_xcorr(use_jax=False)
  (real numpy.fft -- no JAX stand-in involved; tests/golden/make_golden.py xcorr_numpy):
  unmasked, Padfield-masked, unequal patch sizes, 3-d."""
  import os
  g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'xcorr_numpy_golden.npz'))
  for tag in ('a', 'b'):
    prev, curr = g[f'xn_{tag}_prev'], g[f'xn_{tag}_curr']
    want = g[f'xn_{tag}_plain']
    got = fo.masked_xcorr(prev, curr)
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-6 * np.abs(want).max())
    want = g[f'xn_{tag}_masked']
    got = fo.masked_xcorr(prev, curr, g[f'xn_{tag}_pm'], g[f'xn_{tag}_cm'])
    assert got.shape == want.shape and np.abs(want).max() > 0.1
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-6)
  want = g['xn_3d_plain']
  got = fo.masked_xcorr(g['xn_3d_prev'], g['xn_3d_curr'], dim=3)
  assert got.shape == want.shape
  np.testing.assert_allclose(got, want, rtol=0, atol=2e-6 * np.abs(want).max())
  want = g['xn_3dm_masked']  # batch of two masked 3-d volumes (batch-global thresholds)
  got = fo.masked_xcorr(g['xn_3dm_prev'], g['xn_3dm_curr'], g['xn_3dm_pm'], g['xn_3dm_cm'], dim=3)
  assert got.shape == want.shape and np.abs(want).max() > 0.1
  np.testing.assert_allclose(got, want, rtol=0, atol=2e-6)
