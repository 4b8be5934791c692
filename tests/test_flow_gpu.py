"""GPU parity tests: sofima_b200.flow_field (CUDA, through the C ABI) vs the oracle
and the golden vectors generated from the reference source.

Tolerances (north_star): flow offsets 1e-4 rel -- they are integers, so the test
demands equality; NaN patterns must match exactly.  The two statistics channels
(sharpness, peak ratio) are fp32 quotients of FFT outputs; rtol 2e-3 (SURVEY 8d
asks 1e-3 as a goal, non-gating).
"""

import numpy as np
import pytest
import scipy.ndimage as ndi

from oracle import flow_oracle as fo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ff():
  import torch
  if not torch.cuda.is_available():
    pytest.skip('needs a CUDA device')
  from sofima_b200 import flow_field
  return flow_field


def _check_flow(got, want, stats_rtol=2e-3, stats=True):
  assert got.shape == want.shape
  nd = got.shape[0] - 2
  np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
  np.testing.assert_array_equal(got[:nd], want[:nd])
  if stats:
    ok = ~np.isnan(want[nd])
    np.testing.assert_allclose(got[nd][ok], want[nd][ok], rtol=stats_rtol)
    np.testing.assert_allclose(got[nd + 1][ok], want[nd + 1][ok], rtol=stats_rtol,
                               atol=1e-6)


# ---- ports of /root/reference/tests/flow_field_test.py on the CUDA path ----------


def test_kat_delta_and_mask(ff):
  pre = np.zeros((120, 120), np.uint8)
  post = np.zeros((120, 120), np.uint8)
  pre[60, 60] = 255
  post[70, 53] = 255
  calc = ff.JAXMaskedXCorrWithStatsCalculator()
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


def test_kat_peak(ff):
  hy, hx = np.mgrid[:50, :50]
  cy, cx = 20, 28
  r = np.sqrt(2 * (cx - hx) ** 2 + (cy - hy) ** 2)
  xcorr = 10 * np.exp(-r / 4)
  peaks = ff._batched_peaks(xcorr[np.newaxis], (25, 25), min_distance=2,
                            threshold_rel=0.5, peak_radius=(2, 3))
  assert peaks.shape == (1, 4)
  support = np.min(xcorr.astype(np.float32)[cy - 2:cy + 3, cx - 3:cx + 4])
  assert peaks[0, 0] == 3 and peaks[0, 1] == -5
  assert peaks[0, 2] == np.float32(10) / support
  assert peaks[0, 3] == 0


def test_kat_post_targeting(ff):
  pre = np.zeros((120, 120), np.uint8)
  post = np.zeros((120, 120), np.uint8)
  pre[50, 55] = 255
  post[100, 100] = 255
  calc = ff.JAXMaskedXCorrWithStatsCalculator()
  field = calc.flow_field(pre, post, patch_size=80, step=40, batch_size=4)
  assert np.all(np.isnan(field[:, 0, 0]))
  tgt = np.full((2, 2, 2), 40.0, np.float32)
  field = calc.flow_field(pre, post, patch_size=80, step=40, batch_size=4,
                          post_targeting_field=tgt, post_targeting_step=40)
  np.testing.assert_array_equal(-45 * np.ones((2, 2)), field[0])
  np.testing.assert_array_equal(-50 * np.ones((2, 2)), field[1])


def test_kat_3d(ff, flow_golden):  # flow_field_test.py:58-72
  pre = np.zeros((50, 100, 100), np.uint8)
  post = np.zeros((50, 100, 100), np.uint8)
  pre[25, 50, 50] = 255
  post[22, 45, 54] = 255
  flow = ff.JAXMaskedXCorrWithStatsCalculator().flow_field(
      pre, post, patch_size=(40, 80, 80), step=10, batch_size=1)
  np.testing.assert_array_equal([5, 2, 3, 3], flow.shape)
  np.testing.assert_array_equal(np.full([2, 3, 3], -4), flow[0])
  np.testing.assert_array_equal(np.full([2, 3, 3], 5), flow[1])
  np.testing.assert_array_equal(np.full([2, 3, 3], 3), flow[2])
  want = flow_golden['kat_3d_flow']
  np.testing.assert_array_equal(flow[:3], want[:3])
  np.testing.assert_array_equal(flow[4], want[4])


def test_3d_textures_match_oracle(ff):
  rng = np.random.default_rng(4)
  base = ndi.gaussian_filter(rng.standard_normal((40, 90, 100)), 1.2)
  base = ((base - base.min()) / (base.max() - base.min()) * 255).astype(np.uint8)
  pre = np.ascontiguousarray(base[4:34, 8:78, 10:90])
  post = np.ascontiguousarray(base[5:35, 6:76, 13:93])
  kw = dict(patch_size=(16, 32, 40), step=(7, 19, 20), batch_size=5)
  got = ff.JAXMaskedXCorrWithStatsCalculator(peak_radius=(2, 3, 4)).flow_field(pre, post, **kw)
  want = fo.MaskedXCorrWithStatsCalculator(peak_radius=(2, 3, 4)).flow_field(pre, post, **kw)
  assert got.shape == want.shape and got.shape[0] == 5
  _check_flow(got, want)
  # _batched_peaks in 3-d on oracle correlation volumes
  center, xc = fo.batched_xcorr(pre, post, None, None, (16, 32, 40),
                                np.array([[0, 0, 0], [7, 19, 20], [14, 38, 40]]), None)
  np.testing.assert_allclose(ff._batched_peaks(xc, center, 2, 0.5, (2, 3, 4)),
                             fo.batched_peaks(xc, center, 2, 0.5, (2, 3, 4)), rtol=1e-5)
  # masked 3-d patches (csrc/flow3d_masked.cuh): an all-valid mask goes through the
  # Padfield path and must still agree with the oracle's masked path
  m0 = np.zeros(pre.shape, bool)
  got_m = ff.JAXMaskedXCorrWithStatsCalculator(peak_radius=(2, 3, 4)).flow_field(
      pre, post, pre_mask=m0, **kw)
  want_m = fo.MaskedXCorrWithStatsCalculator(peak_radius=(2, 3, 4)).flow_field(
      pre, post, pre_mask=m0, **kw)
  _check_flow(got_m, want_m)


# ---- golden vectors from the reference source -------------------------------------


def test_golden_textured(ff, flow_golden):
  g = flow_golden
  calc = ff.JAXMaskedXCorrWithStatsCalculator()
  pre, post = g['tex160_b4_pre'], g['tex160_b4_post']
  for bs in (4, 7, 64):
    _check_flow(calc.flow_field(pre, post, 160, 40, batch_size=bs),
                g[f'tex160_b{bs}_flow'])


def test_golden_periodic_second_peak(ff, flow_golden):
  g = flow_golden
  got = ff.JAXMaskedXCorrWithStatsCalculator().flow_field(
      g['periodic_pre'], g['periodic_post'], 120, 40, batch_size=5)
  _check_flow(got, g['periodic_flow'])


def test_golden_masked(ff, flow_golden):
  g = flow_golden
  calc = ff.JAXMaskedXCorrWithStatsCalculator()
  kw = dict(pre_mask=g['masked_pre_mask'], post_mask=g['masked_post_mask'],
            batch_size=6)
  _check_flow(calc.flow_field(g['masked_pre'], g['masked_post'], (96, 128),
                              (32, 40), **kw), g['masked_flow'])
  _check_flow(calc.flow_field(g['masked_pre'], g['masked_post'], (96, 128),
                              (32, 40), mask_only_for_patch_selection=True,
                              max_masked=0.4, **kw), g['masked_selonly_flow'])


def test_golden_postpatch_and_targeting(ff, flow_golden):
  g = flow_golden
  calc = ff.JAXMaskedXCorrWithStatsCalculator(mean=40.0, peak_radius=(3, 4))
  got = calc.flow_field(g['postpatch_pre'], g['postpatch_post'], 128, 24,
                        post_patch_size=96, selection_mask=g['postpatch_sel'],
                        batch_size=16)
  _check_flow(got, g['postpatch_flow'])
  got = ff.JAXMaskedXCorrWithStatsCalculator().flow_field(
      g['masked_pre'], g['masked_post'], 128, 24,
      pre_targeting_field=g['pretarget_tg'], pre_targeting_step=24, batch_size=32)
  _check_flow(got, g['pretarget_flow'])


def test_golden_kats(ff, flow_golden):
  g = flow_golden
  calc = ff.JAXMaskedXCorrWithStatsCalculator()
  got = calc.flow_field(g['kat_delta_pre'], g['kat_delta_post'], 80, 40,
                        batch_size=4)
  np.testing.assert_array_equal(got[[0, 1, 3]], g['kat_delta_flow'][[0, 1, 3]])
  got = calc.flow_field(g['kat_notarget_pre'], g['kat_notarget_post'], 80, 40,
                        batch_size=4)
  np.testing.assert_array_equal(np.isnan(got), np.isnan(g['kat_notarget_flow']))
  got = ff._batched_peaks(g['peaks_bump_img'][np.newaxis], (25, 25), 2, 0.5, (2, 3))
  np.testing.assert_array_equal(got, g['peaks_bump'])


# ---- seeded oracle comparisons ------------------------------------------------------


def _texture(seed, shape, sigma=2.0):
  rng = np.random.default_rng(seed)
  base = ndi.gaussian_filter(rng.standard_normal(shape), sigma)
  return ((base - base.min()) / (base.max() - base.min()) * 255).astype(np.uint8)


def test_xcorr_images_match_oracle(ff):
  """The raw correlation images (what _batched_xcorr returns) vs pocketfft fp32."""
  rng = np.random.default_rng(3)
  for (ph, pw), (qh, qw) in (((160, 160), (160, 160)), ((50, 37), (31, 64)),
                             ((9, 9), (9, 9))):
    a = rng.standard_normal((3, ph, pw)).astype(np.float32)
    b = rng.standard_normal((3, qh, qw)).astype(np.float32)
    got = ff.masked_xcorr(a, b)
    want = fo.masked_xcorr(a, b)
    assert got.shape == want.shape
    scale = np.abs(want).max()
    np.testing.assert_allclose(got / scale, want / scale, atol=2e-6)
    am = rng.random(a.shape) > 0.8
    bm = rng.random(b.shape) > 0.7
    got = ff.masked_xcorr(a, b, am, bm)
    want = fo.masked_xcorr(a, b, am, bm)
    np.testing.assert_allclose(got, want, atol=2e-4)


def test_config1_512_tile_pair(ff):
  """BASELINE config 1: 512x512 tile pair, patch 160, step 40 -> 81 patch pairs."""
  rng = np.random.default_rng(0)
  base = ndi.gaussian_filter(rng.standard_normal((640, 640)), 2.0)
  base = ((base - base.min()) / (base.max() - base.min()) * 255).astype(np.uint8)
  pre = base[64:576, 64:576]
  post = np.clip(base[69:581, 61:573].astype(float) + rng.normal(0, 5, (512, 512)),
                 0, 255).astype(np.uint8)
  got = ff.JAXMaskedXCorrWithStatsCalculator().flow_field(pre, post, 160, 40,
                                                          batch_size=1024)
  assert got.shape == (4, 9, 9)
  np.testing.assert_array_equal(got[0], -3)
  np.testing.assert_array_equal(got[1], 5)
  want = fo.MaskedXCorrWithStatsCalculator().flow_field(pre, post, 160, 40,
                                                        batch_size=1024)
  _check_flow(got, want)


@pytest.mark.parametrize('seed,patch,step,bs', [
    (1, 64, 16, 32), (2, (48, 80), (16, 24), 9), (3, 33, 11, 128)])
def test_random_textures_match_oracle(ff, seed, patch, step, bs):
  img = _texture(seed, (260, 300), sigma=1.5)
  rng = np.random.default_rng(seed)
  pre = img[10:210, 10:250]
  post = np.clip(img[13:213, 6:246].astype(float) + rng.normal(0, 12, (200, 240)),
                 0, 255).astype(np.uint8)
  got = ff.JAXMaskedXCorrWithStatsCalculator().flow_field(pre, post, patch, step,
                                                          batch_size=bs)
  want = fo.MaskedXCorrWithStatsCalculator().flow_field(pre, post, patch, step,
                                                        batch_size=bs)
  _check_flow(got, want)


def test_empty_selection(ff):
  pre = np.zeros((100, 100), np.uint8)
  sel = np.zeros((4, 4), bool)
  got = ff.JAXMaskedXCorrWithStatsCalculator().flow_field(
      pre, pre, 40, 20, selection_mask=sel, batch_size=8)
  assert got.shape == (4, 4, 4) and np.isnan(got).all()


def test_full_size_properties(ff):
  """4096^2 tile pair (BASELINE config 2 tile size): 9801 patch pairs.

  Size-independent properties: a pure translation is recovered everywhere, the
  result is invariant to how the grid is cut into images (a crop gives the same
  rows), and repeated runs are bit-identical.
  """
  n = 4096
  base = _texture(9, (n + 64, n + 64), sigma=2.0)
  pre = base[32:32 + n, 32:32 + n]
  post = base[32 + 4:32 + 4 + n, 32 - 7:32 - 7 + n]
  calc = ff.JAXMaskedXCorrWithStatsCalculator()
  got = calc.flow_field(pre, post, 160, 40, batch_size=1024)
  assert got.shape == (4, 99, 99)
  np.testing.assert_array_equal(got[0], -7)
  np.testing.assert_array_equal(got[1], 4)
  again = calc.flow_field(pre, post, 160, 40, batch_size=1024)
  np.testing.assert_array_equal(got, again)
  # A 1024 crop, same batch composition rule -> same offsets on the shared grid.
  crop = calc.flow_field(pre[:1024, :1024], post[:1024, :1024], 160, 40,
                         batch_size=1024)
  np.testing.assert_array_equal(crop[:2], got[:2, :crop.shape[1], :crop.shape[2]])
  np.testing.assert_allclose(crop[2], got[2, :crop.shape[1], :crop.shape[2]],
                             rtol=1e-3)


@pytest.mark.parametrize('dtype,post_patch,mean', [(np.uint8, None, None),
                                                   (np.float32, 48, None),
                                                   (np.uint8, None, 101.5)])
def test_shared_row_spectra_match_plain_path(ff, monkeypatch, dtype, post_patch, mean):
  """The shared-row-transform form (sofima_xcorr_rowcache: one forward row transform per
  image row and patch x start, mean and post flip applied by linearity) against the
  per-patch transforms and against the oracle."""
  from sofima_b200 import _native
  rng = np.random.default_rng(31)
  base = ndi.gaussian_filter(rng.standard_normal((460, 520)), 1.5)
  base = (base - base.min()) / (base.max() - base.min()) * 255
  pre = np.ascontiguousarray(base[20:420, 30:500]).astype(dtype)
  post = np.ascontiguousarray(base[24:424, 27:497]).astype(dtype)
  if dtype == np.float32:
    pre, post = pre / np.float32(3), post / np.float32(3)
  calc = ff.JAXMaskedXCorrWithStatsCalculator(mean=mean)
  kw = dict(patch_size=80 if post_patch else 64, step=16, batch_size=128,
            post_patch_size=post_patch)  # 80 + 48 - 1 -> 128, 64 + 64 - 1 -> 128 = 16 x 8
  ctx = _native.Context.get(0)
  ctx.set_timing(True)
  shared = calc.flow_field(pre, post, **kw)
  rep = ctx.timing_report()
  assert 'flow_rowspec' in rep and 'flow_rows_fwd' not in rep   # the shared form ran
  monkeypatch.setenv('SOFIMA_FLOW_ROWCACHE', '0')
  plain = calc.flow_field(pre, post, **kw)
  rep = ctx.timing_report()
  ctx.set_timing(False)
  assert 'flow_rows_fwd' in rep and 'flow_rowspec' not in rep
  _check_flow(shared, plain, stats_rtol=2e-4)
  want = fo.MaskedXCorrWithStatsCalculator(mean=mean).flow_field(pre, post, **kw)
  _check_flow(shared, want)
  assert np.isfinite(shared[0]).mean() > 0.9


def test_index_tables_are_reused_between_calls(ff):
  """Same geometry -> the cached index tables; another selection mask -> new tables; the
  results do not depend on which one was used."""
  rng = np.random.default_rng(11)
  base = ndi.gaussian_filter(rng.standard_normal((400, 400)), 2.0)
  img = ((base - base.min()) / np.ptp(base) * 255).astype(np.uint8)
  pre, post = img[20:340, 20:340], img[23:343, 18:338]
  calc = ff.JAXMaskedXCorrWithStatsCalculator()
  ff._JOB_CACHE.clear()
  kw = dict(patch_size=160, step=40, batch_size=8)
  a = calc.flow_field(pre, post, **kw)
  assert len(ff._JOB_CACHE) == 1
  job = next(iter(ff._JOB_CACHE.values()))
  b = calc.flow_field(pre, post, **kw)
  assert len(ff._JOB_CACHE) == 1 and next(iter(ff._JOB_CACHE.values())) is job
  np.testing.assert_array_equal(a, b)
  sel = np.ones(a.shape[1:], bool)
  sel[1, 2] = False
  c = calc.flow_field(pre, post, selection_mask=sel, **kw)
  assert len(ff._JOB_CACHE) == 2
  assert np.all(np.isnan(c[:, 1, 2]))
  np.testing.assert_array_equal(c[:2][:, sel], a[:2][:, sel])
  want = fo.MaskedXCorrWithStatsCalculator().flow_field(pre, post, **kw)
  _check_flow(a, want)


def test_masked_xcorr_3d_unequal_volumes(ff):
  """flow_field.masked_xcorr(dim=3) + _batched_peaks as the coarse 3-d offset search of
  notebooks/liconn_inplane_stitching.ipynb (cell 12) uses them: a small query cuboid
  correlated with a larger search volume, no masks, peak relative to a caller-supplied
  centre."""
  rng = np.random.default_rng(21)
  vol = ndi.gaussian_filter(rng.standard_normal((40, 44, 48)), 1.2).astype(np.float32)
  search = vol[4:34, 6:38, 2:42]                      # 30 x 32 x 40
  query = vol[4 + 9:4 + 9 + 8, 6 + 11:6 + 11 + 10, 2 + 13:2 + 13 + 12].copy()   # at (9, 11, 13)
  search = search - search.mean()
  query = query - query.mean()
  got = ff.masked_xcorr(search, query, use_jax=True, dim=3)
  want = fo.masked_xcorr(search, query, dim=3)
  assert got.shape == want.shape == (37, 41, 51)
  scale = np.abs(want).max()
  np.testing.assert_allclose(got / scale, want / scale, atol=3e-6)
  center = (got.shape[0] // 2, got.shape[1] // 2, got.shape[2] // 2)
  pk = ff._batched_peaks(got[None], center, 2, 0.5)
  pk_w = fo.batched_peaks(want[None], center, 2, 0.5)
  np.testing.assert_array_equal(pk[0, :3], pk_w[0, :3])
  np.testing.assert_allclose(pk[0, 3:], pk_w[0, 3:], rtol=2e-3, atol=1e-6)
  # batch axis + the error contract
  b2 = ff.masked_xcorr(np.stack([search, search]), np.stack([query, -query]), dim=3)
  np.testing.assert_allclose(b2[0] / scale, want / scale, atol=3e-6)
  np.testing.assert_allclose(b2[1] / scale, -want / scale, atol=3e-6)
  sm = np.zeros(search.shape, bool)
  sm[:5, :7, :9] = True
  got_m = ff.masked_xcorr(search, query, sm, None, dim=3)
  want_m = fo.masked_xcorr(search, query, sm, None, dim=3)
  np.testing.assert_allclose(got_m, want_m, rtol=0, atol=2e-4)


# ---- Blackwell-specific paths: tensor-core row spectra, fused per-pair kernel ---------


def _rowcache_case(seed=41, n=520):
  rng = np.random.default_rng(seed)
  base = ndi.gaussian_filter(rng.standard_normal((n + 40, n + 40)), 1.5)
  base = ((base - base.min()) / (base.max() - base.min()) * 255).astype(np.uint8)
  pre = np.ascontiguousarray(base[20:20 + n - 8, 20:20 + n])     # width 520: 16-byte rows
  post = np.ascontiguousarray(base[23:23 + n - 8, 16:16 + n])
  return pre, post


@pytest.mark.parametrize('patch,step', [(160, 40), (64, 24)])
def test_tensor_core_row_spectra_match_fft_row_spectra(ff, monkeypatch, patch, step):
  """rowspec_tc_kernel (TMA -> tcgen05.mma kind::i8 with base-128 twiddle digits, exact
  integer DFT; x starts at 24 j exercise three different 16-byte misalignments) against
  rowspec_fast (fp32 FFT on the CUDA cores) and against the oracle."""
  pre, post = _rowcache_case()
  calc = ff.JAXMaskedXCorrWithStatsCalculator()
  kw = dict(patch_size=patch, step=step, batch_size=64)
  monkeypatch.setenv('SOFIMA_FLOW_ROWSPEC_TC', '1')
  tc = calc.flow_field(pre, post, **kw)
  monkeypatch.setenv('SOFIMA_FLOW_ROWSPEC_TC', '0')
  fft = calc.flow_field(pre, post, **kw)
  np.testing.assert_array_equal(np.isnan(tc), np.isnan(fft))
  np.testing.assert_array_equal(tc[:2], fft[:2])              # integer flow vectors
  ok = np.isfinite(fft[2])
  # the two row transforms differ by fp32 rounding only: compare min / peak absolutely
  np.testing.assert_allclose(1 / tc[2][ok], 1 / fft[2][ok], rtol=0, atol=2e-6)
  np.testing.assert_allclose(tc[3][ok], fft[3][ok], rtol=2e-4, atol=1e-6)
  want = fo.MaskedXCorrWithStatsCalculator().flow_field(pre, post, **kw)
  np.testing.assert_array_equal(tc[:2], want[:2])
  assert np.isfinite(tc[0]).mean() > 0.9


@pytest.mark.parametrize('groups', ['3', '4'])
@pytest.mark.parametrize('patch,step,bs', [(160, 40, 16), (64, 16, 100)])
def test_fused_pipeline_matches_three_kernel_path(ff, monkeypatch, groups, patch, step, bs):
  """pair_fused (columns -> product -> inverse columns -> inverse rows -> peaks in one
  persistent launch, SOFIMA_FLOW_FUSED=1) reproduces the three-kernel path BIT FOR BIT,
  statistics channels included -- it executes the same codelets in the same order."""
  pre, post = _rowcache_case(seed=43)
  calc = ff.JAXMaskedXCorrWithStatsCalculator()
  kw = dict(patch_size=patch, step=step, batch_size=bs)
  monkeypatch.setenv('SOFIMA_FLOW_FUSED', '0')
  three = calc.flow_field(pre, post, **kw)
  monkeypatch.setenv('SOFIMA_FLOW_FUSED', '1')
  monkeypatch.setenv('SOFIMA_FLOW_FUSED_GROUPS', groups)
  from sofima_b200 import _native
  ctx = _native.Context.get(0)
  ctx.set_timing(True)
  fused = calc.flow_field(pre, post, **kw)
  rep = ctx.timing_report()
  ctx.set_timing(False)
  assert 'flow_fused' in rep and 'flow_cols' not in rep        # the fused kernel did run
  np.testing.assert_array_equal(fused, three)
  assert np.isfinite(fused[0]).mean() > 0.9


def test_fused_pipeline_second_peak_rule(ff, monkeypatch):
  """A periodic texture gives many local maxima above half the peak height; the fused
  kernel's 32-candidate record and the exact fix-up pass must reproduce the batch-coupled
  erase rule (flow_field.py:263-265) of the three-kernel path."""
  y, x = np.mgrid[:400, :416]
  rng = np.random.default_rng(7)
  tex = (np.sin(2 * np.pi * x / 20.0) + np.sin(2 * np.pi * y / 24.0)) * 50 + 128
  pre = np.clip(tex + rng.normal(0, 4, tex.shape), 0, 255).astype(np.uint8)
  post = np.clip(np.roll(tex, (2, -3), (0, 1)) + rng.normal(0, 4, tex.shape), 0, 255).astype(np.uint8)
  calc = ff.JAXMaskedXCorrWithStatsCalculator()
  kw = dict(patch_size=160, step=40, batch_size=24)
  monkeypatch.setenv('SOFIMA_FLOW_FUSED', '0')
  three = calc.flow_field(pre, post, **kw)
  monkeypatch.setenv('SOFIMA_FLOW_FUSED', '1')
  fused = calc.flow_field(pre, post, **kw)
  np.testing.assert_array_equal(fused, three)
  assert (three[3][np.isfinite(three[3])] > 0).any()            # second peaks do exist


def test_3d_register_codelet_axis_passes(ff, monkeypatch):
  """3-d patches whose padded lengths are 16 * N2 (64 -> 128, 80 -> 160, 96 -> 192: N2 = 8, 10,
  12 in one call) go through axis_fft_fast_kernel with pruned forward passes; the generic
  Stockham passes (SOFIMA_FLOW3D_FAST=0) and the oracle must give the same integer flow."""
  rng = np.random.default_rng(11)
  base = ndi.gaussian_filter(rng.standard_normal((110, 130, 150)), 1.5)
  base = ((base - base.min()) / (base.max() - base.min()) * 255).astype(np.uint8)
  pre = np.ascontiguousarray(base[4:100, 6:118, 8:136])
  post = np.ascontiguousarray(base[6:102, 3:115, 12:140])
  kw = dict(patch_size=(64, 80, 96), step=(32, 32, 32), batch_size=4)
  calc = ff.JAXMaskedXCorrWithStatsCalculator(peak_radius=(3, 3, 3))
  fast = calc.flow_field(pre, post, **kw)
  monkeypatch.setenv('SOFIMA_FLOW3D_FAST', '0')
  generic = calc.flow_field(pre, post, **kw)
  monkeypatch.delenv('SOFIMA_FLOW3D_FAST')
  assert fast.shape == generic.shape and fast.shape[0] == 5
  np.testing.assert_array_equal(fast[:3], generic[:3])
  np.testing.assert_allclose(fast[3:], generic[3:], rtol=2e-3, atol=1e-5)
  assert np.isfinite(fast[:3]).all()
  # pre = base[4:, 6:, 8:], post = base[6:, 3:, 12:]: one shift for every patch (x, y, z)
  assert (fast[0] == 4).all() and (fast[1] == -3).all() and (fast[2] == 2).all()
  want = fo.MaskedXCorrWithStatsCalculator(peak_radius=(3, 3, 3)).flow_field(pre, post, **kw)
  _check_flow(fast, want)
