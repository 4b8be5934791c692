"""Pins oracle/warp_cv_oracle.py (CPU): against the committed outputs of the reference's
warp.warp_subvolume (tests/golden/warp_cv_golden.npz) and, where cv2 / scipy are
importable, operation by operation against the real third-party routines."""

import os

import numpy as np
import pytest
import scipy.ndimage as ndi
from scipy import interpolate

from oracle import warp_cv_oracle as wo
from sofima_b200 import compat

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'warp_cv_golden.npz')


@pytest.fixture(scope='module')
def g():
  return np.load(GOLDEN)


def boxes(g):
  b = g['a_boxes']
  return tuple(compat.BoundingBox(start=b[i], size=b[i + 1]) for i in (0, 2, 4))


@pytest.mark.parametrize('dt', ['u8', 'u16', 'f32'])
@pytest.mark.parametrize('inter', ['nearest', 'linear', 'cubic', 'lanczos'])
def test_oracle_reproduces_reference_outputs(g, dt, inter):
  ib, mb, ob = boxes(g)
  got = wo.warp_subvolume(g[f'a_image_{dt}'], ib, g['a_map'], mb, 8, ob, interpolation=inter)
  want = g[f'a_{dt}_{inter}']
  assert got.dtype == want.dtype
  np.testing.assert_array_equal(got, want)
  assert not want[:, 1].any()  # the all-NaN section is skipped
  assert want[:, 0].any() and want[:, 2].any()


def test_oracle_reference_variants(g):
  ib, mb, ob = boxes(g)
  np.testing.assert_array_equal(
      wo.warp_subvolume(g['a_image_u8'], ib, g['a_map'].astype(np.float32), mb, 8, ob),
      g['a_u8_default_f32map'])
  np.testing.assert_array_equal(
      wo.warp_subvolume(g['a_image_u8'], ib, g['a_map'], mb, 8.0, ob, interpolation='linear',
                        offset=0.5), g['a_u8_offset'])
  got = wo.warp_subvolume(g['a_image_u16'].astype(np.uint32), ib, g['a_map'], mb, 8, ob,
                          interpolation='linear')
  assert got.dtype == np.uint32
  np.testing.assert_array_equal(got, g['a_u32_linear'])
  got = wo.warp_subvolume(g['b_seg'], ib, g['a_map'], mb, 8, ob)
  assert got.dtype == np.uint64 and got.max() > 2**40
  np.testing.assert_array_equal(got, g['b_seg_warped'])
  with pytest.raises(ValueError):
    wo.warp_subvolume(np.full((1, 3, 4, 4), 2**16, np.uint32), ib, g['a_map'], mb, 8, ob)


@pytest.mark.parametrize('dtype', [np.float64, np.float32])
def test_grid_interpolation_equals_scipy(dtype):
  rng = np.random.default_rng(3)
  vals = (rng.standard_normal((13, 17)) * 50).astype(dtype)
  gy = (np.arange(13) + 3) * 7.5 - 11
  gx = (np.arange(17) - 2) * 7.5 - 4
  qy, qx = np.mgrid[:120, :130]
  want = interpolate.RegularGridInterpolator((gy, gx), vals, bounds_error=False,
                                             fill_value=None)((qy, qx))
  np.testing.assert_array_equal(wo.rgi_linear_2d(gy, gx, vals, qy, qx), want)


def _maps(rng, oh, ow, h, w):
  yy, xx = np.mgrid[:oh, :ow]
  smooth = ((xx * 1.07 - 6 + ndi.gaussian_filter(rng.standard_normal((oh, ow)), 8) * 30),
            (yy * 1.15 - 7 + ndi.gaussian_filter(rng.standard_normal((oh, ow)), 8) * 30))
  bad = [m.copy() for m in smooth]
  bad[0][10:20, 30:50] = np.nan
  bad[1][50, 5] = np.inf
  bad[0][60, 6], bad[0][61, 6] = 1e12, -1e12
  return {'random': (rng.random((oh, ow)) * (w + 20) - 10, rng.random((oh, ow)) * (h + 20) - 10),
          'smooth': smooth, 'nonfinite': bad}


@pytest.mark.parametrize('kind', ['random', 'smooth', 'nonfinite'])
def test_remap_equals_opencv(kind):
  cv = pytest.importorskip('cv2')
  flags = {'nearest': cv.INTER_NEAREST, 'linear': cv.INTER_LINEAR, 'cubic': cv.INTER_CUBIC,
           'lanczos': cv.INTER_LANCZOS4}
  rng = np.random.default_rng(3)
  h, w, oh, ow = 120, 140, 100, 130
  imgs = [rng.integers(0, 256, (h, w), dtype=np.uint8),
          rng.integers(0, 65536, (h, w), dtype=np.uint16),
          rng.integers(-32768, 32767, (h, w)).astype(np.int16),
          (rng.random((h, w)) * 1000 - 300).astype(np.float32)]
  dx, dy = (m.astype(np.float32) for m in _maps(rng, oh, ow, h, w)[kind])
  for nn in (True, False):
    c1, c2 = cv.convertMaps(dx, dy, dstmap1type=cv.CV_16SC2, nninterpolation=nn)
    xy, frac = wo.convert_maps(dx, dy, nn)
    np.testing.assert_array_equal(xy, c1)
    if not nn:
      np.testing.assert_array_equal(frac, c2)
    for method in (['nearest'] if nn else ['linear', 'cubic', 'lanczos']):
      for img in imgs:
        want = cv.remap(img, c1, None if nn else c2, interpolation=flags[method])
        np.testing.assert_array_equal(wo.remap(img, xy, frac, method), want,
                                      err_msg=f'{method} {img.dtype}')
  labels = rng.integers(0, 2**31 - 1, (h, w)).astype(np.int32)
  c1, _ = cv.convertMaps(dx, dy, dstmap1type=cv.CV_16SC2, nninterpolation=True)
  np.testing.assert_array_equal(wo.remap(labels, c1, None, 'nearest'),
                                cv.remap(labels, c1, None, interpolation=cv.INTER_NEAREST))


def test_oracle_passes_the_reference_kats():
  wo.check_reference_kats(wo.warp_subvolume)
