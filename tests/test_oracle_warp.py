"""Oracle of the warping path: the restated scipy.ndimage.map_coordinates against the
real routine (SciPy is installed), and warp.ndimage_warp against golden vectors from the
reference's own warp.py (tests/golden/make_warp_golden.py) and its KATs
(tests/warp_test.py:80-113).  CPU only."""

import os

import numpy as np
import pytest
from scipy import ndimage

from oracle import warp_oracle as wo

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'warp_golden.npz')


@pytest.fixture(scope='module')
def g():
  return np.load(GOLDEN)


@pytest.mark.parametrize('shape', [(37, 41), (9, 13, 11)])
@pytest.mark.parametrize('dtype', [np.float64, np.float32, np.uint8, np.uint16])
@pytest.mark.parametrize('order', [0, 1])
def test_map_coordinates_equals_scipy(shape, dtype, order):
  rng = np.random.default_rng(len(shape) * 7 + order)
  a = (rng.random(shape) * 250).astype(dtype)
  coords = [rng.random(5000) * (n + 1) - 1 for n in shape]
  for c, n in zip(coords, shape):  # exact integers and both edges
    c[:50] = rng.integers(0, n, 50)
    c[50:60] = n - 1
    c[60:70] = 0
  np.testing.assert_array_equal(wo.map_coordinates(a, coords, order),
                                ndimage.map_coordinates(a, coords, order=order))


def test_ndimage_warp_golden(g):
  img, cm = g['w2_image'], g['w2_map']
  for order in (0, 1):
    np.testing.assert_array_equal(wo.ndimage_warp(img, cm, (5, 5), order=order),
                                  g[f'w2_u8_o{order}'])
  np.testing.assert_array_equal(
      wo.ndimage_warp(img.astype(np.float32) / np.float32(7), cm.astype(np.float32), (5, 5)),
      g['w2_f32_o1'])
  for order in (0, 1):
    np.testing.assert_array_equal(
        wo.ndimage_warp(g['w3_image'], g['w3_map'], (2, 4, 5), order=order),
        g[f'w3_u16_o{order}'])
  kw = dict(image_start=(20, 30, 2), map_start=(3, 5, 1), out_start=(25, 38, 3),
            out_size=(50, 40, 10))
  np.testing.assert_array_equal(wo.ndimage_warp(g['w3_image'], g['w3_map'], (2, 4, 5), **kw),
                                g['w3_boxes'])
  np.testing.assert_array_equal(
      wo.ndimage_warp(g['w3_image'], g['w3_map'], (2, 4, 5), out_scale=(0.5, 0.5, 1.0), **kw),
      g['w3_scale'])


def test_ndimage_warp_reference_kats():
  # tests/warp_test.py:97-113
  image = np.zeros((10, 100, 100), dtype=np.uint16)
  image[5, 40, 30] = 42
  image[4, 50, 40] = 16
  coord_map = np.zeros((3, 10, 25, 25))
  coord_map[0], coord_map[1], coord_map[2] = 10, 17, 2
  expected = np.zeros((10, 100, 100))
  expected[3, 23, 20] = 42
  expected[2, 33, 30] = 16
  np.testing.assert_array_equal(wo.ndimage_warp(image, coord_map, (1, 4, 5)), expected)
