"""GPU parity tests of sofima_b200.warp.ndimage_warp through the C ABI:
bit-exact against the golden vectors from the reference's warp.py, the reference's KATs
(tests/warp_test.py:80-113) and the oracle on random maps."""

import os

import numpy as np
import pytest
import scipy.ndimage as ndi

from oracle import warp_oracle as wo
from sofima_b200 import compat

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'warp_golden.npz')


@pytest.fixture(scope='module')
def g():
  return np.load(GOLDEN)


@pytest.fixture(scope='module')
def warp():
  import torch
  if not torch.cuda.is_available():
    pytest.skip('needs a CUDA device')
  from sofima_b200 import warp as w
  return w


def test_golden_2d(warp, g):
  img, cm = g['w2_image'], g['w2_map']
  for order in (0, 1):
    got = warp.ndimage_warp(img, cm, (5, 5), (64, 64), (0, 0), order=order)
    assert got.dtype == np.uint8
    np.testing.assert_array_equal(got, g[f'w2_u8_o{order}'])
  got = warp.ndimage_warp(img.astype(np.float32) / np.float32(7), cm.astype(np.float32),
                          (5, 5), (64, 64), (0, 0))
  assert got.dtype == np.float32
  np.testing.assert_array_equal(got, g['w2_f32_o1'])


def test_golden_3d_and_boxes(warp, g):
  vol, cm = g['w3_image'], g['w3_map']
  for order in (0, 1):
    got = warp.ndimage_warp(vol, cm, (2, 4, 5), (32, 32, 8), (2, 2, 2), order=order)
    np.testing.assert_array_equal(got, g[f'w3_u16_o{order}'])
  boxes = dict(image_box=compat.BoundingBox(start=(20, 30, 2), size=(70, 60, 14)),
               map_box=compat.BoundingBox(start=(3, 5, 1), size=(14, 15, 7)),
               out_box=compat.BoundingBox(start=(25, 38, 3), size=(50, 40, 10)))
  got = warp.ndimage_warp(vol, cm, (2, 4, 5), (32, 32, 8), (2, 2, 2), **boxes)
  np.testing.assert_array_equal(got, g['w3_boxes'])
  got = warp.ndimage_warp(vol, cm, (2, 4, 5), (32, 32, 8), (2, 2, 2),
                          out_scale=(0.5, 0.5, 1.0), **boxes)
  np.testing.assert_array_equal(got, g['w3_scale'])


def test_reference_kats(warp):
  # tests/warp_test.py:80-95 (uint64 labels, order 0)
  image = np.zeros((100, 100), dtype=np.uint64)
  image[40, 30] = 42
  image[50, 40] = 2**40
  coord_map = np.zeros((2, 25, 25))
  coord_map[0], coord_map[1] = 10, 17
  warped = warp.ndimage_warp(image, coord_map, (4, 5), (100, 100), (0, 0), order=0)
  expected = np.zeros((100, 100))
  expected[23, 20] = 42
  expected[33, 30] = 2**40
  np.testing.assert_array_equal(warped, expected)
  # tests/warp_test.py:97-113 (3-d uint16)
  image = np.zeros((10, 100, 100), dtype=np.uint16)
  image[5, 40, 30] = 42
  image[4, 50, 40] = 16
  coord_map = np.zeros((3, 10, 25, 25))
  coord_map[0], coord_map[1], coord_map[2] = 10, 17, 2
  warped = warp.ndimage_warp(image, coord_map, (1, 4, 5), (50, 50, 5), (2, 2, 2))
  expected = np.zeros((10, 100, 100))
  expected[3, 23, 20] = 42
  expected[2, 33, 30] = 16
  np.testing.assert_array_equal(warped, expected)


@pytest.mark.parametrize('dtype', [np.uint8, np.uint16, np.float32])
def test_random_maps_vs_oracle(warp, dtype):
  rng = np.random.default_rng(3)
  img = (ndi.gaussian_filter(rng.random((700, 900)), 1.5) * 250).astype(dtype)
  cm = np.stack([ndi.gaussian_filter(rng.standard_normal((36, 46)), 4) * 300,
                 ndi.gaussian_filter(rng.standard_normal((36, 46)), 4) * 300])
  cm[:, 3, 3] = np.nan  # invalid map node: scipy propagates NaN coordinates -> 0
  for order in (0, 1):
    got = warp.ndimage_warp(img, cm, (20, 20), (256, 256), (8, 8), order=order)
    want = wo.ndimage_warp(img, cm, (20, 20), order=order)
    np.testing.assert_array_equal(got, want)
  import torch
  if dtype != np.uint16:
    got_t = warp.ndimage_warp(torch.from_numpy(img).cuda(), cm, (20, 20), (256, 256), (8, 8))
    assert got_t.is_cuda
    np.testing.assert_array_equal(got_t.cpu().numpy(), wo.ndimage_warp(img, cm, (20, 20)))


def test_errors(warp):
  img = np.zeros((10, 10), np.uint8)
  cm = np.zeros((2, 5, 5))
  with pytest.raises(ValueError):
    warp.ndimage_warp(np.zeros((2, 10, 10), np.uint8), cm, (2, 2), (8, 8), (0, 0))
  with pytest.raises(NotImplementedError):
    warp.ndimage_warp(img, cm, (2, 2), (8, 8), (0, 0), order=3)
  with pytest.raises(ValueError):
    warp.ndimage_warp(img, cm, (2, 2), (8, 8), (0, 0),
                      map_box=compat.BoundingBox(start=(0, 0, 0), size=(5, 5, 1)))
