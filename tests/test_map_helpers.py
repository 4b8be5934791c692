"""Host bookkeeping of map_utils (to_absolute / to_relative / outer_box): the reference's own
known-answer tests (tests/map_utils_test.py:93-127, 155-168)."""

import numpy as np
import pytest

from sofima_b200 import compat
from sofima_b200 import map_utils


def test_abs_rel_conversion():
  rng = np.random.default_rng(11111)
  rel = rng.uniform(-0.5, 0.5, [2, 1, 50, 50])
  np.testing.assert_allclose(map_utils.to_relative(map_utils.to_absolute(rel, 10), 10), rel)
  box = compat.BoundingBox(start=(240, 280, 300), size=(50, 50, 1))
  abs_map = map_utils.to_absolute(rel, 10, box)
  np.testing.assert_allclose(map_utils.to_relative(abs_map, 10, box), rel)
  # node (y=3, x=7) of the boxed map sits at ((240 + 7) * 10, (280 + 3) * 10)
  np.testing.assert_allclose(abs_map[:, 0, 3, 7] - rel[:, 0, 3, 7], [2470.0, 2830.0])
  assert rel is not abs_map and np.abs(rel).max() <= 0.5  # the input is not modified


def test_abs_rel_conversion_3d():
  rng = np.random.default_rng(11111)
  rel = rng.uniform(-0.5, 0.5, [3, 25, 50, 50])
  np.testing.assert_allclose(map_utils.to_relative(map_utils.to_absolute(rel, 7), 7), rel)
  box = compat.BoundingBox(start=(240, 280, 300), size=(50, 50, 25))
  np.testing.assert_allclose(
      map_utils.to_relative(map_utils.to_absolute(rel, 7, box), 7, box), rel)


def test_outer_box():
  box = compat.BoundingBox(start=(100, 200, 10), size=(50, 50, 1))
  coord_map = np.zeros([2, 1, 50, 50])
  coord_map[0, 0, 0, 49] = 4
  coord_map[0, 0, 1, 49] = 8
  coord_map[0, 0, 2, 0] = -3
  coord_map[1, 0, 49, 10] = 1
  coord_map[1, 0, 0, 1] = -2
  assert map_utils.outer_box(coord_map, box, stride=5) == compat.BoundingBox(
      start=(99, 199, 10), size=(53, 52, 1))


def test_fill_missing():
  hy, hx = np.mgrid[:50, :50]
  coord_map = np.zeros([2, 1, 50, 50])
  coord_map[0, 0, ...] = np.sin(hx / 25)
  coord_map[1, 0, ...] = np.cos(hy / 25)
  with_gap = coord_map.copy()
  with_gap[:, 0, 24:28, 38:42] = np.nan
  np.testing.assert_array_almost_equal(map_utils.fill_missing(with_gap), coord_map, decimal=2)
  with_gap = coord_map.copy()
  with_gap[:, 0, -1, :] = np.nan
  assert np.all(np.isnan(map_utils.fill_missing(with_gap)[:, 0, -1, :]))
  filled = map_utils.fill_missing(with_gap, extrapolate=True)
  np.testing.assert_array_almost_equal(filled[1, 0, -1, :], coord_map[1, 0, -1, :], decimal=1)
  with_gap[...] = np.nan
  assert np.all(map_utils.fill_missing(with_gap, invalid_to_zero=True) == 0)
  assert map_utils.fill_missing(coord_map) is coord_map  # nothing to do: same object


def test_inner_box():
  box = compat.BoundingBox(start=(100, 200, 10), size=(50, 50, 1))
  coord_map = np.zeros([2, 1, 50, 50])
  coord_map[1, :, ...] = -30
  coord_map[1, :, 0, :] = -40
  coord_map[1, :, -1, :] = -25
  assert map_utils.inner_box(coord_map, box, stride=10) == compat.BoundingBox(
      start=(100, 196, 10), size=(50, 51, 1))
  coord_map = np.zeros([2, 1, 50, 50])
  coord_map[0, :, :, 0] = -9
  coord_map[0, :, :, -1] = 9
  assert map_utils.inner_box(coord_map, box, stride=10) == compat.BoundingBox(
      start=(100, 200, 10), size=(50, 50, 1))


def test_inner_box3d():
  box = compat.BoundingBox(start=(100, 200, 200), size=(50, 50, 50))
  coord_map = np.zeros([3, 50, 50, 50])
  coord_map[2, ...] = -30
  coord_map[2, 0, :, :] = -40
  coord_map[2, -1, :, :] = -25
  assert map_utils.inner_box(coord_map, box, stride=10) == compat.BoundingBox(
      start=(100, 200, 196), size=(50, 50, 51))


def test_invert_map():
  box = compat.BoundingBox(start=(100, 200, 10), size=(50, 50, 1))
  _, hx = np.mgrid[:50, :50]
  coord_map = np.zeros([2, 1, 50, 50])
  coord_map[1, 0, ...] = np.sin(hx / 25) * 20
  inv_map = map_utils.invert_map(coord_map, box, box, 40.0)
  np.testing.assert_array_almost_equal(inv_map[:, :, 1:, 1:], -coord_map[:, :, 1:, 1:],
                                       decimal=5)


def test_invert_map_3d():
  box = compat.BoundingBox(start=(100, 200, 10), size=(20, 20, 5))
  _, _, hx = np.mgrid[:5, :20, :20]
  coord_map = np.zeros([3, 5, 20, 20])
  coord_map[1, ...] = np.sin(hx / 25) * 20
  inv_map = map_utils.invert_map(coord_map, box, box, 40.0)
  np.testing.assert_array_almost_equal(inv_map[:, 1:, 1:, 1:], -coord_map[:, 1:, 1:, 1:],
                                       decimal=5)


def test_resample_map():
  box = compat.BoundingBox(start=(100, 200, 10), size=(50, 50, 1))
  hy, hx = np.mgrid[:50, :50]
  coord_map = np.zeros([2, 1, 50, 50])
  coord_map[0, 0, ...] = np.sin(hx / 25) * 20
  coord_map[1, 0, ...] = np.cos(hy / 25) * 20
  hy, hx = np.mgrid[:100, :100]
  expected = np.zeros([2, 1, 100, 100])
  expected[0, 0, ...] = np.sin(hx / 50) * 20
  expected[1, 0, ...] = np.cos(hy / 50) * 20
  dst_box = compat.BoundingBox(start=(102, 203, 10), size=(48, 47, 1)).scale([2, 2, 1.0])
  resampled = map_utils.resample_map(coord_map, box, dst_box, 40, 20)
  np.testing.assert_array_almost_equal(resampled[:, :, :-1, :-1], expected[:, :, 6:-1, 4:-1],
                                       decimal=2)


def test_tile_blending_weight_equals_distance_transform():
  """StitchAndRender3dTiles._get_dts (analytic, rectangle) == the Euclidean distance
  transform with a black border that the reference computes with `edt`
  (processor/warp.py:151-165), for inner / outer tiles with and without a margin."""
  from sofima_b200.processor import warp as pwarp
  for margin in (0, 3, 7):
    r = pwarp.StitchAndRender3dTiles(tile_map=[[1, 2, 3], [4, 5, 6], [7, 8, 9]],
                                     tile_mesh_path=None, tile_pattern_path='', stride=(8, 8, 8),
                                     margin=margin)
    for tx in range(3):
      for ty in range(3):
        got = r._get_dts((4, 37, 45), tx, ty)
        mask = np.zeros((37, 45), bool)
        if margin > 0:
          x0 = margin if tx > 0 else 0
          x1 = -margin if tx < 2 else -1
          y0 = margin if ty > 0 else 0
          y1 = -margin if ty < 2 else -1
          mask[y0:y1, x0:x1] = 1
        else:
          mask[...] = 1
        want = pwarp._border_distance(mask)
        assert got.dtype == np.float32
        np.testing.assert_array_equal(got, want)


def test_compose_maps():
  """tests/map_utils_test.py:248-264: a map composed with its inverse is the identity."""
  box = compat.BoundingBox(start=(100, 200, 10), size=(50, 50, 1))
  coord_map = np.zeros([2, 1, 50, 50])
  hy, hx = np.mgrid[:50, :50]
  coord_map[0, 0, ...] = np.sin(hx / 25)
  coord_map[1, 0, ...] = np.cos(hy / 25)
  inverted = map_utils.invert_map(coord_map, box, box, 5)
  composed = map_utils.compose_maps(coord_map, box, 5, inverted, box, 5)[:, :, 1:-2, 1:-2]
  np.testing.assert_array_almost_equal(composed, np.zeros_like(composed), decimal=3)


def test_make_affine_map():
  box = compat.BoundingBox(start=(10, 20, 3), size=(6, 5, 4))
  ident = np.hstack([np.eye(3), np.zeros((3, 1))])
  np.testing.assert_array_equal(map_utils.make_affine_map(ident, box, (1, 2, 2)),
                                np.zeros((3, 4, 5, 6)))
  shift = ident.copy()
  shift[:, 3] = (1.5, -2.0, 0.25)
  m = map_utils.make_affine_map(shift, box, (1, 2, 2))
  np.testing.assert_allclose(m[0], 1.5)
  np.testing.assert_allclose(m[1], -2.0)
  np.testing.assert_allclose(m[2], 0.25)
  scale = ident.copy()
  scale[0, 0] = 2.0  # x' = 2 x: relative x offset = absolute x position of the node
  m = map_utils.make_affine_map(scale, box, (1, 2, 2))
  np.testing.assert_allclose(m[0, 0, 0], (np.arange(6) * 2 + 10).astype(float))


def test_warp_points():
  """tests/warp_test.py:115-126."""
  from sofima_b200 import warp
  coord_map = np.zeros((2, 10, 3, 3))
  coord_map[0, 0, ...] = 10
  coord_map[1, 1, ...] = 20
  points = np.array([[101, 201, 0], [105, 205, 1]])
  map_box = compat.BoundingBox(start=(10, 20, 0), size=(3, 3, 10))
  warped = warp.warp_points(points, coord_map, map_box, 10)
  np.testing.assert_array_equal(warped, np.array([[111, 201, 0], [105, 225, 1]]))
  fpts = points.astype(np.float64) + 0.25
  fpts[:, 2] = points[:, 2]
  np.testing.assert_allclose(warp.warp_points(fpts, coord_map, map_box, 10),
                             fpts + np.array([[10, 0, 0], [0, 20, 0]]))


def test_mask_irregular_lives_in_map_utils():
  from sofima_b200.processor import mesh as pmesh
  assert pmesh.mask_irregular is map_utils.mask_irregular
  cmap = np.zeros((2, 6, 7), np.float32)
  cmap[0, 2, 3] = -35.0  # node pushed far to the left: the link to it is folded
  bad = map_utils.mask_irregular(cmap, (40.0, 40.0), 0.5, dilation_iters=0)
  assert bad[2, 2] and bad.sum() >= 1 and np.isnan(cmap[0, 2, 2])


def test_map_processors():
  """processor/maps.py:332-498 -- InvertMap, ResampleMap, MaskIrregularities, FillMissing."""
  from sofima_b200.processor import maps
  box = compat.BoundingBox(start=(10, 20, 3), size=(30, 30, 1))
  _, hx = np.mgrid[:30, :30]
  cmap = np.zeros([2, 1, 30, 30], np.float32)
  cmap[1, 0] = np.sin(hx / 15) * 10
  vol = np.zeros((2, 8, 100, 100), np.float32)
  inv = maps.InvertMap(maps.InvertMap.Config(stride=40.0, crop_output=False), vol)
  (out,) = inv.process(compat.Subvolume(cmap, box))
  assert out.bbox == box
  np.testing.assert_array_almost_equal(out.data[:, :, 1:, 1:], -cmap[:, :, 1:, 1:], decimal=4)
  cropped = maps.InvertMap(maps.InvertMap.Config(stride=40.0), vol).process(
      compat.Subvolume(cmap, box))[0]
  assert cropped.bbox == map_utils.inner_box(cmap.astype(np.float64), box, 40.0)
  assert maps.InvertMap(maps.InvertMap.Config(stride=40.0), vol).process(
      compat.Subvolume(np.full_like(cmap, np.nan), box)) == []
  with pytest.raises(ValueError):
    maps.InvertMap(maps.InvertMap.Config(stride=40.0))
  # ResampleMap: twice the node density, values follow the smooth field
  rs = maps.ResampleMap(maps.ResampleMap.Config(stride=40, out_stride=20))
  (fine,) = rs.process(compat.Subvolume(cmap, box))
  assert fine.bbox == box.scale([2, 2, 1.0]) and fine.data.shape == (2, 1, 60, 60)
  np.testing.assert_allclose(fine.data[1, 0, ::2, ::2][:29, :29], cmap[1, 0, :29, :29], atol=1e-6)
  np.testing.assert_allclose(rs.pixelsize(np.array([8.0, 8.0, 30.0])), [4.0, 4.0, 30.0])
  # MaskIrregularities: 3 nodes of context are cropped, a folded node is masked with its ring
  mi = maps.MaskIrregularities((40.0, 40.0), 0.5)
  bad = np.zeros([2, 1, 30, 30], np.float32)
  bad[0, 0, 15, 15] = -35
  got = mi.process(compat.Subvolume(bad, box))
  assert got.data.shape == (2, 1, 24, 24) and np.isnan(got.data[0, 0, 12, 12])
  assert np.isnan(got.data).sum() >= 2 and np.isfinite(got.data[0, 0, 0, 0])
  # FillMissing
  holes = cmap.copy()
  holes[:, 0, 10:12, 10:12] = np.nan
  filled = maps.FillMissing().process(compat.Subvolume(holes, box))
  assert np.isfinite(filled.data).all()
  np.testing.assert_allclose(filled.data, cmap, atol=0.05)
