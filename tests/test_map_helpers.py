"""Host bookkeeping of map_utils (to_absolute / to_relative / outer_box): the reference's own
known-answer tests (tests/map_utils_test.py:93-127, 155-168)."""

import numpy as np

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
