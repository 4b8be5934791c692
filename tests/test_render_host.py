"""Host logic of the render path on CPU: `warp.render_tiles` and `processor.warp.WarpByMap`
with the per-pixel kernel replaced by the oracle (`oracle/warp_cv_oracle.warp_subvolume`).
The canvas must equal the one the reference itself produced (tests/golden/warp_cv_golden.npz),
which pins the boxes, the SciPy map inversion and the paste rules independently of the GPU."""

import os

import numpy as np
import pytest

from oracle import warp_cv_oracle as wo
from sofima_b200 import compat
from sofima_b200 import warp
from sofima_b200.processor import warp as pwarp

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'warp_cv_golden.npz')


@pytest.fixture()
def oracle_kernel(monkeypatch):
  def fake(image, image_box, coord_map, map_box, stride, out_box, interpolation=None,
           offset=0.0, parallelism=1):
    del parallelism
    return wo.warp_subvolume(image, image_box, coord_map, map_box, stride, out_box,
                             interpolation=interpolation, offset=offset)
  monkeypatch.setattr(warp, 'warp_subvolume', fake)
  return fake


def test_render_tiles_host_logic(oracle_kernel):
  g = np.load(GOLDEN)
  keys = [(0, 0), (0, 1), (1, 0), (1, 1)]
  tiles = {k: g[f'c_tile_{k[0]}{k[1]}'] for k in keys}
  maps = {k: g[f'c_map_{k[0]}{k[1]}'] for k in keys}
  canvas, covered, wt = warp.render_tiles(
      tiles, maps, stride=(20, 20), margin=10, return_warped_tiles=True,
      tile_masks={(0, 1): g['c_mask_01']}, margin_overrides={(1, 1): (5, 8, 12, 3)})
  np.testing.assert_array_equal(canvas, g['c_canvas'])
  np.testing.assert_array_equal(covered, g['c_covered'])
  for k in keys:
    x0, y0, w = wt[k]
    assert [x0, y0] == g[f'c_pos_{k[0]}{k[1]}'].tolist()
    np.testing.assert_array_equal(w, g[f'c_warped_{k[0]}{k[1]}'])


def test_warp_by_map_host_logic(oracle_kernel):
  rng = np.random.default_rng(5)
  data = rng.integers(0, 256, (1, 2, 120, 140), dtype=np.uint8)
  stride = 20
  cmap = np.zeros((2, 2, 120 // stride + 1, 140 // stride + 1), np.float32)
  cmap[0] += 6
  cmap[1] -= 4
  cmap[:, 1] = np.nan
  box = compat.BoundingBox(start=(20, 20, 0), size=(80, 60, 2))
  proc = pwarp.WarpByMap(pwarp.WarpByMap.Config(stride=stride, map_volinfo=cmap,
                                                data_volinfo=data, interpolation='nearest'))
  out = proc.process(compat.Subvolume(np.zeros((1, 2, 60, 80), np.uint8), box))[0]
  assert out.bbox == box
  want = np.zeros((1, 2, 60, 80), np.uint8)
  want[:, 0] = data[:, 0, 20 - 4:80 - 4, 20 + 6:100 + 6]
  np.testing.assert_array_equal(out.data, want)
  # map box: output box in map pixels plus two nodes of context, clipped to the map volume
  map_box, rel = proc._get_map_for_box(compat.BoundingBox(start=(20, 20, 0), size=(80, 60, 1)))
  assert map_box == compat.BoundingBox(start=(0, 0, 0), size=(7, 6, 1))
  assert rel.dtype == np.float64 and rel.shape == (2, 1, 6, 7)
  # a box whose map is all NaN yields nothing to warp
  assert list(proc._generate_boxes_to_warp(
      data, compat.BoundingBox(start=(20, 20, 1), size=(80, 60, 1)))) == []
