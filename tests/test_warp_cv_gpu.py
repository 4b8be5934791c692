"""GPU parity tests of sofima_b200.warp.warp_subvolume through the C ABI:
bit-exact against the outputs of the reference's warp.warp_subvolume (scipy + OpenCV,
tests/golden/warp_cv_golden.npz) and against the oracle on seeded random cases."""

import os

import numpy as np
import pytest
import scipy.ndimage as ndi

from oracle import warp_cv_oracle as wo
from sofima_b200 import compat

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'warp_cv_golden.npz')


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


def boxes(g):
  b = g['a_boxes']
  return tuple(compat.BoundingBox(start=b[i], size=b[i + 1]) for i in (0, 2, 4))


@pytest.mark.parametrize('dt', ['u8', 'u16', 'f32'])
@pytest.mark.parametrize('inter', ['nearest', 'linear', 'cubic', 'lanczos'])
def test_reference_outputs(warp, g, dt, inter):
  ib, mb, ob = boxes(g)
  got = warp.warp_subvolume(g[f'a_image_{dt}'], ib, g['a_map'], mb, 8, ob, interpolation=inter)
  want = g[f'a_{dt}_{inter}']
  assert got.dtype == want.dtype and got.shape == want.shape
  np.testing.assert_array_equal(got, want)


def test_reference_variants(warp, g):
  ib, mb, ob = boxes(g)
  np.testing.assert_array_equal(
      warp.warp_subvolume(g['a_image_u8'], ib, g['a_map'].astype(np.float32), mb, 8, ob),
      g['a_u8_default_f32map'])
  np.testing.assert_array_equal(
      warp.warp_subvolume(g['a_image_u8'], ib, g['a_map'], mb, 8.0, ob, interpolation='linear',
                          offset=0.5), g['a_u8_offset'])
  got = warp.warp_subvolume(g['a_image_u16'].astype(np.uint32), ib, g['a_map'], mb, 8, ob,
                            interpolation='linear')
  assert got.dtype == np.uint32
  np.testing.assert_array_equal(got, g['a_u32_linear'])
  got = warp.warp_subvolume(g['b_seg'], ib, g['a_map'], mb, 8, ob)
  assert got.dtype == np.uint64
  np.testing.assert_array_equal(got, g['b_seg_warped'])
  # the cv2 flag values are accepted like the names (INTER_LANCZOS4 = 4)
  np.testing.assert_array_equal(
      warp.warp_subvolume(g['a_image_u8'], ib, g['a_map'], mb, 8, ob, interpolation=4),
      g['a_u8_lanczos'])


def test_errors(warp, g):
  ib, mb, ob = boxes(g)
  with pytest.raises(ValueError):
    warp.warp_subvolume(np.full((1, 3, 4, 4), 2**16, np.uint32), ib, g['a_map'], mb, 8, ob)
  with pytest.raises(KeyError):
    warp.warp_subvolume(g['a_image_u8'], ib, g['a_map'], mb, 8, ob, interpolation='spline')
  with pytest.raises(NotImplementedError):
    warp.warp_subvolume(g['a_image_u8'].astype(np.float64), ib, g['a_map'], mb, 8, ob)


@pytest.mark.parametrize('seed', [0, 1, 2])
def test_random_cases_against_oracle(warp, seed):
  rng = np.random.default_rng(100 + seed)
  n, nz = int(rng.integers(1, 3)), int(rng.integers(1, 4))
  h, w = int(rng.integers(60, 200)), int(rng.integers(60, 200))
  stride = [4, 6.5, 10][seed]
  my, mx = int(rng.integers(6, 20)), int(rng.integers(6, 20))
  ib = compat.BoundingBox(start=rng.integers(0, 50, 3), size=(w, h, nz))
  mb = compat.BoundingBox(start=(int(ib.start[0] // stride), int(ib.start[1] // stride),
                                 ib.start[2]), size=(mx, my, nz))
  ob = compat.BoundingBox(start=(ib.start[0] + int(rng.integers(-10, 10)),
                                 ib.start[1] + int(rng.integers(-10, 10)), ib.start[2]),
                          size=(int(rng.integers(40, 220)), int(rng.integers(40, 220)), nz))
  cmap = np.stack([ndi.gaussian_filter(rng.standard_normal((nz, my, mx)), 1.5) * 30
                   for _ in range(2)]).astype([np.float64, np.float32, np.float64][seed])
  if seed == 1:
    cmap[:, :, 2, 3] = np.nan
  base = ndi.gaussian_filter(rng.random((n, nz, h, w)), (0, 0, 1, 1))
  base = (base - base.min()) / (base.max() - base.min())
  for img in ((base * 255).astype(np.uint8), (base * 65535).astype(np.uint16),
              (base * 60000 - 30000).astype(np.int16), (base * 50 - 10).astype(np.float32)):
    for inter in ('nearest', 'linear', 'cubic', 'lanczos'):
      want = wo.warp_subvolume(img, ib, cmap, mb, stride, ob, interpolation=inter)
      got = warp.warp_subvolume(img, ib, cmap, mb, stride, ob, interpolation=inter)
      np.testing.assert_array_equal(got, want, err_msg=f'{img.dtype} {inter}')
      assert want.any()


def test_device_resident_section_stack(warp, g):
  import torch
  ib, mb, ob = boxes(g)
  img = torch.from_numpy(g['a_image_u8']).cuda()
  got = warp.warp_subvolume(img, ib, g['a_map'], mb, 8, ob, interpolation='lanczos')
  assert got.is_cuda and got.dtype == torch.uint8
  np.testing.assert_array_equal(got.cpu().numpy(), g['a_u8_lanczos'])


def test_full_size_section_properties(warp):
  """A 4096 x 4096 section: the identity map reproduces the image for every method, and an
  integer translation reproduces the shifted image (zero outside)."""
  rng = np.random.default_rng(7)
  img = rng.integers(0, 256, (1, 1, 4096, 4096), dtype=np.uint8)
  box = compat.BoundingBox(start=(0, 0, 0), size=(4096, 4096, 1))
  mbox = compat.BoundingBox(start=(0, 0, 0), size=(129, 129, 1))
  zero = np.zeros((2, 1, 129, 129))
  for inter in ('nearest', 'linear', 'cubic', 'lanczos'):
    np.testing.assert_array_equal(
        warp.warp_subvolume(img, box, zero, mbox, 32, box, interpolation=inter), img)
  shift = zero.copy()
  shift[0] += 5
  shift[1] -= 3
  got = warp.warp_subvolume(img, box, shift, mbox, 32, box, interpolation='lanczos')
  want = np.zeros_like(img)
  want[0, 0, 3:, :-5] = img[0, 0, :-3, 5:]
  np.testing.assert_array_equal(got, want)


def test_warp_by_map_processor(warp):
  """processor.warp.WarpByMap (processor/warp.py:541-623) over in-memory volumes: every
  section equals warp_subvolume on the boxes the processor derives; an integer shift map
  renders the shifted data; 2x downsampling is the block mean of the full-resolution render."""
  from sofima_b200.processor import warp as pwarp
  rng = np.random.default_rng(5)
  data = rng.integers(0, 256, (1, 3, 300, 320), dtype=np.uint8)
  stride = 20
  cmap = np.zeros((2, 3, 300 // stride + 1, 320 // stride + 1), np.float32)
  cmap[0] += 6
  cmap[1] -= 4
  cmap[:, 1] = np.nan  # a section without a map stays empty
  box = compat.BoundingBox(start=(40, 20, 0), size=(200, 160, 3))
  proc = pwarp.WarpByMap(pwarp.WarpByMap.Config(stride=stride, map_volinfo=cmap,
                                                data_volinfo=data, interpolation='lanczos'))
  out = proc.process(compat.Subvolume(np.zeros((1, 3, 160, 200), np.uint8), box))[0]
  assert out.bbox == box
  want = np.zeros((1, 3, 160, 200), np.uint8)
  want[:, ::2] = data[:, ::2, 20 - 4:180 - 4, 40 + 6:240 + 6]
  np.testing.assert_array_equal(out.data, want)
  # smooth map + downsampling
  cmap2 = np.stack([ndi.gaussian_filter(rng.standard_normal(cmap.shape[1:]), (0, 2, 2)) * 20
                    for _ in range(2)]).astype(np.float32)
  full = pwarp.WarpByMap(pwarp.WarpByMap.Config(stride=stride, map_volinfo=cmap2,
                                                data_volinfo=data, interpolation='linear'))
  ref = full.process(compat.Subvolume(np.zeros((1, 3, 160, 200), np.uint8), box))[0].data
  half = pwarp.WarpByMap(pwarp.WarpByMap.Config(stride=stride // 2, map_volinfo=cmap2,
                                                data_volinfo=data, interpolation='linear',
                                                downsample=2))
  hbox = compat.BoundingBox(start=(20, 10, 0), size=(100, 80, 3))
  got = half.process(compat.Subvolume(np.zeros((1, 3, 80, 100), np.uint8), hbox))[0].data
  mean = ref.astype(np.int64).reshape(1, 3, 80, 2, 100, 2).sum(axis=(3, 5)) / 4.0
  np.testing.assert_array_equal(got, mean.astype(np.uint8))


def test_render_tiles_reference_outputs(warp, g):
  """warp.render_tiles (warp.py:338-535) on a 2 x 2 tile grid: canvas, coverage mask and the
  per-tile renders equal the reference's (scipy map inversion + OpenCV Lanczos remap)."""
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
  # without the extras only two values come back
  assert len(warp.render_tiles(tiles, maps, stride=(20, 20), margin=10)) == 2
  with pytest.raises(NotImplementedError):
    warp.render_tiles(tiles, maps, stride=(20, 10))


def test_stitch_and_render_3d_tiles(warp):
  """processor.warp.StitchAndRender3dTiles (processor/warp.py:38-343): 2 x 2 tiles cut from
  one volume with 12 px of overlap and meshes that encode exactly that placement render back
  to the volume (blended regions may lose one grey level to the float32 average, as in the
  reference); a subvolume in the middle equals the same region of the full render."""
  from sofima_b200.processor import warp as pwarp
  rng = np.random.default_rng(9)
  big = ndi.gaussian_filter(rng.random((16, 200, 200)), 1.0)
  big = ((big - big.min()) / (big.max() - big.min()) * 250 + 2).astype(np.uint8)
  tz, ty, tx, step = 16, 96, 96, 84
  tiles, key_to_idx = {}, {}
  mesh = np.zeros((3, 4, 2, 6, 6))
  for y in range(2):
    for x in range(2):
      idx = len(key_to_idx)
      key_to_idx[x, y] = idx
      tiles[10 + idx] = big[:, y * step:y * step + ty, x * step:x * step + tx]
      mesh[0, idx] = -(tx - step) * x
      mesh[1, idx] = -(ty - step) * y

  class Renderer(pwarp.StitchAndRender3dTiles):
    def _open_tile_volume(self, tile_id):
      return tiles[tile_id]

  Renderer.reset_cache()
  r = Renderer(tile_map=[[10, 11], [12, 13]], tile_pattern_path='{tile_id}',
               tile_mesh_path={'key_to_idx': key_to_idx, 'x': mesh}, stride=(8, 16, 16))
  box = compat.BoundingBox(start=(0, 0, 0), size=(180, 180, 16))
  out = r.process(compat.Subvolume(np.zeros((1, 16, 180, 180), np.uint8), box))
  assert out.bbox == box and out.data.dtype == np.uint8
  want = big[None, :, :180, :180].astype(int)
  diff = want - out.data.astype(int)
  assert diff.min() >= 0 and diff.max() <= 1, (diff.min(), diff.max())
  assert (diff == 0).mean() > 0.9
  # interior of a single tile: one source, weight / weight
  assert np.abs(diff[0, :, 20:60, 20:60]).max() <= 1
  sub = compat.BoundingBox(start=(60, 70, 4), size=(64, 48, 8))
  part = r.process(compat.Subvolume(np.zeros((1, 8, 48, 64), np.uint8), sub))
  np.testing.assert_array_equal(part.data, out.data[:, 4:12, 70:118, 60:124])
  # the margin keeps the seam away from the inner tile edges
  Renderer.reset_cache()
  r2 = Renderer(tile_map=[[10, 11], [12, 13]], tile_pattern_path='{tile_id}', margin=4,
                tile_mesh_path={'key_to_idx': key_to_idx, 'x': mesh}, stride=(8, 16, 16))
  out2 = r2.process(compat.Subvolume(np.zeros((1, 16, 180, 180), np.uint8), box))
  d2 = want - out2.data.astype(int)
  # (the last row / column of outer tiles carries no weight with a margin: reference quirk)
  assert np.abs(d2[0, :, :179, :179]).max() <= 1
  Renderer.reset_cache()


def test_reference_kats(warp):
  """The reference's own warp_subvolume tests (tests/warp_test.py:27-78): label translation
  with ids beyond int32 and a 45-degree rotation of a rhombus (Lanczos default)."""
  wo.check_reference_kats(warp.warp_subvolume, compat.BoundingBox)
