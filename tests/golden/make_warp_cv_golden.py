"""Golden vectors for `warp.warp_subvolume` from the reference's own source
(/root/reference/warp.py:58-186), run with the real cv2 and scipy of this image:

  python tests/golden/make_warp_cv_golden.py      # needs /root/reference and cv2

Only the un-vendored connectomics helpers are stubbed: bounding boxes (sofima_b200.compat)
and `labels.make_contiguous` / `labels.relabel` (sorted unique ids with 0 pinned at index 0,
and a table lookup), plus skimage and BoxGenerator, which warp_subvolume never calls.
"""
import os
import sys
import types

import numpy as np
import scipy.ndimage as ndi

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, '..', '..'))
import _jax_shim as shim  # pylint: disable=g-import-not-at-top


def _make_contiguous(labels):
  orig = np.unique(np.append(labels.ravel(), np.uint64(0)))
  low = np.arange(len(orig), dtype=np.uint64)
  return np.searchsorted(orig, labels).astype(np.uint64), list(zip(orig, low))


def _relabel(labels, orig_ids, new_ids):
  order = np.argsort(orig_ids)
  pos = np.searchsorted(orig_ids[order], labels)
  return new_ids[order][pos]


def load_warp():
  import cv2  # the real one  # pylint: disable=unused-import
  shim.install()
  from sofima_b200 import compat
  for name in ('skimage', 'skimage.exposure', 'connectomics.segmentation',
               'connectomics.segmentation.labels', 'connectomics.common.box_generator'):
    sys.modules.setdefault(name, types.ModuleType(name))
  sys.modules['skimage'].exposure = sys.modules['skimage.exposure']
  sys.modules['connectomics.common.box_generator'].BoxGenerator = object
  lab = sys.modules['connectomics.segmentation.labels']
  lab.make_contiguous, lab.relabel = _make_contiguous, _relabel
  sys.modules['connectomics.segmentation'].labels = lab
  import connectomics
  connectomics.common.box_generator = sys.modules['connectomics.common.box_generator']
  connectomics.segmentation = sys.modules['connectomics.segmentation']
  return shim.load_reference('warp'), compat


def main():
  warp, compat = load_warp()
  rng = np.random.default_rng(23)
  out = {}
  Box = compat.BoundingBox

  def smooth(shape, sig, amp):
    return ndi.gaussian_filter(rng.standard_normal(shape), sig) * amp

  # Case A: [2, 3, 104, 120] images; the map (stride 8) has context around the output box,
  # the output box sticks out of the image on two sides; section 1 of the map is all NaN
  # (skipped), section 2 has a NaN hole and a far-away node.
  n, nz, h, w = 2, 3, 104, 120
  image_box = Box(start=(40, 64, 5), size=(w, h, nz))
  map_box = Box(start=(3, 6, 5), size=(19, 16, nz))
  out_box = Box(start=(30, 70, 5), size=(128, 96, nz))
  cmap = np.stack([smooth((nz, 16, 19), 2, 25), smooth((nz, 16, 19), 2, 25)])
  cmap[:, 1] = np.nan
  cmap[:, 2, 4:6, 7:9] = np.nan
  cmap[0, 2, 12, 15] = 1e7
  base = ndi.gaussian_filter(rng.random((n, nz, h, w)), (0, 0, 1.2, 1.2))
  base = (base - base.min()) / (base.max() - base.min())
  out['a_map'] = cmap
  out['a_boxes'] = np.array([image_box.start, image_box.size, map_box.start, map_box.size,
                             out_box.start, out_box.size])
  imgs = {'u8': (base * 255).astype(np.uint8), 'u16': (base * 65535).astype(np.uint16),
          'f32': (base * 900 - 300).astype(np.float32)}
  for name, img in imgs.items():
    out[f'a_image_{name}'] = img
    for inter in ('nearest', 'linear', 'cubic', 'lanczos'):
      out[f'a_{name}_{inter}'] = warp.warp_subvolume(
          img, image_box, cmap, map_box, 8, out_box, interpolation=inter)
  # default interpolation, float32 map, float stride, the deprecated offset
  out['a_u8_default_f32map'] = warp.warp_subvolume(
      imgs['u8'], image_box, cmap.astype(np.float32), map_box, 8, out_box)
  out['a_u8_offset'] = warp.warp_subvolume(
      imgs['u8'], image_box, cmap, map_box, 8.0, out_box, interpolation='linear', offset=0.5)
  # uint32 intensities below 2**16 are warped as uint16 (warp.py:109-115)
  out['a_u32_linear'] = warp.warp_subvolume(
      imgs['u16'].astype(np.uint32), image_box, cmap, map_box, 8, out_box,
      interpolation='linear')
  # Case B: uint64 labels beyond int32 (nearest neighbour on contiguous ids)
  seg = (rng.integers(0, 40, (1, nz, h // 10 + 1, w // 10 + 1)).astype(np.uint64)
         * np.uint64(2**40 + 12345))
  seg = np.kron(seg, np.ones((1, 1, 10, 10), np.uint64))[:, :, :h, :w]
  out['b_seg'] = seg
  out['b_seg_warped'] = warp.warp_subvolume(seg, image_box, cmap, map_box, 8, out_box)
  # Case C: render_tiles on a 2 x 2 grid of 160 x 200 tiles (stride 20), smooth forward maps,
  # one NaN node, one tile mask, margin overrides; canvas inferred.
  th, tw, st = 160, 200, 20
  big = ndi.gaussian_filter(rng.random((2 * th + 60, 2 * tw + 60)), 1.5)
  big = ((big - big.min()) / (big.max() - big.min()) * 254 + 1).astype(np.uint8)
  tiles, maps = {}, {}
  for ty_ in range(2):
    for tx_ in range(2):
      tiles[tx_, ty_] = np.ascontiguousarray(
          big[30 + ty_ * (th - 20):30 + ty_ * (th - 20) + th,
              30 + tx_ * (tw - 20):30 + tx_ * (tw - 20) + tw])
      m = np.stack([smooth((1, th // st, tw // st), 2, 12), smooth((1, th // st, tw // st), 2, 12)])
      m[0] -= 10 * tx_
      m[1] -= 10 * ty_
      maps[tx_, ty_] = m
  maps[1, 0][:, 0, 3, 4] = np.nan
  mask = np.ones((th, tw), np.uint8)
  mask[40:70, 50:90] = 0
  for (tx_, ty_), t in tiles.items():
    out[f'c_tile_{tx_}{ty_}'] = t
    out[f'c_map_{tx_}{ty_}'] = maps[tx_, ty_]
  out['c_mask_01'] = mask
  canvas, cov, wt = warp.render_tiles(tiles, maps, stride=(st, st), margin=10,
                                      return_warped_tiles=True, tile_masks={(0, 1): mask},
                                      margin_overrides={(1, 1): (5, 8, 12, 3)})
  out['c_canvas'], out['c_covered'] = canvas, cov
  for (tx_, ty_), (x0, y0, w) in wt.items():
    out[f'c_warped_{tx_}{ty_}'] = w
    out[f'c_pos_{tx_}{ty_}'] = np.array([x0, y0])
  path = os.path.join(HERE, 'warp_cv_golden.npz')
  np.savez_compressed(path, **out)
  print('warp_cv_golden.npz', os.path.getsize(path) // 1024, 'KiB')


if __name__ == '__main__':
  main()
