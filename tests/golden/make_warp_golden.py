"""Golden vectors for the warping path from the reference's own warp.ndimage_warp
(/root/reference/warp.py:189-335; NumPy + SciPy, no JAX):

  python tests/golden/make_warp_golden.py      # needs /root/reference

cv2, skimage and the un-vendored connectomics helpers are stubbed; the only stub that is
executed is BoxGenerator, here a single box covering the whole output -- the reference's
box tiling does not change any output value (every voxel is sampled independently).
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


def load_warp():
  shim.install()
  from sofima_b200 import compat

  class BoxGenerator:
    def __init__(self, outer_box, box_size, box_overlap, back_shift_small_boxes=True):
      self._box = outer_box
      self.num_boxes = 1

    def generate(self, i):
      return (0, 0, 0), self._box

    def index_to_cropped_box(self, i):
      return self._box

  for name in ('cv2', 'skimage', 'skimage.exposure', 'connectomics.segmentation',
               'connectomics.segmentation.labels', 'connectomics.common.box_generator'):
    sys.modules.setdefault(name, types.ModuleType(name))
  sys.modules['skimage'].exposure = sys.modules['skimage.exposure']
  sys.modules['connectomics.common.box_generator'].BoxGenerator = BoxGenerator
  sys.modules['connectomics.segmentation'].labels = sys.modules['connectomics.segmentation.labels']
  import connectomics
  connectomics.common.box_generator = sys.modules['connectomics.common.box_generator']
  connectomics.segmentation = sys.modules['connectomics.segmentation']
  return shim.load_reference('warp'), compat


def main():
  warp, compat = load_warp()
  rng = np.random.default_rng(17)
  out = {}

  def smooth(shape, sig, amp):
    return ndi.gaussian_filter(rng.standard_normal(shape), sig) * amp

  # 2-d: uint8 and float32 images, float64 and float32 maps, order 0 / 1
  img = (ndi.gaussian_filter(rng.random((150, 170)), 1.0) * 255).astype(np.uint8)
  cmap = np.stack([smooth((31, 35), 3, 40), smooth((31, 35), 3, 40)])
  cmap[:, 4, 5] += 7.3
  out['w2_image'], out['w2_map'] = img, cmap
  for order in (0, 1):
    out[f'w2_u8_o{order}'] = warp.ndimage_warp(img, cmap, (5, 5), (64, 64), (0, 0), order=order)
  imgf = img.astype(np.float32) / np.float32(7)
  out['w2_f32_o1'] = warp.ndimage_warp(imgf, cmap.astype(np.float32), (5, 5), (64, 64), (0, 0))
  # 3-d
  vol = (ndi.gaussian_filter(rng.random((14, 60, 70)), 1.0) * 60000).astype(np.uint16)
  cmap3 = np.stack([smooth((7, 15, 14), 2, 9), smooth((7, 15, 14), 2, 9), smooth((7, 15, 14), 2, 3)])
  out['w3_image'], out['w3_map'] = vol, cmap3
  for order in (0, 1):
    out[f'w3_u16_o{order}'] = warp.ndimage_warp(vol, cmap3, (2, 4, 5), (32, 32, 8), (2, 2, 2),
                                                order=order)
  # boxes (3-d only in the reference: the offset is reshaped to [dim, 1, 1, 1]): map with
  # context around the output, output smaller than the image, out_scale
  image_box = compat.BoundingBox(start=(20, 30, 2), size=(70, 60, 14))
  map_box = compat.BoundingBox(start=(3, 5, 1), size=(14, 15, 7))
  out_box = compat.BoundingBox(start=(25, 38, 3), size=(50, 40, 10))
  out['w3_boxes'] = warp.ndimage_warp(vol, cmap3, (2, 4, 5), (32, 32, 8), (2, 2, 2),
                                      image_box=image_box, map_box=map_box, out_box=out_box)
  out['w3_scale'] = warp.ndimage_warp(vol, cmap3, (2, 4, 5), (32, 32, 8), (2, 2, 2),
                                      image_box=image_box, map_box=map_box, out_box=out_box,
                                      out_scale=(0.5, 0.5, 1.0))
  np.savez_compressed(os.path.join(HERE, 'warp_golden.npz'), **out)
  print('warp_golden.npz', os.path.getsize(os.path.join(HERE, 'warp_golden.npz')) // 1024, 'KiB')


if __name__ == '__main__':
  main()
