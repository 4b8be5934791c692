"""Warping processors (reference processor/warp.py).

`WarpByMap` (processor/warp.py:346-538) renders a subvolume through an inverse coordinate
map with `warp.warp_subvolume`; its `Config` is the EM-2D pipeline's `warp_config`
(processor/defaults/em_2d.py:244-262).  Volume I/O goes through the same hooks as the other
processors of this package (`_open_volume`, `_build_mask`): the connectomics volume stack is
not installable here.
"""

from __future__ import annotations

import dataclasses
import json
from typing import Any

import numpy as np

from .. import compat
from .. import map_utils
from .. import warp
from ..compat import config as cfg_lib

BoundingBox = compat.BoundingBox
Subvolume = compat.Subvolume


class WarpByMap(compat.SubvolumeProcessor):
  """Warps volume data according to a coordinate map (processor/warp.py:346)."""

  crop_at_borders = False
  ignores_input_data = True

  @dataclasses.dataclass(eq=True)
  class Config(cfg_lib.JsonMixin):
    """Same fields as the reference's Config (processor/warp.py:366-400)."""
    stride: float
    map_volinfo: Any
    data_volinfo: Any
    map_decorator_specs: Any = None
    data_decorator_specs: Any = None
    map_scale: float = 1.0
    interpolation: str | None = None
    downsample: int = 1
    offset: float = 0.0
    mask_configs: Any = None
    source_cache_bytes: int = int(1e9)

  def __init__(self, config: 'WarpByMap.Config', input_volinfo=None):
    del input_volinfo
    self._config = config
    self._map_volinfo = config.map_volinfo
    self._data_volinfo = config.data_volinfo
    self._scale = config.map_scale
    self._interpolation = config.interpolation

    def _get_specs(specs):
      if specs is None:
        return []
      return json.loads(specs) if isinstance(specs, str) else specs

    self._data_decorator_specs = _get_specs(config.data_decorator_specs)
    self._map_decorator_specs = _get_specs(config.map_decorator_specs)
    self._downsample = np.array([config.downsample, config.downsample, 1])  # xyz
    self._target_stride = config.stride
    self._source_stride = config.stride * config.downsample
    self._offset = config.offset
    self._mask_configs = config.mask_configs

  # --- hooks (volume / mask I/O stays with the caller's storage layer) -------------
  def _open_volume(self, volinfo):
    """Returns an array-like [C, Z, Y, X] volume for `volinfo` (already opened arrays pass
    through)."""
    if hasattr(volinfo, 'shape'):
      return volinfo
    raise NotImplementedError('volume I/O is a hook: pass arrays or override _open_volume')

  def _build_mask(self, mask_configs, box):
    raise NotImplementedError('mask building is a hook: override _build_mask')

  def _read(self, vol, box: BoundingBox) -> np.ndarray:
    return np.asarray(vol[box.to_slice4d()])

  @staticmethod
  def _clip_box_to_volume(vol, box: BoundingBox):
    full = BoundingBox(start=(0, 0, 0), size=tuple(vol.shape[:0:-1]))
    return box.intersection(full)

  def _load_and_warp(self, data_box, data_vol, map_data, map_box, out_box):
    """Warped data for `out_box` (processor/warp.py:444-474); None when fully masked."""
    data = np.array(self._read(data_vol, data_box))
    if self._mask_configs is not None:
      mask = self._build_mask(self._mask_configs, data_box)
      for c in range(data.shape[0]):
        data[c, ...][mask] = 0
      if np.all(mask):
        return None
    return warp.warp_subvolume(data, data_box, map_data, map_box, self._source_stride, out_box,
                               self._interpolation, self._offset)

  def _get_map_for_box(self, box: BoundingBox):
    """Coordinate map nodes needed for an output box (processor/warp.py:476-503): the box in
    map pixels plus two nodes of context, clipped to the map volume."""
    s = 1.0 / self._target_stride
    map_box = box.scale([s, s, 1.0]).adjusted_by(start=(-2, -2, 0), end=(2, 2, 0))
    map_vol = self._open_volume(self._map_volinfo)
    map_box = self._clip_box_to_volume(map_vol, map_box)
    if map_box is None or np.any(map_box.size == 0):
      return None, None
    rel_map = self._read(map_vol, map_box).astype(np.float64) * self._scale
    if np.all(np.isnan(rel_map)):
      return None, None
    return map_box, rel_map

  def _generate_boxes_to_warp(self, data_vol, box: BoundingBox):
    """Work items (out box, data box, map, map box) for `box` (processor/warp.py:505-539).
    OpenCV's remap addresses sources with int16, and the kernel keeps that quantisation for
    parity, so sources of 2**15 pixels or more are split 2 x 2 like in the reference."""
    map_box, rel_map = self._get_map_for_box(box)
    if map_box is None or np.any(map_box.size == 0):
      return
    data_box = map_utils.outer_box(rel_map, map_box, self._source_stride, 1)
    data_box = self._clip_box_to_volume(data_vol, data_box)
    if data_box is None or np.any(data_box.size == 0):
      return
    if np.all(data_box.size < 2**15):
      yield box, data_box, rel_map, map_box
      return
    if np.any(box.size[:2] < self._target_stride * 3):
      return
    sub = np.array(list(-(-box.size[:2] // 2)) + [box.size[2]])
    sub = -(-sub // self._downsample) * self._downsample
    for x0 in range(int(box.start[0]), int(box.end[0]), int(sub[0])):
      for y0 in range(int(box.start[1]), int(box.end[1]), int(sub[1])):
        sub_box = BoundingBox(start=(x0, y0, box.start[2]), end=(
            min(x0 + sub[0], box.end[0]), min(y0 + sub[1], box.end[1]), box.end[2]))
        yield from self._generate_boxes_to_warp(data_vol, sub_box)

  def _downsample_area(self, warped_sec: np.ndarray, warp_box: BoundingBox, dtype):
    """XY area average over `downsample` x `downsample` blocks aligned to the global grid
    (processor/warp.py:588-612; the reference goes through an integral image of int64 /
    float64 values, i.e. exact block sums, divides by the block area and casts)."""
    d = int(self._downsample[0])
    if warped_sec.dtype in (np.uint8, np.uint32):
      wide = warped_sec.astype(np.int64)
    elif warped_sec.dtype == np.float32:
      wide = np.nan_to_num(warped_sec.astype(np.float64))
    else:
      raise NotImplementedError(f'Downsampling of {warped_sec.dtype} not supported.')
    start = -(-warp_box.start[:2] // d)  # first fully covered output pixel
    end = warp_box.end[:2] // d
    x0, y0 = start * d - warp_box.start[:2]
    nx, ny = (end - start)
    blk = wide[:, :, y0:y0 + ny * d, x0:x0 + nx * d]
    sums = blk.reshape(blk.shape[0], blk.shape[1], ny, d, nx, d).sum(axis=(3, 5))
    down_box = BoundingBox(start=(start[0], start[1], warp_box.start[2]),
                           size=(nx, ny, warp_box.size[2]))
    return down_box, (sums / float(d * d)).astype(dtype)

  def process(self, subvol: Subvolume):
    """Renders the data under `subvol.bbox` section by section (processor/warp.py:541-623):
    map for the section, source box reached by the map, warp, optional area downsampling."""
    box = subvol.bbox
    data_vol = self._open_volume(self._data_volinfo)
    n_chan = subvol.data.shape[0] if subvol.data is not None else data_vol.shape[0]
    dtype = subvol.data.dtype if subvol.data is not None else data_vol.dtype
    warped = np.zeros([n_chan] + box.size[::-1].tolist(), dtype=dtype)
    for z in range(warped.shape[1]):
      curr_box = BoundingBox(start=box.start + [0, 0, z], size=[box.size[0], box.size[1], 1])
      for out_box, data_box, map_data, map_box in self._generate_boxes_to_warp(data_vol,
                                                                               curr_box):
        warp_box = out_box.scale(self._downsample)
        warped_sec = self._load_and_warp(data_box, data_vol, map_data, map_box, warp_box)
        if warped_sec is None:
          continue
        if warp_box != out_box:
          down_box, down = self._downsample_area(warped_sec, warp_box, warped.dtype)
          warped[down_box.translate(-box.start).to_slice4d()] = down
        else:
          warped[out_box.translate(-box.start).to_slice4d()] = warped_sec
    return [self.crop_box_and_data(box, warped)]
