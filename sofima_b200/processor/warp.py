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
    return np.asarray(vol[(slice(None),) + box.to_slice3d()])

  def process(self, subvol: Subvolume) -> Subvolume:
    """Renders the data under `subvol.bbox` (processor/warp.py:441-538): reads the inverse
    coordinate map for the box, the source data its vectors point to, and warps."""
    box = subvol.bbox
    map_vol = self._open_volume(self._map_volinfo)
    data_vol = self._open_volume(self._data_volinfo)
    stride = self._source_stride
    # map nodes covering the output box (one node of context on every side)
    start = np.array([box.start[0] // stride - 1, box.start[1] // stride - 1, box.start[2]])
    end = np.array([-(-box.end[0] // stride) + 1, -(-box.end[1] // stride) + 1, box.end[2]])
    map_shape_xyz = np.array(map_vol.shape[:0:-1])
    start = np.maximum(start, 0).astype(int)
    end = np.minimum(end, map_shape_xyz).astype(int)
    map_box = BoundingBox(start=start, end=end)
    coord_map = self._read(map_vol, map_box).astype(np.float32) * np.float32(self._scale)
    if np.all(np.isnan(coord_map)):
      return Subvolume(np.zeros((data_vol.shape[0],) + tuple(box.size[::-1]),
                                dtype=data_vol.dtype), box)
    # source data reached by the map over the output box
    abs_x = coord_map[0] + (np.arange(map_box.start[0], map_box.end[0]) * stride)[None, None, :]
    abs_y = coord_map[1] + (np.arange(map_box.start[1], map_box.end[1]) * stride)[None, :, None]
    data_shape_xyz = np.array(data_vol.shape[:0:-1])
    lo = np.array([np.floor(np.nanmin(abs_x)) - 8, np.floor(np.nanmin(abs_y)) - 8, box.start[2]])
    hi = np.array([np.ceil(np.nanmax(abs_x)) + 9, np.ceil(np.nanmax(abs_y)) + 9, box.end[2]])
    lo = np.maximum(lo, 0).astype(int)
    hi = np.minimum(hi, data_shape_xyz).astype(int)
    image_box = BoundingBox(start=lo, end=hi)
    image = self._read(data_vol, image_box)
    if self._mask_configs is not None:
      image = np.where(self._build_mask(self._mask_configs, image_box), 0, image)
    warped = warp.warp_subvolume(image, image_box, coord_map, map_box, stride, box,
                                 self._interpolation, self._offset)
    return Subvolume(warped, box)
