"""Coordinate-map processors of the reference's processor/maps.py.

`InvertMap`, `ResampleMap`, `MaskIrregularities` and `FillMissing` (processor/maps.py:332-498)
are thin subvolume wrappers around `map_utils` -- host geometry on the map nodes, as upstream.
`ReconcileCrossBlockMaps` -- the reconciliation of block-wise and cross-block maps
(processor/maps.py:35-330) -- is outside the scope of this backend:
only its `Config` is provided so that `pipeline.mesh_config.default_em_2d()` round-trips;
constructing the processor raises."""

from __future__ import annotations

import dataclasses
from typing import Any

import numpy as np

from .. import compat
from .. import map_utils
from ..compat import config as cfg_lib

Subvolume = compat.Subvolume


class ReconcileCrossBlockMaps(compat.SubvolumeProcessor):
  """processor/maps.py:35 (configuration only, see the module docstring)."""

  @dataclasses.dataclass(eq=True)
  class Config(cfg_lib.JsonMixin):
    """Same fields as the reference's Config (processor/maps.py:55-84)."""
    cross_block: Any
    cross_block_inv: Any
    last_inv: Any
    main_inv: Any
    z_map: dict[str, int]
    stride: int
    xy_overlap: int = 128
    backward: bool = False

  crop_at_borders = False

  def __init__(self, config: 'ReconcileCrossBlockMaps.Config', input_volinfo=None):
    del config, input_volinfo
    raise NotImplementedError(
        'ReconcileCrossBlockMaps runs on the reference\'s CPU map inversion '
        '(map_utils.invert_map); it is outside the scope of the CUDA backend.')


class InvertMap(compat.SubvolumeProcessor):
  """Inverts a coordinate map (processor/maps.py:332-397)."""

  @dataclasses.dataclass(eq=True)
  class Config(cfg_lib.JsonMixin):
    """stride: [z]yx stride of the map; crop_output: emit the inner box of the map (what the
    inversion can fill) instead of the input box; input_volume: the map volume."""
    stride: Any
    crop_output: bool = True
    input_volume: Any = None

  crop_at_borders = False

  def __init__(self, config: 'InvertMap.Config', input_path_or_metadata=None):
    source = input_path_or_metadata if input_path_or_metadata is not None \
        else config.input_volume
    if source is None:
      raise ValueError('No source volume specified.')
    self._config = config
    if hasattr(source, 'shape'):      # [C, Z, Y, X] array-like
      size_xyz = tuple(int(v) for v in source.shape[:0:-1])
    elif hasattr(source, 'volume_size'):  # metadata-like
      vs = source.volume_size
      size_xyz = (int(vs.x), int(vs.y), int(vs.z))
    else:
      raise NotImplementedError('volume metadata I/O is a hook: pass an array-like map volume')
    self._volume_bbox = compat.BoundingBox(start=(0, 0, 0), size=size_xyz)

  def process(self, subvol: Subvolume):
    config, box = self._config, subvol.bbox
    if np.all(np.isnan(subvol.data)):
      return []  # nothing to invert
    rel_map = subvol.data.astype(np.float64)
    if config.crop_output:
      dst_box = map_utils.inner_box(rel_map, box, config.stride)
      dst_box = dst_box.intersection(self._volume_bbox)
    else:
      dst_box = box
    if dst_box is None:
      return []
    return [Subvolume(map_utils.invert_map(rel_map, box, dst_box, config.stride), dst_box)]


class ResampleMap(compat.SubvolumeProcessor):
  """Resamples a coordinate map on a grid of another spacing (processor/maps.py:400-444)."""

  @dataclasses.dataclass(eq=True)
  class Config(cfg_lib.JsonMixin):
    stride: int
    out_stride: int
    scale: float = 1.0
    method: str = 'linear'

  crop_at_borders = False

  def __init__(self, config: 'ResampleMap.Config', input_volinfo_or_ts=None):
    del input_volinfo_or_ts
    self._config = config

  def pixelsize(self, psize):
    psize = np.array(psize, dtype=np.float32)
    psize[:2] *= self._config.out_stride / self._config.stride
    return psize

  def process(self, subvol: Subvolume):
    config, box = self._config, subvol.bbox
    if np.all(np.isnan(subvol.data)):
      return []
    rel_map = subvol.data.astype(np.float64) * config.scale
    ratio = config.stride / config.out_stride
    dst_box = self.crop_box(box).scale([ratio, ratio, 1.0])
    out = map_utils.resample_map(rel_map, box, dst_box, config.stride, config.out_stride,
                                 config.method)
    return [Subvolume(out, dst_box)]


class MaskIrregularities(compat.SubvolumeProcessor):
  """NaNs stretched / folded nodes of every section (processor/maps.py:447-472)."""

  crop_at_borders = False

  def __init__(self, stride, frac, input_volinfo=None):
    del input_volinfo
    self._stride = stride
    self._frac = frac

  def context(self):
    return (3, 3, 0), (3, 3, 0)  # covers the dilation inside mask_irregular

  def process(self, subvol: Subvolume):
    out = np.zeros_like(subvol.data)
    for z in range(subvol.data.shape[1]):
      section = subvol.data[:, z, ...].copy()
      map_utils.mask_irregular(section, self._stride, self._frac)
      out[:, z, ...] = section
    return self.crop_box_and_data(subvol.bbox, out)


class FillMissing(compat.SubvolumeProcessor):
  """Fills invalid nodes by inter- / extrapolation (processor/maps.py:475-498)."""

  @dataclasses.dataclass(eq=True)
  class Config(cfg_lib.JsonMixin):
    """Empty, required by the processing framework."""

  crop_at_borders = False

  def __init__(self, input_volinfo=None):
    del input_volinfo

  def process(self, subvol: Subvolume):
    mesh = subvol.data
    if not np.all(np.isnan(mesh)):
      mesh = map_utils.fill_missing(mesh, extrapolate=True)
    return self.crop_box_and_data(subvol.bbox, mesh)
