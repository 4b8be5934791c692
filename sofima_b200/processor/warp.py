"""Warping processors (reference processor/warp.py).

`StitchAndRender3dTiles` (processor/warp.py:38-343) renders a grid of 3-d tiles through their
solved meshes with distance-weighted blending (`warp.ndimage_warp` on the GPU; the blend
accumulators stay on the device).  `WarpByMap` (processor/warp.py:346-538) renders a subvolume through an inverse coordinate
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
from .. import mesh as _mesh_lib
from .. import warp
from ..compat import config as cfg_lib

BoundingBox = compat.BoundingBox
Subvolume = compat.Subvolume


def _border_distance(mask: np.ndarray) -> np.ndarray:
  """Euclidean distance of every True pixel to the nearest False pixel, with everything
  outside the array counting as False -- `edt.edt(mask, black_border=True)` of the un-vendored
  `edt` package the reference calls (processor/warp.py:165), float32 like it."""
  from scipy import ndimage  # pylint: disable=g-import-not-at-top
  padded = np.pad(mask.astype(bool), 1)
  return ndimage.distance_transform_edt(padded)[1:-1, 1:-1].astype(np.float32)


class StitchAndRender3dTiles(compat.SubvolumeProcessor):
  """Renders a volume by warping 3-d tiles through their meshes and blending them
  (processor/warp.py:38-343).

  The tile meshes are shared by all instances, like the reference's class-level cache;
  `reset_cache()` drops them.  `tile_mesh_path` is an .npz path or any mapping with
  'key_to_idx' ({(x, y): column of 'x'}) and 'x' ([3, n_tiles, mz, my, mx] relative meshes).
  """

  _tile_meshes = None
  _tile_idx_to_xy = None
  _tile_boxes: dict = {}
  _inverted_meshes: dict = {}

  crop_at_borders = False

  def __init__(self, *, tile_map, tile_mesh_path, tile_pattern_path, stride, offset=(0, 0, 0),
               margin: int = 0, work_size=(128, 128, 128), order: int = 1,
               parallelism: int = 16, input_volinfo=None):
    del input_volinfo
    self._tile_map = np.array(tile_map)
    self._tile_mesh_path = tile_mesh_path
    self._tile_pattern_path = tile_pattern_path
    self._stride = stride  # zyx
    self._offset = offset  # xyz
    self._margin = margin
    self._order = order
    self._parallelism = parallelism
    self._work_size = work_size
    self._key_to_idx = {(x, y): tile_id for y, row in enumerate(tile_map)
                        for x, tile_id in enumerate(row)}

  @classmethod
  def reset_cache(cls):
    cls._tile_meshes = None
    cls._tile_idx_to_xy = None
    cls._tile_boxes = {}
    cls._inverted_meshes = {}

  def _open_tile_volume(self, tile_id: int):
    """Returns a ZYX-shaped array-like with the data of a tile."""
    raise NotImplementedError('This function needs to be defined in a subclass.')

  def context(self):
    return (0, 0, 0), (0, 0, 0)

  def _load_meshes(self) -> bool:
    cls = StitchAndRender3dTiles
    if cls._tile_meshes is not None:
      return False
    data = self._tile_mesh_path
    if isinstance(data, (str, bytes)) or hasattr(data, '__fspath__'):
      data = np.load(data, allow_pickle=True)
    key_to_idx = data['key_to_idx']
    if isinstance(key_to_idx, np.ndarray):
      key_to_idx = key_to_idx.item()
    cls._tile_idx_to_xy = {v: k for k, v in key_to_idx.items()}
    cls._tile_meshes = np.asarray(data['x'])
    assert cls._tile_meshes.shape[1] == len(cls._tile_idx_to_xy)
    return True

  def _collect_tile_boxes(self, tile_shape_zyx):
    """Global box every tile can render, and its box in mesh nodes
    (processor/warp.py:118-149)."""
    cls = StitchAndRender3dTiles
    meshes = cls._tile_meshes
    map_box = BoundingBox(start=(0, 0, 0), size=meshes.shape[2:][::-1])
    sz, sy, sx = self._stride
    for i in range(meshes.shape[1]):
      tx, ty = cls._tile_idx_to_xy[i]
      tg_box = map_utils.outer_box(meshes[:, i, ...], map_box, self._stride)
      out_box = BoundingBox(
          start=(tg_box.start[0] * sx + tx * tile_shape_zyx[-1] + self._offset[0],
                 tg_box.start[1] * sy + ty * tile_shape_zyx[-2] + self._offset[1],
                 tg_box.start[2] * sz + self._offset[2]),
          size=(tg_box.size[0] * sx, tg_box.size[1] * sy, tg_box.size[2] * sz))
      cls._tile_boxes[i] = out_box, tg_box

  def _get_dts(self, shape, tx: int, ty: int) -> np.ndarray:
    """Blending weight of a tile: distance from its usable area's edge
    (processor/warp.py:151-165).  `margin` pixels are ignored on inner edges only.  The
    usable area is a rectangle, so its Euclidean distance transform (everything outside the
    array counts as background) is the distance to the nearest side -- computed directly;
    `_border_distance` is the general form it is checked against."""
    h, w = int(shape[1]), int(shape[2])
    if self._margin > 0:
      x0 = self._margin if tx > 0 else 0
      x1 = -self._margin if tx < self._tile_map.shape[-1] - 1 else -1
      y0 = self._margin if ty > 0 else 0
      y1 = -self._margin if ty < self._tile_map.shape[-2] - 1 else -1
      ys, xs = slice(y0, y1).indices(h), slice(x0, x1).indices(w)
      y0, y1, x0, x1 = ys[0], ys[1], xs[0], xs[1]
    else:
      y0, y1, x0, x1 = 0, h, 0, w
    yy = np.arange(h, dtype=np.float32)[:, None]
    xx = np.arange(w, dtype=np.float32)[None, :]
    dist = np.minimum(np.minimum(yy - (y0 - 1), y1 - yy), np.minimum(xx - (x0 - 1), x1 - xx))
    return np.maximum(dist, 0).astype(np.float32)

  def _dts_on_device(self, shape, tx: int, ty: int):
    """`_get_dts` of a tile, computed once per instance and kept on the device."""
    import torch  # pylint: disable=g-import-not-at-top
    cache = self.__dict__.setdefault('_dts_cache', {})
    key = (tuple(shape[1:]), tx, ty)
    if key not in cache:
      cache[key] = torch.from_numpy(self._get_dts(shape, tx, ty)).cuda()
    return cache[key]

  def _tile_work(self, box: BoundingBox, tile_shape_zyx):
    """Work items of the tiles that reach into `box` (processor/warp.py:167-255): inverse
    mesh (cached), the part of the box the tile covers, the tile data needed for it."""
    cls = StitchAndRender3dTiles
    meshes = cls._tile_meshes
    image_box = BoundingBox(start=(0, 0, 0), size=tile_shape_zyx[::-1])
    map_box = BoundingBox(start=(0, 0, 0), size=meshes.shape[2:][::-1])
    for i, (out_box, tg_box) in cls._tile_boxes.items():
      sub_box = out_box.intersection(box)
      if sub_box is None:
        continue
      tx, ty = cls._tile_idx_to_xy[i]
      if i not in cls._inverted_meshes:
        # one node of context against rounding; holes can only be outside the hull
        tg_box = tg_box.adjusted_by(start=(-1, -1, -1), end=(1, 1, 1))
        inverse = map_utils.invert_map(meshes[:, i, ...], map_box, tg_box, stride=self._stride)
        inverse = map_utils.fill_missing(inverse, extrapolate=True, interpolate_first=False)
        cls._inverted_meshes[i] = tg_box, inverse
      else:
        tg_box, inverse = cls._inverted_meshes[i]
      # frame with the source tile at the origin
      local_out_box = out_box.translate((-tx * tile_shape_zyx[-1] - self._offset[0],
                                         -ty * tile_shape_zyx[-2] - self._offset[1],
                                         -self._offset[2]))
      local_warp_box = sub_box.translate(-out_box.start).translate(local_out_box.start)
      s = 1.0 / np.array(self._stride)[::-1]
      local_map_box = local_warp_box.scale(s).adjusted_by(start=(-2, -2, -2), end=(2, 2, 2))
      local_map_box = local_map_box.intersection(tg_box)
      if local_map_box is None:
        continue
      query = local_map_box.translate(-tg_box.start)
      assert np.all(query.start >= 0)
      sub_map = inverse[query.to_slice4d()]
      data_box = map_utils.outer_box(sub_map, local_map_box, self._stride, 1)
      data_box = data_box.intersection(image_box)
      if data_box is None:
        continue
      # the 2-d weight repeated over the sections of the data box (on the device)
      dts = self._dts_on_device(tile_shape_zyx, tx, ty)
      sub_dts = dts[data_box.to_slice3d()[1:]][None, ...].expand(
          int(data_box.size[2]), -1, -1).contiguous()
      yield i, inverse, tg_box, local_warp_box, sub_box, sub_dts, data_box

  def process(self, subvol: Subvolume):
    """Distance-weighted average of all tiles that reach into `subvol.bbox`
    (processor/warp.py:258-343)."""
    import torch  # pylint: disable=g-import-not-at-top
    box = subvol.bbox
    cls = StitchAndRender3dTiles
    mesh_init = self._load_meshes()
    volstores = {}
    for i in range(cls._tile_meshes.shape[1]):
      volstores[i] = self._open_tile_volume(self._key_to_idx[cls._tile_idx_to_xy[i]])
    tile_shape_zyx = tuple(next(iter(volstores.values())).shape)
    if mesh_init or not cls._tile_boxes:
      self._collect_tile_boxes(tile_shape_zyx)
    # blend accumulators in float32 on the device
    shape = tuple(int(v) for v in subvol.data.shape[1:])
    img = torch.zeros(shape, dtype=torch.float32, device='cuda')
    norm = torch.zeros(shape, dtype=torch.float32, device='cuda')
    for i, inverse, tg_box, local_warp_box, sub_box, sub_dts, data_box in self._tile_work(
        box, tile_shape_zyx):
      image = np.asarray(volstores[i][data_box.to_slice3d()])
      kwargs = dict(work_size=self._work_size, overlap=(0, 0, 0), image_box=data_box,
                    map_box=tg_box, out_box=local_warp_box, parallelism=self._parallelism)
      if image.dtype in (np.uint8, np.float32):
        image = torch.from_numpy(np.ascontiguousarray(image)).cuda()
      warped = warp.ndimage_warp(image, inverse, self._stride, order=self._order, **kwargs)
      if not torch.is_tensor(warped):
        warped = torch.from_numpy(warped.astype(np.float32)).cuda()
      weight = warp.ndimage_warp(sub_dts, inverse, self._stride, **kwargs)
      sel = sub_box.translate(-box.start).to_slice3d()
      img[sel] += warped * weight
      norm[sel] += weight
    filled = norm > 0
    img[filled] /= norm[filled]
    out_dtype = np.dtype(self.output_type(subvol.data.dtype))
    if out_dtype == np.uint8:  # same truncation as astype, 4x less to copy back
      out = _mesh_lib._to_host(img.clamp_(0, 255).to(torch.uint8))
    else:
      out = img.cpu().numpy().astype(out_dtype)
    return self.crop_box_and_data(box, out[None, ...])


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
