"""Drop-in for `sofima.warp.ndimage_warp` (reference warp.py:189-335) on the B200.

The coordinate-map preparation (relative -> absolute, box offsets, `out_scale`) is the
reference's NumPy arithmetic on the small map array; the per-voxel work -- interpolate
the map, sample the image -- runs in one CUDA kernel that follows
`scipy.ndimage.map_coordinates` operation by operation in float64, so the output equals
the reference's bit for bit.  `work_size`, `overlap` and `parallelism` only bound host
memory / threads in the reference and do not change any output value; they are accepted
and ignored.  Interpolation orders 0 and 1 (the reference's default) are built.
"""

from __future__ import annotations

import ctypes
from typing import Sequence

import numpy as np
from scipy import ndimage

from . import _native
from . import compat
from . import mesh as _mesh

_DTYPES = {np.dtype(np.uint8): 0, np.dtype(np.float32): 1, np.dtype(np.uint16): 2,
           np.dtype(np.uint32): 3}


def _to_absolute(coord_map: np.ndarray, stride) -> np.ndarray:
  """map_utils.to_absolute (map_utils.py:150-185) without a box."""
  coord_map = coord_map.copy()
  dim = coord_map.shape[0]
  grids = np.mgrid[tuple(slice(0, s) for s in coord_map.shape[-dim:])]
  for i in range(dim):
    coord_map[i, ...] += grids[dim - 1 - i] * stride[dim - 1 - i]
  return coord_map


def ndimage_warp(image, coord_map: np.ndarray, stride: Sequence[float],
                 work_size: Sequence[int] = (), overlap: Sequence[int] = (), order=1,
                 map_coordinates=ndimage.map_coordinates, image_box=None, map_box=None,
                 out_box=None, parallelism: int = 1, out_scale=(1.0, 1.0, 1.0)):
  """Warps `image` through a coordinate map.

  Args:
    image: [z, ] y, x data to warp (uint8, uint16, uint32, float32; uint64 label
      volumes are warped with order 0 like in the reference)
    coord_map: [N, [z,] y, x] coordinate map in relative format
    stride: [z,] y, x image pixels per coordinate-map pixel
    work_size, overlap, parallelism: accepted for compatibility (host tiling knobs)
    order: interpolation order, 0 or 1
    map_coordinates: must be scipy.ndimage.map_coordinates (the routine the kernel
      reproduces)
    image_box: bounding box of the image data (XYZ)
    map_box: bounding box of the coordinate map; needs image_box
    out_box: bounding box of the output; defaults to the image box
    out_scale: xy[z] out_image_voxel_size / source_image_voxel_size

  Returns:
    warped image (NumPy in -> NumPy out; a CUDA tensor image stays on the device)
  """
  shape = coord_map.shape[1:]
  dim = len(shape)
  assert dim == len(stride)
  if work_size and overlap:
    assert dim == len(overlap) == len(work_size)
  if dim != image.ndim:
    raise ValueError(f'Dimension mismatch: image: {image.ndim} vs coord map: {dim}')
  if map_coordinates is not ndimage.map_coordinates:
    raise NotImplementedError(
        'The CUDA kernel reproduces scipy.ndimage.map_coordinates; other callables '
        'cannot be traced into it.')
  if order not in (0, 1):
    raise NotImplementedError(f'interpolation order {order}: only 0 and 1 are built')

  labels_back = None
  if not _mesh._is_tensor(image) and image.dtype == np.uint64:
    # Label volumes: contiguous ids, nearest neighbour (warp.py:240-243).
    # Id 0 (background) is pinned at index 0 whether or not it occurs in the volume, as
    # labels.make_contiguous does upstream: the kernel writes cval = 0 for samples outside
    # the image, and that has to read back as label 0, not as the smallest real label.
    ids = np.unique(np.append(image.ravel(), np.uint64(0)))
    if len(ids) >= 2**32:
      raise ValueError('too many distinct labels')
    image = np.searchsorted(ids, image).astype(np.uint32)
    labels_back, order = ids, 0

  src_map = _to_absolute(np.asarray(coord_map), stride)
  if map_box is not None:
    if image_box is None:
      raise ValueError('image_box has to be specified when map_box is used.')
    # (3-d only in the reference: the offset is shaped [dim, 1, 1, 1])
    src_map += (map_box.start[:dim] * np.asarray(stride)[::-1]
                - image_box.start[:dim] / np.asarray(out_scale)[:dim]).reshape(dim, 1, 1, 1)
  reshaper = tuple([slice(None)] + [np.newaxis] * dim)
  src_map = src_map.copy() * np.array(out_scale[:dim])[reshaper]

  if out_box is not None:
    out_shape = tuple(int(v) for v in out_box.size[::-1])[-dim:]
  else:
    out_shape = tuple(image.shape)
    size_xyz = list(image.shape[::-1]) + ([1] if dim == 2 else [])
    out_box = compat.BoundingBox(start=(0, 0, 0), size=size_xyz)
  if map_box is not None:
    offset = (map_box.start * np.asarray(stride)[::-1] - out_box.start)[::-1]
  else:
    offset = (0,) * dim

  dev = image.device.index if _mesh._is_tensor(image) and image.is_cuda else None
  ctx = _native.Context.get(dev)
  torch = _mesh._torch()
  device = torch.device('cuda', ctx.device)
  if _mesh._is_tensor(image):
    img_d = image.to(device).contiguous()
    np_dtype = np.dtype(str(img_d.dtype).replace('torch.', ''))
  else:
    host = np.ascontiguousarray(image)
    np_dtype = host.dtype
    if np_dtype not in _DTYPES:
      raise NotImplementedError(f'image dtype {np_dtype} is not supported by the CUDA warp')
    # torch has no uint16 / uint32 arithmetic, but can carry the bytes
    img_d = torch.from_numpy(host.view(np.uint8)).to(device)
  if np_dtype not in _DTYPES:
    raise NotImplementedError(f'image dtype {np_dtype} is not supported by the CUDA warp')
  map_d = torch.from_numpy(np.ascontiguousarray(src_map, dtype=np.float64)).to(device)
  out_d = torch.empty(int(np.prod(out_shape)) * np_dtype.itemsize, dtype=torch.uint8,
                      device=device)
  i64 = ctypes.c_int64 * dim
  f64 = ctypes.c_double * dim
  ctx.bind_stream()
  rc = _native.lib().sofima_warp_image(
      ctx.handle, dim, img_d.data_ptr(), _DTYPES[np_dtype], i64(*image.shape),
      map_d.data_ptr(), i64(*shape), f64(*[float(v) for v in offset]),
      f64(*[float(v) for v in stride]), int(order), out_d.data_ptr(), i64(*out_shape))
  _native.check(ctx.handle, rc)
  if _mesh._is_tensor(image):
    return out_d.view(img_d.dtype).reshape(out_shape)
  warped = _mesh._to_host(out_d).view(np_dtype).reshape(out_shape)
  if labels_back is not None:
    warped = labels_back[warped]
  return warped


_INTERPOLATION = {'nearest': 0, 'linear': 1, 'cubic': 2, 'lanczos': 3}
_CV_FLAGS = {0: 0, 1: 1, 2: 2, 4: 3}  # cv2.INTER_NEAREST / _LINEAR / _CUBIC / _LANCZOS4
_SECTION_DTYPES = {np.dtype(np.uint8): 0, np.dtype(np.float32): 1, np.dtype(np.uint16): 2,
                   np.dtype(np.int32): 3, np.dtype(np.uint32): 3, np.dtype(np.int16): 4}


def _interpolation_code(interpolation) -> int:
  """warp.py:33-40 (names) and the cv2 flag values the reference also accepts."""
  if isinstance(interpolation, str):
    return _INTERPOLATION[interpolation]  # KeyError for unknown names, like the reference
  try:
    return _CV_FLAGS[int(interpolation)]
  except (KeyError, TypeError, ValueError):
    raise ValueError(f'unknown interpolation {interpolation!r}') from None


def warp_subvolume(image, image_box, coord_map, map_box, stride, out_box, interpolation=None,
                   offset: float = 0.0, parallelism: int = 1):
  """Warps a [n, z, y, x] subvolume through an xy inverse coordinate map (warp.py:58-186).

  Args:
    image: [n, z, y, x] data to warp (uint8, uint16, int16, float32; uint32 below 2**16 is
      warped as uint16; uint64 is treated as segmentation and sampled nearest-neighbour on
      contiguous ids); NumPy array or CUDA tensor
    image_box: bounding box of `image` within the volume
    coord_map: [2, z, y, x] relative xy coordinate map (source position of every node)
    map_box: bounding box of `coord_map`
    stride: image pixels per coordinate-map pixel
    out_box: bounding box of the warped output
    interpolation: 'nearest' | 'linear' | 'cubic' | 'lanczos' (or the cv2 flag); defaults to
      Lanczos for images
    offset: (deprecated upstream) shift applied to the map and its node positions
    parallelism: accepted for compatibility (host threads in the reference)

  Returns:
    warped image covering `out_box`, [n, z, out_y, out_x]; values equal the reference's
    scipy + OpenCV pipeline bit for bit (one CUDA kernel, csrc/warp_cv.cu).
  """
  del parallelism
  torch = _mesh._torch()
  is_tensor = _mesh._is_tensor(image)
  if image.ndim != 4:
    raise ValueError(f'expected an [n, z, y, x] image, got shape {tuple(image.shape)}')
  labels_back = None
  orig_dtype = None
  if not is_tensor:
    image = np.asarray(image)
    orig_dtype = image.dtype
    if image.dtype == np.uint64:  # warp.py:93-100
      # Id 0 stays at index 0 whether or not it occurs, as labels.make_contiguous does.
      ids = np.unique(np.append(image.ravel(), np.uint64(0)))
      assert len(ids) < 2**31
      image = np.searchsorted(ids, image).astype(np.int32)
      labels_back, interpolation = ids, 'nearest'
    elif image.dtype == np.uint32:  # warp.py:109-115
      if image.size and image.max() >= 2**16:
        raise ValueError('Image warping supported up to uint16 only. For segmentation '
                         'data, use uint64.')
      image = image.astype(np.uint16)
    np_dtype = image.dtype
  else:
    np_dtype = np.dtype(str(image.dtype).replace('torch.', ''))
    if np_dtype == np.uint32 or np_dtype == np.uint64:
      raise NotImplementedError('pass uint32 / uint64 volumes as NumPy arrays (they are '
                                'converted on the host like in the reference)')
  code = 3 if interpolation is None else _interpolation_code(interpolation)
  if np_dtype not in _SECTION_DTYPES:
    raise NotImplementedError(f'image dtype {np_dtype} is not supported by the CUDA warp')
  if _SECTION_DTYPES[np_dtype] == 3 and code != 0:
    raise NotImplementedError('32-bit integer data is only warped nearest-neighbour '
                              '(OpenCV has no other mode for it either)')

  coord_map = np.asarray(coord_map)
  if coord_map.ndim != 4 or coord_map.shape[0] != 2 or coord_map.shape[1] != image.shape[1]:
    raise ValueError(f'coordinate map {coord_map.shape} does not match image '
                     f'{tuple(image.shape)}')
  if not np.issubdtype(coord_map.dtype, np.floating):
    coord_map = coord_map.astype(np.float64)
  if coord_map.shape[2] < 2 or coord_map.shape[3] < 2:
    raise ValueError('The points in dimension 0 must have at least 2 points')  # as scipy
  skipped = np.all(np.isnan(coord_map), axis=(0, 2, 3))  # warp.py:117-119
  # Map nodes -> coordinates within `image` (warp.py:123-126), in the map's own dtype.
  abs_map = _to_absolute(coord_map, (float(stride), float(stride)))
  mstart = np.asarray(map_box.start)
  abs_map += (mstart[:2] * stride - np.asarray(image_box.start)[:2] + offset).reshape(2, 1, 1, 1)
  # Node positions within the output (warp.py:130-134).
  ostart = np.asarray(out_box.start)
  grid_y = (np.arange(coord_map.shape[2]) + mstart[1]) * stride - ostart[1] + offset
  grid_x = (np.arange(coord_map.shape[3]) + mstart[0]) * stride - ostart[0] + offset
  out_shape = (int(image.shape[0]), int(out_box.size[2]), int(out_box.size[1]),
               int(out_box.size[0]))
  if out_shape[1] != image.shape[1]:
    raise ValueError(f'out_box has {out_shape[1]} sections, the image {image.shape[1]}')

  dev = image.device.index if is_tensor and image.is_cuda else None
  ctx = _native.Context.get(dev)
  device = torch.device('cuda', ctx.device)
  is_f64 = int(abs_map.dtype == np.float64)
  grids = [np.asarray(grid_y, np.float64), np.asarray(grid_x, np.float64)]
  my, mx = int(coord_map.shape[2]), int(coord_map.shape[3])

  def launch(img_d, z0, z1, out_d):
    """Sections [z0, z1) of a device-resident [n, z1 - z0, y, x] block -> out_d."""
    nzc = z1 - z0
    host = np.concatenate([np.ascontiguousarray(abs_map[:, z0:z1], dtype=np.float64).ravel()]
                          + grids)
    map_d = torch.from_numpy(host).to(device, non_blocking=True)
    n_map = 2 * nzc * my * mx
    skip_d = None
    if skipped[z0:z1].any():
      skip_d = torch.from_numpy(skipped[z0:z1].astype(np.uint8)).to(device)
    ctx.bind_stream()
    rc = _native.lib().sofima_warp_subvolume(
        ctx.handle, img_d.data_ptr(), _SECTION_DTYPES[np_dtype],
        (ctypes.c_int64 * 4)(int(image.shape[0]), nzc, int(image.shape[2]),
                             int(image.shape[3])),
        map_d.data_ptr(), is_f64, map_d.data_ptr() + 8 * n_map,
        map_d.data_ptr() + 8 * (n_map + my), my, mx,
        None if skip_d is None else skip_d.data_ptr(), code, out_d.data_ptr(), out_shape[2],
        out_shape[3])
    _native.check(ctx.handle, rc)
    return map_d, skip_d  # referenced until the stream has consumed them

  if is_tensor:
    img_d = image.to(device).contiguous()
    out_d = torch.empty(int(np.prod(out_shape)) * np_dtype.itemsize, dtype=torch.uint8,
                        device=device)
    launch(img_d, 0, out_shape[1], out_d)
    return out_d.view(img_d.dtype).reshape(out_shape)

  # Host arrays: blocks of sections travel through two pairs of pinned buffers, so that the
  # host copies of one block overlap the transfers and the kernel of the previous one.
  n, nz = out_shape[0], out_shape[1]
  out_bytes_total = int(np.prod(out_shape)) * np_dtype.itemsize
  # The result lives in pinned memory from torch's caching host allocator when it is not
  # huge: the device writes every block straight into it (no bounce copy, no page faults on
  # fresh pageable memory).  The array keeps the block alive until the caller drops it.
  direct = 0 < out_bytes_total <= (1 << 30)
  if direct:
    warped_t = torch.empty(out_bytes_total, dtype=torch.uint8, pin_memory=True)
    warped = warped_t.numpy().view(np_dtype).reshape(out_shape)
  else:
    warped_t = None
    warped = np.empty(out_shape, np_dtype)
  in_sec = n * image.shape[2] * image.shape[3] * np_dtype.itemsize
  out_sec = n * out_shape[2] * out_shape[3] * np_dtype.itemsize
  zc = max(1, min(nz, _BLOCK_BYTES // max(in_sec, out_sec, 1)))
  pins = _pinned_blocks(zc * in_sec, 0 if direct else zc * out_sec)
  chan_out = out_shape[2] * out_shape[3] * np_dtype.itemsize  # bytes of one section, 1 channel
  pending = None  # (z0, z1, pinned out view, event, keep-alive)
  for k, z0 in enumerate(range(0, nz, zc)):
    z1 = min(nz, z0 + zc)
    pin_in, pin_out, done = pins[k % 2]
    if done[0] is not None:
      done[0].synchronize()  # the block that used this pair two rounds ago has left it
    src = pin_in[:(z1 - z0) * in_sec].view(torch.uint8)
    np.copyto(src.numpy().view(np_dtype).reshape((n, z1 - z0) + tuple(image.shape[2:])),
              image[:, z0:z1])
    img_d = src.to(device, non_blocking=True)
    out_d = torch.empty((z1 - z0) * out_sec, dtype=torch.uint8, device=device)
    keep = launch(img_d, z0, z1, out_d)
    if direct:
      for c in range(n):  # [c, z0:z1] is contiguous in the result
        o = (c * nz + z0) * chan_out
        warped_t[o:o + (z1 - z0) * chan_out].copy_(
            out_d[c * (z1 - z0) * chan_out:(c + 1) * (z1 - z0) * chan_out], non_blocking=True)
      dst = None
    else:
      dst = pin_out[:(z1 - z0) * out_sec]
      dst.copy_(out_d, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    done[0] = ev
    if pending is not None and not direct:
      _collect(warped, pending, np_dtype)
    pending = (z0, z1, dst, ev, (img_d, out_d, keep))
  if pending is not None:
    if direct:
      pending[3].synchronize()  # the last block (stream order: all earlier ones too)
    else:
      _collect(warped, pending, np_dtype)
  if labels_back is not None:
    return labels_back[warped]
  return warped.astype(orig_dtype, copy=False)


_BLOCK_BYTES = 48 << 20
_PINNED: dict = {}


def _pinned_blocks(in_bytes: int, out_bytes: int):
  """Two (pinned in, pinned out, [event]) triples per thread, grown on demand."""
  import threading
  torch = _mesh._torch()
  key = threading.get_ident()
  cur = _PINNED.get(key)
  if cur is None or cur[0][0].numel() < in_bytes or cur[0][1].numel() < out_bytes:
    if cur is not None:  # keep the larger of the old and new sizes of either buffer
      in_bytes = max(in_bytes, cur[0][0].numel())
      out_bytes = max(out_bytes, cur[0][1].numel())
    cur = [(torch.empty(in_bytes, dtype=torch.uint8, pin_memory=True),
            torch.empty(out_bytes, dtype=torch.uint8, pin_memory=True), [None])
           for _ in range(2)]
    _PINNED[key] = cur
  return cur


def _collect(warped: np.ndarray, pending, np_dtype):
  z0, z1, dst, ev, _ = pending
  ev.synchronize()
  n = warped.shape[0]
  np.copyto(warped[:, z0:z1],
            dst.numpy().view(np_dtype).reshape((n, z1 - z0) + warped.shape[2:]))


def render_tiles(tiles, coord_maps, stride=(20, 20), margin: int = 50, parallelism: int = 1,
                 width=None, height=None, use_clahe: bool = False, clahe_kwargs=None,
                 margin_overrides=None, return_warped_tiles: bool = False, tile_masks=None):
  """Warps a grid of tiles into one canvas (warp.py:338-535).

  Args:
    tiles: (x, y) tile coordinate -> [y, x] image; all tiles have the same shape
    coord_maps: (x, y) -> [2, 1, my, mx] forward coordinate map of the tile
    stride: map stride in pixels (equal in x and y)
    margin: pixels at the tile edges that are not rendered
    parallelism: accepted for compatibility (host threads in the reference; the tiles
      are warped one after the other on the GPU)
    width, height: canvas size; inferred from the tile grid when missing
    use_clahe, clahe_kwargs: CLAHE (skimage.exposure.equalize_adapthist) before warping
    margin_overrides: (x, y) -> (top, bottom, left, right) margins
    return_warped_tiles: also return {(x, y): (x0, y0, warped tile)}
    tile_masks: (x, y) -> array like the tile; only non-zero pixels are rendered

  Returns:
    (canvas [height, width], bool array of the pixels covered by tile content)
    [+ the dict of warped tiles].
  The map inversion runs on the map nodes with SciPy exactly as in the reference
  (map_utils.invert_map / fill_missing); the per-pixel warp is `warp_subvolume` on the GPU.
  """
  del parallelism
  from . import map_utils  # pylint: disable=g-import-not-at-top
  if stride[0] != stride[1]:
    raise NotImplementedError('Currently only equal strides in XY are supported.')
  first = next(iter(tiles.values()))
  ty, tx = first.shape
  image_box = compat.BoundingBox(start=(0, 0, 0), size=(tx, ty, 1))
  my, mx = next(iter(coord_maps.values())).shape[-2:]
  map_box = compat.BoundingBox(start=(0, 0, 0), size=(mx, my, 1))
  if width is None or height is None:
    height = ty * (max(y for _, y in tiles) + 1)
    width = tx * (max(x for x, _ in tiles) + 1)
  canvas = np.zeros((height, width), dtype=first.dtype)
  covered = np.zeros((height, width), dtype=bool)
  warped_tiles = {}
  clahe_kwargs = clahe_kwargs or {}

  for (tile_x, tile_y), coord_map in coord_maps.items():
    img = tiles.get((tile_x, tile_y))
    if img is None:
      continue
    # inverse map on a node box covering everything the tile maps to (+1 node of context)
    tg_box = map_utils.outer_box(coord_map, map_box, stride[0])
    tg_box = tg_box.adjusted_by(start=(-1, -1, 0), end=(1, 1, 0))
    inverse = map_utils.invert_map(coord_map, map_box, tg_box, stride[0])
    inverse = map_utils.fill_missing(inverse, extrapolate=True)
    # 1 inside the margins (and the tile mask), warped together with the image
    keep = np.zeros_like(img)
    if margin_overrides is not None and (tile_x, tile_y) in margin_overrides:
      top, bottom, left, right = margin_overrides[tile_x, tile_y]
      keep[top:-(bottom + 1), left:-(right + 1)] = 1
    else:
      keep[margin:-(margin + 1), margin:-(margin + 1)] = 1
    if use_clahe:
      try:
        import skimage.exposure  # pylint: disable=g-import-not-at-top
      except ImportError as e:
        raise NotImplementedError('use_clahe needs scikit-image') from e
      img = (skimage.exposure.equalize_adapthist(img, **clahe_kwargs)
             * np.iinfo(img.dtype).max).astype(img.dtype)
    if tile_masks is not None and tile_masks.get((tile_x, tile_y)) is not None:
      keep[tile_masks[tile_x, tile_y] == 0] = 0
    out_box = compat.BoundingBox(
        start=((tg_box.start[0] + 1) * stride[1], (tg_box.start[1] + 1) * stride[0], 0),
        size=(tg_box.size[0] * stride[1], tg_box.size[1] * stride[0], 1))
    w_img, w_keep = warp_subvolume(np.stack([img, keep])[:, None], image_box, inverse, tg_box,
                                   stride[0], out_box=out_box)
    w_img, w_keep = w_img[0], w_keep[0].astype(bool)
    # canvas position relative to the nominal tile position; trim what sticks out
    y0 = ty * tile_y + int(out_box.start[1])
    x0 = tx * tile_x + int(out_box.start[0])
    if x0 < 0:
      w_img, w_keep, x0 = w_img[:, -x0:], w_keep[:, -x0:], 0
    if y0 < 0:
      w_img, w_keep, y0 = w_img[-y0:], w_keep[-y0:], 0
    dst = canvas[y0:y0 + w_img.shape[0], x0:x0 + w_img.shape[1]]
    w_img = w_img[:dst.shape[0], :dst.shape[1]]
    w_keep = w_keep[:dst.shape[0], :dst.shape[1]]
    if return_warped_tiles:
      warped_tiles[tile_x, tile_y] = x0, y0, w_img
    covered[y0:y0 + w_img.shape[0], x0:x0 + w_img.shape[1]][w_keep] = True
    w_keep = w_keep & (w_img > 0)  # never paint unrendered (zero) pixels
    dst[w_keep] = w_img[w_keep]
  if return_warped_tiles:
    return canvas, covered, warped_tiles
  return canvas, covered


def warp_points(points: np.ndarray, coord_map: np.ndarray, map_box, stride: float) -> np.ndarray:
  """Maps [n, 3] xyz points through an in-plane [2, z, y, x] coordinate map
  (warp.py:541-605).  Host bookkeeping on a handful of points: the map of every section that
  holds a point is interpolated linearly (extrapolating) at the points, as upstream; integer
  point arrays get rounded coordinates, z is unchanged."""
  import collections  # pylint: disable=g-import-not-at-top
  from scipy import interpolate  # pylint: disable=g-import-not-at-top
  from . import map_utils  # pylint: disable=g-import-not-at-top
  abs_map = map_utils.to_absolute(coord_map, stride)
  abs_map += np.array(map_box.start[:2] * stride).reshape((2, 1, 1, 1))
  by_z = collections.defaultdict(list)
  for i, pt in enumerate(points):
    by_z[pt[2]].append(i)
  points = np.array(points)
  assert points.ndim == 2 and points.shape[1] == 3
  assert coord_map.shape[0] == 2
  out = points.copy()
  node_y = (np.arange(coord_map.shape[2]) + map_box.start[1]) * stride
  node_x = (np.arange(coord_map.shape[3]) + map_box.start[0]) * stride
  for z, idx in by_z.items():
    z_rel = int(z - map_box.start[2])
    query = points[idx, 1], points[idx, 0]  # yx
    new = []
    for comp in (0, 1):
      dense = interpolate.RegularGridInterpolator((node_y, node_x), abs_map[comp, z_rel, ...],
                                                  bounds_error=False, fill_value=None)
      val = dense(query).astype(np.float32)
      if np.issubdtype(out.dtype, np.integer):
        val = np.round(val).astype(out.dtype)
      new.append(val)
    out[idx, 0], out[idx, 1] = new
  return out
