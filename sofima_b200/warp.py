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
  warped = out_d.cpu().numpy().view(np_dtype).reshape(out_shape)
  if labels_back is not None:
    warped = labels_back[warped]
  return warped


def warp_subvolume(image, image_box, coord_map, map_box, stride, out_box, interpolation=None,
                   offset: float = 0.0, parallelism: int = 1):
  """Warps a [n, z, y, x] subvolume through an xy inverse coordinate map (warp.py:58-186).

  The reference evaluates this with OpenCV's fixed-point `remap` (Lanczos / linear /
  nearest on CV_16SC2 maps); a CUDA restatement of those interpolation tables is not part of
  this backend yet, so the call fails loudly instead of falling back to the CPU.  Use
  `ndimage_warp` (bit-exact against SciPy) for map-based rendering on the device.
  """
  del image, image_box, coord_map, map_box, stride, out_box, interpolation, offset, parallelism
  raise NotImplementedError(
      'warp_subvolume (OpenCV remap semantics) is not built in the CUDA backend; '
      'ndimage_warp renders through the same kind of map on the device')
