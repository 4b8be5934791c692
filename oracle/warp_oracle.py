"""NumPy float64 restatement of the image-warping path (TEST ORACLE).

Follows /root/reference/warp.py:189-335 (`ndimage_warp`) and the SciPy routine it calls
twice per output voxel, `scipy.ndimage.map_coordinates` (third-party: scipy, unpinned in
setup.cfg; scipy IS installed in this image, so `map_coordinates` below is checked
bit-for-bit against the real routine in tests/test_oracle_warp.py).  Restated algorithm
(order 0 / 1, mode 'constant', cval 0 -- the defaults ndimage_warp uses):
  * all arithmetic in float64;
  * a sample with ANY coordinate outside [0, n - 1] is exactly cval;
  * order 1: per axis i0 = floor(c), t = c - i0, weights (1 - t, t), upper index clamped
    to n - 1; corners visited with the last axis fastest; each term is
    ((value * w_axis0) * w_axis1) ...; terms added in visiting order;
  * order 0: index floor(c + 0.5);
  * unsigned outputs: t > 0 ? (T)(min(t + 0.5, max)) : 0; floats: plain cast.
The box tiling of ndimage_warp (work_size / overlap / parallelism) does not change any
output value and is not restated.

Test infrastructure only: never imported from `sofima_b200/`.
"""

from __future__ import annotations

import itertools

import numpy as np


def map_coordinates(arr, coords, order=1):
  """scipy.ndimage.map_coordinates(arr, coords, order=order) for order in (0, 1)."""
  arr = np.asarray(arr)
  coords = [np.asarray(c, np.float64) for c in coords]
  shape = coords[0].shape
  inside = np.ones(shape, bool)
  for c, n in zip(coords, arr.shape):
    inside &= (c >= 0) & (c <= n - 1)
  if order == 0:
    idx = [np.clip(np.floor(np.where(inside, c, 0) + 0.5).astype(np.int64), 0, n - 1)
           for c, n in zip(coords, arr.shape)]
    return np.where(inside, arr[tuple(idx)], 0).astype(arr.dtype)
  assert order == 1
  i0, w = [], []
  for c in coords:
    cc = np.where(inside, c, 0.0)
    f = np.floor(cc)
    t = cc - f
    i0.append(f.astype(np.int64))
    w.append((1.0 - t, t))
  acc = np.zeros(shape, np.float64)
  for corner in itertools.product((0, 1), repeat=arr.ndim):
    idx = [np.minimum(i0[d] + k, arr.shape[d] - 1) for d, k in enumerate(corner)]
    term = arr[tuple(idx)].astype(np.float64)
    for d, k in enumerate(corner):
      term = term * w[d][k]
    acc = acc + term
  acc = np.where(inside, acc, 0.0)
  if arr.dtype.kind == 'u':
    t = np.where(acc > 0, acc + 0.5, 0.0)
    return np.minimum(t, float(np.iinfo(arr.dtype).max)).astype(arr.dtype)
  return acc.astype(arr.dtype)


def source_map(coord_map, stride, image_start=None, map_start=None, out_scale=None):
  """Absolute source coordinates per map node, in image voxels (warp.py:245-260,
  map_utils.to_absolute map_utils.py:150-185).  Starts are xyz like BoundingBox.start."""
  coord_map = np.array(coord_map, copy=True)
  dim = coord_map.shape[0]
  grids = np.mgrid[tuple(slice(0, s) for s in coord_map.shape[-dim:])]
  for i in range(dim):
    coord_map[i, ...] += grids[dim - 1 - i] * stride[dim - 1 - i]
  out_scale = (1.0,) * dim if out_scale is None else out_scale
  if map_start is not None:
    coord_map += (np.asarray(map_start)[:dim] * np.asarray(stride)[::-1]
                  - np.asarray(image_start)[:dim] / np.asarray(out_scale)[:dim]
                  ).reshape((dim,) + (1,) * dim)
  reshaper = (slice(None),) + (np.newaxis,) * dim
  return coord_map.copy() * np.array(out_scale[:dim])[reshaper]


def ndimage_warp(image, coord_map, stride, order=1, image_start=None, map_start=None,
                 out_start=None, out_size=None, out_scale=None):
  """warp.ndimage_warp (warp.py:189-335) without the box tiling; starts / size are xyz."""
  image = np.asarray(image)
  dim = coord_map.shape[0]
  assert dim == image.ndim == len(stride)
  src_map = source_map(coord_map, stride, image_start, map_start, out_scale)
  out_shape = image.shape if out_size is None else tuple(out_size)[::-1][-dim:]
  if map_start is not None:
    ms = np.asarray(map_start)[:dim]
    os_ = np.asarray(out_start if out_start is not None else image_start)[:dim]
    offset = (ms * np.asarray(stride)[::-1] - os_)[::-1]
  else:
    offset = (0,) * dim
  grids = np.mgrid[tuple(slice(0, s) for s in out_shape)]
  src_coords = [(c - o) / s for c, s, o in zip(grids, stride, offset)]
  dense = [map_coordinates(m, src_coords, order=1) for m in src_map[::-1]]
  return map_coordinates(image, dense, order=order)
