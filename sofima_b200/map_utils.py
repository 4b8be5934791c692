"""Drop-in for the device-side part of `sofima.map_utils` (reference map_utils.py).

  compose_maps_fast        map_utils.py:616-734
  to_absolute / to_relative  map_utils.py:150-224   (host bookkeeping on the small map array)
  outer_box / inner_box    map_utils.py:307-389   (host bookkeeping)
  fill_missing             map_utils.py:227-304   (host, SciPy qhull like the reference)
  invert_map               map_utils.py:392-487   (host, SciPy qhull like the reference)
  resample_map             map_utils.py:490-546   (host, SciPy qhull like the reference)
  compose_maps             map_utils.py:549-613   (host, the slow scattered form; the per-node
                                                   gather is compose_maps_fast on the GPU)
  make_affine_map          map_utils.py:789-811   (host bookkeeping)
  mask_irregular           map_utils.py:737-786   (CUDA tensors: csrc/flowfilt.cu; NumPy: as upstream)

The scattered-data steps (Delaunay triangulation + piecewise-linear / nearest lookup) work
on the coarse map nodes -- thousands of points, not pixels -- and are scipy.spatial /
scipy.interpolate calls in the reference; they stay SciPy calls here, so the values are the
reference's.  The per-pixel work that follows them (warp.warp_subvolume, ndimage_warp) is
what runs on the GPU.

Coordinate maps are in the reference's relative format `[2 or 3, z, y, x]`.  NumPy
in -> NumPy out, CUDA torch tensors stay on the device.  The bilinear sampling
follows `jax.scipy.ndimage.map_coordinates(order=1)` operation by operation
(csrc/stitch.cuh); there is no CPU path.
"""

from __future__ import annotations

import collections.abc
import ctypes
from typing import Sequence

import numpy as np
from scipy import interpolate as _interpolate
from scipy import ndimage
from scipy import spatial as _spatial

from . import _native
from . import mesh as _mesh


def _as_vec(value, dim: int):
  if not isinstance(value, collections.abc.Sequence) and not isinstance(value, np.ndarray):
    return (value,) * dim
  assert len(value) == dim, f'Dimension mismatch: {value=} vs {dim=}'
  return tuple(value)


def _integral_starts(start, dim: int, name: str) -> list[int]:
  """The C ABI carries map origins as int64 (every caller in the reference passes box
  corners); a fractional origin would silently shift the composed map, so it is refused."""
  vals = np.asarray(start, dtype=np.float64).reshape(-1)[-dim:]
  if not np.all(np.isfinite(vals)) or np.any(vals != np.round(vals)):
    raise ValueError(f'{name} must be integral (got {vals.tolist()}): fractional map '
                     'origins are not supported by the CUDA backend')
  return [int(v) for v in vals]


def _identity_offsets(shape, stride, box=None):
  """Per-axis [z]yx absolute positions of the nodes of a map of `shape` (map_utils.py:128-147
  plus the box shift of :169-177)."""
  dim = len(shape)
  stride = _as_vec(stride, dim)
  grids = np.mgrid[tuple(slice(0, s) for s in shape)]
  offs = [g * step for g, step in zip(grids, stride)]
  if box is not None:
    if not np.all(np.asarray(shape)[::-1] == np.asarray(box.size)[:dim]):
      raise ValueError(f'box shape ({box.size}) mismatch with coord map ({shape})')
    offs = [o + start * step for o, step, start in zip(offs, stride, box.start[:dim][::-1])]
  return offs


def to_absolute(coord_map: np.ndarray, stride, box=None) -> np.ndarray:
  """Relative -> absolute coordinate map (map_utils.py:150-185): channel i (x, y[, z]) gets
  the position of its node added; `box` places the map in the global frame."""
  coord_map = np.array(coord_map)
  dim = coord_map.shape[0]
  offs = _identity_offsets(coord_map.shape[-dim:], stride, box)
  for i in range(dim):
    coord_map[i, ...] += offs[-(i + 1)]
  return coord_map


def to_relative(coord_map: np.ndarray, stride, box=None) -> np.ndarray:
  """Absolute -> relative coordinate map (map_utils.py:190-224)."""
  coord_map = np.array(coord_map)
  dim = coord_map.shape[0]
  offs = _identity_offsets(coord_map.shape[-dim:], stride, box)
  for i in range(dim):
    coord_map[i, ...] -= offs[-(i + 1)]
  return coord_map


def outer_box(coord_map: np.ndarray, box, stride, target_len=None):
  """Box covering every target position the map refers to (map_utils.py:307-342), in units
  of `target_len` (default: the map stride)."""
  from . import compat  # pylint: disable=g-import-not-at-top
  abs_map = to_absolute(coord_map, stride, box)
  dim = coord_map.shape[0]
  lens_xyz = _as_vec(target_len if target_len is not None else stride, dim)[::-1]
  start = np.array(box.start).copy()
  size = np.array(box.size).copy()
  for i, tl in enumerate(lens_xyz):
    lo, hi = np.nanmin(abs_map[i]), np.nanmax(abs_map[i])
    lo = int(lo) // tl
    start[i] = lo
    size[i] = -(int(-hi) // tl) - lo + 1
  return compat.BoundingBox(start=start, size=size)


def _interpolate_points(data_points, query_points, *values, method: str = 'linear'):
  """Scattered-data interpolation of several fields at once (map_utils.py:70-117):
  `data_points` / `query_points` are per-axis coordinate arrays (x, y[, z]); returns
  [len(values), n_query].  Outside the convex hull linear / cubic give NaN."""
  if len(data_points) != len(query_points):
    raise ValueError('Data and query points dimensionalities needs to match, are: '
                     f'{len(data_points)} and {len(query_points)}')
  if method == 'nearest':
    lookup = _interpolate.NearestNDInterpolator(data_points, values[0])
    fields = [lookup(query_points)]
    for val in values[1:]:
      lookup.values = val
      fields.append(lookup(query_points))
    return np.array(fields)
  if method not in ('linear', 'cubic'):
    raise ValueError(f'unknown interpolation method {method!r}')
  pts = np.ascontiguousarray(np.array(data_points).T, dtype=np.double)
  tri = _spatial.Delaunay(pts)
  stacked = np.array(values).T  # [n_points, n_fields]
  if method == 'linear':
    lookup = _interpolate.LinearNDInterpolator(tri, stacked, fill_value=np.nan)
  else:
    lookup = _interpolate.CloughTocher2DInterpolator(tri, stacked, fill_value=np.nan)
  return lookup(query_points).T


_QhullError = getattr(_spatial, 'QhullError', None) or _spatial.qhull.QhullError


def fill_missing(coord_map: np.ndarray, *, extrapolate=False, invalid_to_zero=False,
                 interpolate_first=True) -> np.ndarray:
  """Replaces non-finite nodes of a relative map (map_utils.py:227-304): linear
  interpolation inside the hull of the valid nodes, then (optionally) nearest-neighbour
  extrapolation; 2-d maps are processed section by section."""
  if not np.any(np.isnan(coord_map)):
    return coord_map
  dim = coord_map.shape[0]
  grid = np.mgrid[tuple(slice(0, n) for n in coord_map.shape[-dim:])]  # [z]yx
  axes_xyz = grid[::-1]
  queries = tuple(a.ravel() for a in axes_xyz)
  node_shape = coord_map.shape[-dim:]

  def fill(section):
    out = section.copy()
    valid = np.all(np.isfinite(section), axis=0)
    if not np.any(valid) and invalid_to_zero:
      out[...] = 0
      return out
    if interpolate_first:
      try:
        fields = _interpolate_points(tuple(a[valid] for a in axes_xyz), queries,
                                     *[c[valid] for c in section])
        for i, f in enumerate(fields):
          out[i, ...] = f.reshape(node_shape)
      except _QhullError:
        pass
    if extrapolate:
      valid = np.all(np.isfinite(out), axis=0)
      if not np.all(valid):
        fields = _interpolate_points(tuple(a[valid] for a in axes_xyz), queries,
                                     *[c[valid] for c in out], method='nearest')
        for i, f in enumerate(fields):
          out[i, ...] = f.reshape(node_shape)
    return out

  if dim == 2:
    return np.stack([fill(coord_map[:, z, ...]) for z in range(coord_map.shape[1])], axis=1)
  return fill(coord_map)


def inner_box(coord_map: np.ndarray, box, stride):
  """Box all of whose nodes are reached by the map (map_utils.py:345-389)."""
  from . import compat  # pylint: disable=g-import-not-at-top
  dim = coord_map.shape[0]
  assert dim in (2, 3)
  stride = _as_vec(stride, dim)
  full = to_absolute(fill_missing(coord_map, extrapolate=True), stride, box)
  lo, hi = [], []
  for i in range(dim):  # x, y[, z]; component i varies along array axis -(i + 1)
    axis = -(i + 1)
    step = stride[axis]
    a = np.max(np.min(full[i, ...], axis=axis))
    b = np.min(np.max(full[i, ...], axis=axis))
    lo.append(int(-(-a // step)))
    hi.append(b // step)
  if dim == 2:
    return compat.BoundingBox(start=(lo[0], lo[1], box.start[2]),
                              size=(hi[0] - lo[0] + 1, hi[1] - lo[1] + 1, box.size[2]))
  return compat.BoundingBox(start=lo, size=[h - l + 1 for l, h in zip(lo, hi)])


def invert_map(coord_map: np.ndarray, src_box, dst_box, stride) -> np.ndarray:
  """(x, y[, z]) -> (u, v[, w]) map inverted on the nodes of `dst_box`
  (map_utils.py:392-487): the valid source nodes are scattered at their targets and the
  source positions are interpolated linearly at the regular destination nodes."""
  coord_map = coord_map.astype(np.float64)
  dim = coord_map.shape[0]
  if dim not in (2, 3):
    raise NotImplementedError()
  stride = _as_vec(stride, dim)
  # frame with its origin at the first destination node
  src_box = src_box.adjusted_by(start=-dst_box.start, end=-dst_box.start)
  dst_box = dst_box.adjusted_by(start=-dst_box.start, end=-dst_box.start)
  coord_map = to_absolute(coord_map, stride, src_box)

  def node_positions(box):  # [z]yx arrays of node positions in pixels
    grid = np.mgrid[tuple(slice(0, int(n)) for n in box.size[:dim][::-1])]
    for i in range(dim):
      grid[i] = (grid[i] + box.start[dim - i - 1]) * stride[i]
    return grid

  src_pos = node_positions(src_box)
  dst_pos = node_positions(dst_box)
  queries = tuple(q.ravel() for q in dst_pos[::-1])  # uv[w]
  if dim == 2:
    out = np.full((2, coord_map.shape[1], dst_box.size[1], dst_box.size[0]), np.nan,
                  dtype=coord_map.dtype)
    for z in range(coord_map.shape[1]):
      valid = np.all(np.isfinite(coord_map[:, z, ...]), axis=0)
      if not np.any(valid):
        continue
      try:
        u, v = _interpolate_points(tuple(c[z][valid] for c in coord_map), queries,
                                   *[p[valid] for p in src_pos[::-1]])
      except _QhullError:
        continue
      out[0, z, ...] = u.reshape(dst_pos[1].shape)
      out[1, z, ...] = v.reshape(dst_pos[0].shape)
    return to_relative(out, stride, dst_box)
  out = np.full((3, dst_box.size[2], dst_box.size[1], dst_box.size[0]), np.nan,
                dtype=coord_map.dtype)
  valid = np.all(np.isfinite(coord_map), axis=0)
  if not np.any(valid):
    return out
  try:
    fields = _interpolate_points(tuple(c[valid] for c in coord_map), queries,
                                 *[p[valid] for p in src_pos[::-1]])
    for i, f in enumerate(fields):
      out[i, ...] = f.reshape(dst_pos[0].shape)
  except _QhullError:
    pass
  return to_relative(out, stride, dst_box)


def resample_map(coord_map: np.ndarray, src_box, dst_box, src_stride: float,
                 dst_stride: float, method='linear') -> np.ndarray:
  """Relative [2, z, y, x] map resampled on a grid of another spacing
  (map_utils.py:490-546)."""
  assert coord_map.shape[0] == 2
  sy, sx = np.mgrid[:src_box.size[1], :src_box.size[0]]
  sy = (sy + src_box.start[1]) * src_stride
  sx = (sx + src_box.start[0]) * src_stride
  ty, tx = np.mgrid[:dst_box.size[1], :dst_box.size[0]]
  ty = (ty + dst_box.start[1]) * dst_stride
  tx = (tx + dst_box.start[0]) * dst_stride
  out = np.full((2, coord_map.shape[1], dst_box.size[1], dst_box.size[0]), np.nan,
                dtype=coord_map.dtype)
  for z in range(coord_map.shape[1]):
    valid = np.isfinite(coord_map[0, z, ...])
    if not np.any(valid):
      continue
    try:
      u, v = _interpolate_points((sx[valid], sy[valid]), (tx.ravel(), ty.ravel()),
                                 coord_map[0, z, ...][valid], coord_map[1, z, ...][valid],
                                 method=method)
    except _QhullError:
      continue
    out[0, z, ...] = u.reshape(tx.shape)
    out[1, z, ...] = v.reshape(ty.shape)
  return out


def compose_maps(map1: np.ndarray, box1, stride1: float, map2: np.ndarray, box2,
                 stride2: float) -> np.ndarray:
  """map2(map1(x, y)) for [2, z, y, x] relative maps by scattered-data interpolation
  (map_utils.py:549-613): map2's valid nodes are the data points, map1's absolute targets
  the queries; invalid nodes of map2 are thereby interpolated over."""
  assert map1.shape[0] == 2
  assert map2.shape[0] == 2
  abs1 = to_absolute(map1, stride1, box1)
  abs2 = to_absolute(map2, stride2, box2)
  out = np.full_like(map1, np.nan)
  sy, sx = np.mgrid[box2.start[1]:box2.end[1], box2.start[0]:box2.end[0]]
  sx = sx * stride2
  sy = sy * stride2
  for z in range(map1.shape[1]):
    ok1 = np.all(np.isfinite(abs1[:, z, ...]), axis=0)
    ok2 = np.all(np.isfinite(abs2[:, z, ...]), axis=0)
    if not np.any(ok1) or not np.any(ok2):
      continue
    try:
      u, v = _interpolate_points((sx[ok2], sy[ok2]),
                                 (abs1[0, z, ...][ok1], abs1[1, z, ...][ok1]),
                                 abs2[0, z, ...][ok2], abs2[1, z, ...][ok2])
    except _QhullError:
      continue
    out[0, z, ...][ok1] = u
    out[1, z, ...][ok1] = v
  return to_relative(out, stride1, box1)


def make_affine_map(matrix: np.ndarray, box, stride) -> np.ndarray:
  """Relative [3, z, y, x] map of an affine transform (map_utils.py:789-811); `matrix` is
  [3, 4] in the format of ndimage.affine_transform, acting on (x, y, z)."""
  zyx = _identity_offsets(tuple(int(v) for v in box.size[::-1]), stride)
  coords = np.array(zyx[::-1], dtype=np.float64)  # x, y, z
  for i in range(3):
    coords[i, ...] += box.start[i]
  moved = (np.dot(matrix[:3, :3], coords.reshape((3, -1)))
           + matrix[:, 3][:, np.newaxis]).reshape(coords.shape)
  return moved - coords


def mask_irregular(coord_map: np.ndarray, stride: Sequence[float], frac: float,
                   max_frac: float | None = None, dilation_iters: int = 1) -> np.ndarray:
  """Marks stretched / folded nodes of a [2, y, x] relative map (map_utils.py:737-786).

  A node is bad if the distance to its +x or +y neighbour is below `frac` or above
  `max_frac` (default 2 - frac) times the stride; the bad set is dilated with a full
  3x3 structuring element.  Bad nodes are set to NaN in place; returns the mask.
  """
  assert coord_map.ndim == 3 and coord_map.shape[0] == 2
  if type(coord_map).__module__.startswith('torch') and coord_map.is_cuda:
    # device-resident map: csrc/flowfilt.cu (same result, the mesh never leaves the GPU)
    import ctypes
    import torch
    if coord_map.dtype != torch.float32 or not coord_map.is_contiguous():
      raise ValueError('device mask_irregular needs a contiguous float32 [2, y, x] tensor')
    ctx = _native.Context.get(coord_map.device.index)
    ctx.bind_stream()
    bad = torch.empty(coord_map.shape[1:], dtype=torch.uint8, device=coord_map.device)
    rc = _native.lib().sofima_mask_irregular(
        ctx.handle, coord_map.data_ptr(), int(coord_map.shape[1]), int(coord_map.shape[2]),
        (ctypes.c_double * 2)(float(stride[0]), float(stride[1])), float(frac),
        float(2 - frac if max_frac is None else max_frac), int(dilation_iters), bad.data_ptr())
    _native.check(ctx.handle, rc)
    return bad.bool()
  # as upstream: the stride (a NumPy scalar of the `stride` array) promotes the fp32
  # differences to float64 before the comparisons
  stride = np.asarray(stride)
  stride_x, stride_y = stride
  hi = 2 - frac if max_frac is None else max_frac
  diff_x = np.pad(np.diff(coord_map[0], axis=-1), [[0, 0], [0, 1]], mode='constant') + stride_x
  diff_y = np.pad(np.diff(coord_map[1], axis=-2), [[0, 1], [0, 0]], mode='constant') + stride_y
  with np.errstate(invalid='ignore'):
    bad = (diff_x < frac * stride_x) | (diff_y < frac * stride_y)
    bad |= (diff_x > hi * stride_x) | (diff_y > hi * stride_y)
  if dilation_iters > 0:
    bad = ndimage.binary_dilation(bad, ndimage.generate_binary_structure(2, 2),
                                  iterations=dilation_iters)
  coord_map[0][bad] = np.nan
  coord_map[1][bad] = np.nan
  return bad


def compose_maps_fast(map1, start1: Sequence[float], stride1, map2,
                      start2: Sequence[float], stride2, mode='nearest'):
  """Composes two coordinate maps: map2(map1(z, y, x)) on the grid of map1.

  Same contract as map_utils.compose_maps_fast (map_utils.py:616-734): invalid
  (NaN) values in either map are NOT interpolated.

  Args:
    map1: [2 or 3, z, y, x] 1st coordinate map in relative format
    start1: [z]yx origin coordinates for map1 (longer sequences: the last entries)
    stride1: distance between nearest neighbors of map1 (scalar or [z]yx)
    map2, start2, stride2: same for the 2nd map
    mode: 'nearest' or 'constant' (cval NaN), as passed to map_coordinates

  Returns:
    [2 or 3, z, y, x] composed map covering the area of map1 (with stride1)
  """
  assert map1.shape[0] == map2.shape[0]
  dim = int(map1.shape[0])
  if dim not in (2, 3) or len(map1.shape) != 4 or len(map2.shape) != 4:
    raise ValueError('maps must be [2 or 3, z, y, x]')
  if mode not in ('nearest', 'constant'):
    raise NotImplementedError(f"mode {mode!r}: only 'nearest' and 'constant' are built")
  stride1 = _as_vec(stride1, dim)
  stride2 = _as_vec(stride2, dim)
  s1 = _integral_starts(start1, dim, 'start1')
  s2 = _integral_starts(start2, dim, 'start2')
  if len(s1) != dim or len(s2) != dim:
    raise ValueError('start1 / start2 need at least `dim` entries')
  if dim == 2 and map1.shape[1] != map2.shape[1]:
    raise ValueError('2-d maps need the same number of sections')

  dev = map1.device.index if _mesh._is_tensor(map1) and map1.is_cuda else None
  ctx = _native.Context.get(dev)
  m1 = _mesh._to_device(map1, ctx, copy=False)
  m2 = _mesh._to_device(map2, ctx, copy=False)
  out = _mesh._torch().empty_like(m1)
  i64x3 = ctypes.c_int64 * 3
  ctx.bind_stream()
  rc = _native.lib().sofima_compose_maps(
      ctx.handle, dim, m1.data_ptr(), i64x3(*m1.shape[1:]), (ctypes.c_int64 * dim)(*s1),
      (ctypes.c_double * dim)(*[float(v) for v in stride1]), m2.data_ptr(),
      i64x3(*m2.shape[1:]), (ctypes.c_int64 * dim)(*s2),
      (ctypes.c_double * dim)(*[float(v) for v in stride2]),
      int(mode == 'constant'), out.data_ptr())
  _native.check(ctx.handle, rc)
  return _mesh._from_device(out, map1)
