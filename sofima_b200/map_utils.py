"""Drop-in for the device-side part of `sofima.map_utils` (reference map_utils.py).

  compose_maps_fast        map_utils.py:616-734
  to_absolute / to_relative  map_utils.py:150-224   (host bookkeeping on the small map array)
  outer_box                map_utils.py:307-342   (host bookkeeping)

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
