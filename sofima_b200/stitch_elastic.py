"""Drop-in for the solver-facing part of `sofima.stitch_elastic` (reference
stitch_elastic.py): the per-step target mesh of elastic tile stitching.

  NeighborInfo             stitch_elastic.py:43-72
  compute_flow_map         stitch_elastic.py:197-282   (host geometry + CUDA flow_field)
  compute_flow_map3d       stitch_elastic.py:84-193
  aggregate_arrays         stitch_elastic.py:285-453   (host, NumPy -- as in the reference)
  compute_target_mesh      stitch_elastic.py:624-676   (CUDA, csrc/stitch.cuh)
  target_mesh_fn           the `prev_fn` closure the notebooks build around it
                           (notebooks/em_stitching.ipynb:545-549)

The reference's `prev_fn` is a JAX callable traced into the jitted integrator; here it
is a device-side description (`StitchTarget`) that `mesh.relax_mesh(..., prev_fn=...)`
hands to the C ABI, which re-evaluates the target inside every integration step.
"""

from __future__ import annotations

import ctypes
import enum
from typing import Mapping, Sequence

import numpy as np

from . import _native
from . import compat
from . import flow_field
from . import mesh as _mesh


class NeighborInfo(enum.IntEnum):
  """Indices into a neighbour-info row (stitch_elastic.py:43-72)."""
  nbor_idx = 0             # neighbouring tile index
  flow_idx = 1             # index within the flow array
  coarse_offset_ortho = 2  # coarse offset orthogonal to the overlap dim (pixels)
  flow_size_ortho = 3      # flow extent orthogonal to the overlap dim
  flow_size_overlap = 4    # flow extent along the overlap dim
  fine_off_x = 5           # offset vector used when the flow was computed
  fine_off_y = 6
  dim = 7                  # 0: x neighbour, 1: y neighbour
  coarse_offset_z = 8      # 3-d meshes only
  flow_size_z = 9
  fine_off_z = 10


def _relative_intersection(box1, box2):
  """The overlap of two boxes in the coordinates of each (stitch_elastic.py:75-81)."""
  common = box1.intersection(box2)
  return (compat.BoundingBox(start=common.start - box1.start, size=common.size),
          compat.BoundingBox(start=common.start - box2.start, size=common.size))


def compute_flow_map(tile_map: Mapping[tuple[int, int], np.ndarray], offset_map: np.ndarray,
                     axis: int, patch_size: Sequence[int] = (120, 120),
                     stride: Sequence[int] = (20, 20), batch_size: int = 256):
  """Fine flow between horizontally (axis 0) or vertically (axis 1) adjacent 2-d tiles.

  Same contract as stitch_elastic.compute_flow_map (stitch_elastic.py:197-282): the
  overlap strips of the two tiles are cut so that they start at a multiple of `stride`
  in the pre tile, the flow between them is estimated on the GPU, and the result is
  NaN-padded by half a patch so that it is aligned with the mesh nodes.

  Args:
    tile_map: (x, y) -> tile image
    offset_map: [2, y, x] coarse XY offset between tile (x, y) and its +axis neighbour
    axis: 0 = neighbour at (x + 1, y), 1 = neighbour at (x, y + 1)
    patch_size: YX patch size in pixels
    stride: YX stride of the flow map in pixels
    batch_size: flow vectors estimated per batch

  Returns:
    ((x, y) -> [4, fy, fx] flow, (x, y) -> XY offset the flow was computed with)
  """
  ny_t, nx_t = offset_map.shape[-2:]
  calc = flow_field.JAXMaskedXCorrWithStatsCalculator()
  step = tuple(int(v) for v in stride)
  stride = np.asarray(step)
  pad = [patch_size[0] // 2 // step[0], patch_size[1] // 2 // step[1]]
  flows, offsets, pending = {}, {}, {}
  # Every strip pair is queued on the GPU before the first flow field is read back: the
  # strips are short (a few hundred patch pairs), and a host round trip per strip would cost
  # more than the correlation itself.
  for y in range(ny_t - axis):
    for x in range(nx_t - (1 - axis)):
      if np.isnan(offset_map[0, y, x]):
        continue
      pre = tile_map[x, y]
      post = tile_map[x + (1 - axis), y + axis]
      offset = offset_map[:, y, x]  # XY
      rounded = stride[::-1] * np.round(offset / stride[::-1])
      par = 1 - axis  # array axis along the tile-tile direction
      overlap = -int(offset[axis])
      overlap = pre.shape[par] - (pre.shape[par] - overlap) // stride[par] * stride[par]
      ortho = int(rounded[1 - axis])

      pre_sel = [slice(None), slice(None)]
      post_sel = [slice(None), slice(None)]
      pre_sel[par] = slice(-overlap, None)
      post_sel[par] = slice(None, overlap)
      if ortho > 0:    # post is shifted towards +ortho relative to pre
        pre_sel[axis] = slice(ortho, None)
        post_sel[axis] = slice(None, -ortho)
      elif ortho < 0:
        pre_sel[axis] = slice(None, ortho)
        post_sel[axis] = slice(-ortho, None)
      kw = {'_async': True} if getattr(calc, 'supports_async', False) else {}
      pending[x, y] = calc.flow_field(pre[tuple(pre_sel)], post[tuple(post_sel)],
                                      patch_size=tuple(patch_size), step=step,
                                      batch_size=batch_size, **kw)
      offsets[x, y] = (-overlap, ortho) if axis == 0 else (ortho, -overlap)
  for key, handle in pending.items():
    # The inverse flow (post, pre) is -f: it is not computed separately.
    f = handle.result() if hasattr(handle, 'result') else handle
    flows[key] = np.pad(f, [[0, 0], [pad[0], pad[0] - 1], [pad[1], pad[1] - 1]],
                        constant_values=np.nan)
  return flows, offsets


def compute_flow_map3d(tile_map: Mapping[tuple[int, int], np.ndarray],
                       tile_shape: Sequence[int], offset_map: np.ndarray, axis: int,
                       patch_size: Sequence[int] = (120, 120, 120),
                       stride: Sequence[int] = (40, 40, 40), batch_size: int = 16):
  """Fine flow between adjacent 3-d tiles (stitch_elastic.py:84-193).

  Args:
    tile_map: (x, y) -> [1, z, y, x] tile data
    tile_shape: XYZ shape of a tile
    offset_map: [3, 1, y, x] coarse XYZ offsets between (x, y) and its +axis neighbour
    axis: 0 = x neighbour, 1 = y neighbour
    patch_size, stride: ZYX, in pixels
    batch_size: flow vectors estimated per batch

  Returns:
    ((x, y) -> [5, fz, fy, fx] flow, (x, y) -> XYZ offset the flow was computed with)
  """
  calc = flow_field.JAXMaskedXCorrWithStatsCalculator()
  flows, offsets = {}, {}
  ny_t, nx_t = offset_map.shape[-2:]
  step = tuple(int(v) for v in stride)
  pad = np.array(patch_size) // 2 // np.array(step)
  tile_shape = tuple(int(v) for v in tile_shape)
  s = step[2 - axis]
  for y in range(ny_t - axis):
    for x in range(nx_t - (1 - axis)):
      coarse = offset_map[:, 0, y, x]  # XYZ
      here = compat.BoundingBox(start=(0, 0, 0), size=tile_shape)
      nbor = compat.BoundingBox(
          start=(tile_shape[0] * (1 - axis) + coarse[0], tile_shape[1] * axis + coarse[1],
                 coarse[2]), size=tile_shape)
      isec_here, isec_nbor = _relative_intersection(here, nbor)
      # Along the tile-tile direction the strip starts at a multiple of the stride
      # inside the preceding tile ...
      overlap = isec_here.size[axis]
      aligned_start = (tile_shape[axis] - overlap) // s * s
      shift = np.zeros(3)
      shift[axis] = -((tile_shape[axis] - aligned_start) - overlap)
      # ... and so do the starts in the two orthogonal directions.
      for ax in range(3):
        if ax == axis:
          continue
        if isec_here.start[ax] > 0:
          shift[ax] = s * np.round(isec_here.start[ax] / s) - isec_here.start[ax]
        elif isec_nbor.start[ax] > 0:
          shift[ax] = -(s * np.round(isec_nbor.start[ax] / s) - isec_nbor.start[ax])
      nbor = nbor.translate(shift)
      isec_here, isec_nbor = _relative_intersection(here, nbor)
      assert np.all(isec_here.start % s == 0)
      assert np.all(isec_nbor.start % s == 0)
      offset = np.array(nbor.start - here.start)
      offset[axis] = -isec_here.size[axis]
      offsets[x, y] = tuple(offset.tolist())
      pre = tile_map[x, y][isec_here.to_slice4d()].squeeze(axis=0)
      post = tile_map[x + (1 - axis), y + axis][isec_nbor.to_slice4d()].squeeze(axis=0)
      assert pre.shape == post.shape
      f = calc.flow_field(pre, post, patch_size=tuple(patch_size), step=step,
                          batch_size=batch_size)
      flows[x, y] = np.pad(f, [[0, 0]] + [[int(q), int(q) - 1] for q in pad],
                           constant_values=np.nan)
  return flows, offsets


def aggregate_arrays(x_data, y_data, tile_coords: Sequence[tuple[int, int]],
                     coarse_mesh: np.ndarray, stride: Sequence[float],
                     tile_shape: Sequence[int]):
  """Aggregates the per-pair flow fields of a tile grid into dense arrays.

  Same contract as stitch_elastic.aggregate_arrays (stitch_elastic.py:285-453).

  Args:
    x_data: (coarse offsets [2 or 3, ty, tx] between (x, y) and (x+1, y); dict
      tile -> fine flow; dict tile -> offset vector used for the fine flow)
    y_data: the same for (x, y) and (x, y+1)
    tile_coords: (x, y) tile coordinates
    coarse_mesh: [2 or 3, ty, tx] rigid solution (initial tile positions)
    stride: [z]yx stride of mesh and flow grids in pixels
    tile_shape: [z]yx tile shape in pixels

  Returns:
    fx [dim, N, ...], fy [dim, N, ...] NaN-padded flows; x [dim, N, ...] initial
    meshes; nbors [N, 4, 8 or 11] int NeighborInfo table; dict tile -> index
  """
  cx, fine_x, offsets_x = x_data
  cy, fine_y, offsets_y = y_data
  assert cx.ndim == 3 and cy.ndim == 3
  index = {tuple(c): i for i, c in enumerate(tile_coords)}
  dim = len(stride)
  ntiles = len(index)

  def dense(fine: Mapping[tuple[int, int], np.ndarray]) -> np.ndarray:
    extent = np.ones(dim, dtype=int)
    for f in fine.values():
      extent = np.maximum(extent, f.shape[1:])
    arr = np.full((dim, ntiles) + tuple(int(e) for e in extent), np.nan)
    for key, i in index.items():
      f = fine.get(key)
      if f is not None:
        arr[(slice(None), i) + tuple(slice(0, s) for s in f.shape[1:])] = f[:dim]
    return arr

  fx, fy = dense(fine_x), dense(fine_y)

  def row(nbor_key, flow_key, coarse, fine, offsets, axis):
    shape = fine[flow_key].shape
    ortho, overlap = shape[-2], shape[-1]
    if axis == 1:
      ortho, overlap = overlap, ortho
    off = offsets[flow_key]
    vals = [index[nbor_key], index[flow_key], coarse[1] if axis == 0 else coarse[0],
            ortho, overlap, off[0], off[1], axis]
    if dim == 3:
      vals += [coarse[2], shape[-3], off[2]]
    return vals

  nbors = np.full((ntiles, 4, 8 if dim == 2 else 11), -1, dtype=int)
  for tx, ty in tile_coords:
    i = index[tx, ty]
    if (tx - 1, ty) in fine_x:   # left neighbour: its flow, we are the 'post' tile
      nbors[i, 0] = row((tx - 1, ty), (tx - 1, ty), cx[:, ty, tx - 1], fine_x, offsets_x, 0)
    if (tx, ty) in fine_x:       # right neighbour: our flow
      nbors[i, 1] = row((tx + 1, ty), (tx, ty), cx[:, ty, tx], fine_x, offsets_x, 0)
    if (tx, ty - 1) in fine_y:   # neighbour above
      nbors[i, 2] = row((tx, ty - 1), (tx, ty - 1), cy[:, ty - 1, tx], fine_y, offsets_y, 1)
    if (tx, ty) in fine_y:       # neighbour below
      nbors[i, 3] = row((tx, ty + 1), (tx, ty), cy[:, ty, tx], fine_y, offsets_y, 1)

  mesh_shape = [int(s) for s in (np.array(tile_shape) // np.array(stride))]
  x = np.zeros([dim, ntiles] + mesh_shape, dtype=np.float32)
  for tx, ty in tile_coords:
    x[:, index[tx, ty]] = np.asarray(coarse_mesh[:, ty, tx]).reshape((dim,) + (1,) * dim)
  return fx, fy, x, nbors, index


class StitchTarget:
  """Device-side `prev_fn` for elastic stitching.

  Calling it evaluates the target mesh of every tile for the tile meshes `x`
  ([2, N, y, x]) -- what the notebooks' `prev_fn` returns.  Passed as
  `mesh.relax_mesh(x, None, config, prev_fn=target)` the same computation runs inside
  the integrator, before every force evaluation (mesh.py:429-430).
  """

  def __init__(self, nbors, fx, fy, stride: Sequence[float] = (20, 20)):
    nbors = np.asarray(nbors)
    if nbors.ndim != 3 or nbors.shape[1] != 4:
      raise ValueError(f'nbors must be [n, 4, 8 or 11], got {nbors.shape}')
    self.dim = len(stride)
    if self.dim not in (2, 3) or nbors.shape[2] != (8 if self.dim == 2 else 11):
      raise ValueError('stride must be [z]yx and nbors [n, 4, 8] (2-d) or [n, 4, 11] (3-d)')
    n = nbors.shape[0]
    if np.any(nbors[:, :, 0] < -1) or np.any(nbors[:, :, 0] >= n) or np.any(
        (nbors[:, :, 0] >= 0) & ((nbors[:, :, 1] < 0) | (nbors[:, :, 1] >= n))):
      raise ValueError('neighbour / flow indices out of range')
    self.ntiles = n
    self.stride = tuple(float(s) for s in stride)
    self._nbors_host = np.ascontiguousarray(nbors, dtype=np.int32)
    self._fx_in, self._fy_in = fx, fy
    for f in (fx, fy):
      if len(f.shape) != self.dim + 2 or f.shape[0] != self.dim or f.shape[1] != n:
        raise ValueError(
            f'flow arrays must be [{self.dim}, {n}, [z,] y, x], got {tuple(f.shape)}')
    self._dev = {}  # device index -> (fx, fy, nbors) tensors

  def _arrays(self, ctx):
    arrs = self._dev.get(ctx.device)
    if arrs is None:
      torch = _mesh._torch()
      arrs = (_mesh._to_device(self._fx_in, ctx, copy=False),
              _mesh._to_device(self._fy_in, ctx, copy=False),
              torch.from_numpy(self._nbors_host).to(torch.device('cuda', ctx.device)))
      self._dev[ctx.device] = arrs
    return arrs

  def _sofima_device_target(self, x_shape, ctx) -> _native.StitchTargetPod:
    d = self.dim
    if len(x_shape) != d + 2 or x_shape[0] != d or x_shape[1] != self.ntiles:
      raise ValueError(f'x must be [{d}, {self.ntiles}, [z,] y, x], got {tuple(x_shape)}')
    fx, fy, nb = self._arrays(ctx)
    pod = _native.StitchTargetPod()
    pod.fx, pod.fy, pod.nbors = fx.data_ptr(), fy.data_ptr(), nb.data_ptr()
    pod.ndim = d
    for a in range(3):
      j = a - (3 - d)  # index into the [z]yx tuples
      pod.fx_shape[a] = fx.shape[2 + j] if j >= 0 else 1
      pod.fy_shape[a] = fy.shape[2 + j] if j >= 0 else 1
      pod.stride[a] = self.stride[j] if j >= 0 else 0.0
    return pod

  def __call__(self, x):
    dev = x.device.index if _mesh._is_tensor(x) and x.is_cuda else None
    ctx = _native.Context.get(dev)
    xd = _mesh._to_device(x, ctx, copy=False)
    pod = self._sofima_device_target(tuple(xd.shape), ctx)
    if self.dim == 2:
      shape = _native.MeshShape(2, xd.shape[1], 1, xd.shape[2], xd.shape[3], 0)
    else:
      shape = _native.MeshShape(3, xd.shape[1], xd.shape[2], xd.shape[3], xd.shape[4], 1)
    out = _mesh._torch().empty_like(xd)
    ctx.bind_stream()
    rc = _native.lib().sofima_stitch_target_mesh(
        ctx.handle, xd.data_ptr(), ctypes.byref(shape), ctypes.byref(pod), out.data_ptr())
    _native.check(ctx.handle, rc)
    return _mesh._from_device(out, x)


def target_mesh_fn(nbors, fx, fy, stride: Sequence[float] = (20, 20)) -> StitchTarget:
  """`prev_fn` for `mesh.relax_mesh`: vmap(compute_target_mesh)(nbors) on the device."""
  return StitchTarget(nbors, fx, fy, stride)


def compute_target_mesh(nbor_data, x, fx, fy, stride: Sequence[float] = (20, 20)):
  """Target mesh of ONE tile (stitch_elastic.py:624-676).

  Args:
    nbor_data: [4, 8 or 11] neighbour info of the tile; -1 in the nbor and flow
      indices marks missing entries
    x: [2 or 3, n, [z,] y, x] node positions of all tiles
    fx, fy: [2 or 3, n, [z,] y, x] flows between horizontal / vertical neighbours
    stride: [z]yx stride of flow and mesh data

  Returns:
    [2 or 3, [z,] y, x] target positions (NaN where no neighbour provides one)
  """
  nbor_data = np.asarray(nbor_data)
  n = x.shape[1]
  table = np.full((n,) + nbor_data.shape, -1, dtype=int)
  table[0] = nbor_data
  return StitchTarget(table, fx, fy, stride)(x)[:, 0]
