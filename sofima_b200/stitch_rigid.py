"""Drop-in for the coarse tile-offset estimation of `sofima.stitch_rigid` (reference
stitch_rigid.py):

  _estimate_offset[_horiz|_vert]   stitch_rigid.py:39-101
  compute_coarse_offsets           stitch_rigid.py:104-273
  interpolate_missing_offsets      stitch_rigid.py:277-327 (host NumPy, as upstream)

The offset between two neighbouring tiles is the peak of ONE masked normalised
cross-correlation of their whole overlap strips (e.g. 4096 x 300 px -> 8192 x 600-point
transforms), computed by the CUDA flow path (long-column form of csrc/flow.cu).  The
low-contrast masks are SciPy min / max filters on the host, as in the reference.  The
rigid mesh optimisation on top of the offsets (`optimize_coarse_mesh` with
`elastic_tile_mesh[_3d]`, stitch_rigid.py:330-545) relaxes a one-node-per-tile mesh with
its own linear spring force: csrc/tile_mesh.cu runs a whole chunk of integration steps in
one thread block (`sofima_tile_mesh_chunk`).
"""

from __future__ import annotations

import ctypes
import dataclasses
from typing import Mapping, Sequence

import numpy as np
from scipy import ndimage

from . import _native
from . import flow_field
from . import mesh

MaskMap = Mapping[tuple[int, int], np.ndarray]


def _estimate_offset(a: np.ndarray, b: np.ndarray, range_limit: float, filter_size: int = 10,
                     masks=None):
  """Global offset between two equally shaped strips: ([x, y], |peak ratio|)."""
  def flat(img):  # areas with insufficient dynamic range are masked
    return (ndimage.maximum_filter(img, filter_size)
            - ndimage.minimum_filter(img, filter_size)) < range_limit
  a_mask, b_mask = flat(a), flat(b)
  if masks is not None:
    a_mask |= masks[0]
    b_mask |= masks[1]
  calc = flow_field.JAXMaskedXCorrWithStatsCalculator()
  xo, yo, _, pr = calc.flow_field(a, b, pre_mask=a_mask, post_mask=b_mask,
                                  patch_size=a.shape, step=(1, 1), batch_size=1).squeeze()
  return [xo, yo], abs(pr)


def _estimate_offset_horiz(overlap, left, right, range_limit, filter_size, masks=None):
  return _estimate_offset(left[:, -overlap:], right[:, :overlap], range_limit,
                          filter_size=filter_size, masks=masks)


def _estimate_offset_vert(overlap, top, bot, range_limit, filter_size, masks=None):
  return _estimate_offset(top[-overlap:, :], bot[:overlap, :], range_limit,
                          filter_size=filter_size, masks=masks)


def compute_coarse_offsets(yx_shape: tuple[int, int],
                           tile_map: Mapping[tuple[int, int], np.ndarray],
                           overlaps_xy=((200, 300), (200, 300)),
                           min_range=(10, 100, 0), min_overlap=160, filter_size=10,
                           mask_map: MaskMap | None = None):
  """Coarse XY offset between every pair of neighbouring tiles.

  Same contract as stitch_rigid.compute_coarse_offsets (stitch_rigid.py:104-273): for
  every pair the overlap widths in `overlaps_xy` are tried with decreasing contrast
  thresholds `min_range`; a single-peak correlation ends the search, otherwise the
  estimate that is consistent between two consecutive overlap widths (< 20 px apart)
  wins, else the valid one with the highest peak ratio.

  Returns:
    (conn_x, conn_y), each [2, 1, ny, nx]: XY offsets between tile (x, y) and
    (x + 1, y) resp. (x, y + 1); inf where no acceptable estimate exists, NaN where
    a tile is missing.
  """

  def find_offset(estimate_fn, pre, post, overlaps, max_ortho_shift, axis, masks=None):
    def acceptable(off):
      return abs(off[1 - axis]) < max_ortho_shift and abs(off[axis]) >= min_overlap

    found = False
    offset = None
    for range_limit in min_range:
      if found:
        break
      best_idx, best_pr = -1, 0
      estimates = []
      for overlap in overlaps:
        ov_masks = None
        if masks is not None:
          ma = masks[0][:, -overlap:] if axis == 0 else masks[0][-overlap:, :]
          mb = masks[1][:, :overlap] if axis == 0 else masks[1][:overlap, :]
          # a completely masked overlap disables the custom mask
          ma = np.full_like(ma, fill_value=False) if np.all(ma) else ma
          mb = np.full_like(mb, fill_value=False) if np.all(mb) else mb
          ov_masks = (ma, mb)
        offset, pr = estimate_fn(overlap, pre, post, range_limit, filter_size, ov_masks)
        offset[axis] -= overlap
        if pr == 0.0:  # single peak: accept immediately
          found = True
          break
        estimates.append(offset)
        if pr > best_pr and acceptable(offset):
          best_pr, best_idx = pr, len(estimates) - 1
      if found:
        break
      closest, closest_idx = np.inf, 0
      for i, (first, second) in enumerate(zip(estimates, estimates[1:])):
        gap = np.abs(second[axis] - first[axis])
        if gap < closest and acceptable(second):
          closest, closest_idx = gap, i
      if closest < 20:
        offset = estimates[closest_idx + 1]
        found = True
      elif best_idx >= 0:
        offset = estimates[best_idx]
        found = True
    if not found or abs(offset[axis]) < min_overlap:
      offset = np.inf, np.inf
    return offset

  ny, nx = yx_shape
  conn_x = np.full((2, 1, ny, nx), np.nan)
  for x in range(nx - 1):
    for y in range(ny):
      if (x, y) not in tile_map or (x + 1, y) not in tile_map:
        continue
      masks = None
      if mask_map is not None:
        width = max(overlaps_xy[0])
        masks = (mask_map[x, y][:, -width:], mask_map[x + 1, y][:, :width])
      conn_x[:, 0, y, x] = find_offset(_estimate_offset_horiz, tile_map[x, y],
                                       tile_map[x + 1, y], overlaps_xy[0],
                                       max(overlaps_xy[1]), 0, masks)
  conn_y = np.full((2, 1, ny, nx), np.nan)
  for y in range(ny - 1):
    for x in range(nx):
      if (x, y) not in tile_map or (x, y + 1) not in tile_map:
        continue
      masks = None
      if mask_map is not None:
        width = max(overlaps_xy[1])
        masks = (mask_map[x, y][-width:], mask_map[x, y + 1][:width])
      conn_y[:, 0, y, x] = find_offset(_estimate_offset_vert, tile_map[x, y],
                                       tile_map[x, y + 1], overlaps_xy[1],
                                       max(overlaps_xy[0]), 1, masks)
  return conn_x, conn_y


def interpolate_missing_offsets(conn: np.ndarray, axis: int, max_r: int = 4) -> np.ndarray:
  """Estimates missing coarse offsets (stitch_rigid.py:277-327).

  Offsets marked inf ("no acceptable estimate") are replaced by the mean of the finite
  offsets found at the smallest distance 1 <= r < max_r on either side along `axis`.

  Args:
    conn: [2 or 3, 1, y, x] coarse offsets as returned by compute_coarse_offsets;
      modified in place
    axis: axis of `conn` along which neighbours are searched (-1: x, -2: y)
    max_r: search radius (exclusive)

  Returns:
    conn; still inf where no finite neighbour lies within the radius
  """
  if conn.ndim != 4:
    raise ValueError('conn array must have rank 4')
  length = conn.shape[axis]
  for y, x in zip(*np.where(np.isinf(conn[0, 0, ...]))):
    centre = (y, x)[axis + 2]
    for r in range(1, max_r):
      hits = []
      for q in (centre - r, centre + r):
        if not 0 <= q < length:
          continue
        where = [0, 0, y, x]
        where[axis] = q
        if np.isfinite(conn[tuple(where)]):
          where[0] = slice(None)
          hits.append(conn[tuple(where)])
      if hits:
        conn[:, 0, y, x] = np.mean(hits, axis=0)
        break
  return conn


def _tile_arrays(ctx, *arrays):
  """fp32 contiguous CUDA copies of [ncomp, z, y, x] arrays of one common shape."""
  torch = mesh._torch()
  dev = torch.device('cuda', ctx.device)
  out = []
  for a in arrays:
    if mesh._is_tensor(a):
      t = a.to(dev, dtype=torch.float32).contiguous()
    else:
      t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
    out.append(t)
  shape = tuple(out[0].shape)
  if len(shape) != 4 or any(tuple(t.shape) != shape for t in out):
    raise ValueError('x, cx and cy must be [2 or 3, z, y, x] arrays of the same shape')
  return out


def _tile_shape(shape, ncomp: int) -> _native.MeshShape:
  if shape[0] != ncomp:
    raise ValueError(f'expected {ncomp} components, got an array of shape {shape}')
  return _native.MeshShape(ncomp, 1, shape[1], shape[2], shape[3], 0)


def _tile_force(x, cx, cy, ncomp: int):
  dev = x.device.index if mesh._is_tensor(x) and x.is_cuda else None
  ctx = _native.Context.get(dev)
  xd, cxd, cyd = _tile_arrays(ctx, x, cx, cy)
  out = mesh._torch().empty_like(xd)
  shape = _tile_shape(tuple(xd.shape), ncomp)
  ctx.bind_stream()
  rc = _native.lib().sofima_tile_mesh_force(ctx.handle, xd.data_ptr(), cxd.data_ptr(),
                                            cyd.data_ptr(), ctypes.byref(shape),
                                            out.data_ptr())
  _native.check(ctx.handle, rc)
  return out if mesh._is_tensor(x) else out.cpu().numpy()


def elastic_tile_mesh(x, cx, cy, k=None, stride=None, prefer_orig_order=False, links=None):
  """Force on the nodes of a 2-d tile mesh (stitch_rigid.py:330-391).

  Args:
    x: [2, z, y, x] mesh where every node represents a tile
    cx: desired XY offsets between (x, y) and (x+1, y) tiles, same shape
    cy: desired XY offsets between (x, y) and (x, y+1) tiles, same shape
    k, stride, prefer_orig_order, links: unused (mesh solver compatibility)

  Returns:
    force field acting on the mesh, same shape as x
  """
  del k, stride, prefer_orig_order, links
  return _tile_force(x, cx, cy, 2)


def elastic_tile_mesh_3d(x, cx, cy, k=None, stride=None, prefer_orig_order=False, links=None):
  """Force on the nodes of a 3-d tile mesh, XYZ offsets (stitch_rigid.py:394-473)."""
  del k, stride, prefer_orig_order, links
  return _tile_force(x, cx, cy, 3)


def default_coarse_mesh_config() -> mesh.IntegrationConfig:
  """The settings optimize_coarse_mesh falls back to (stitch_rigid.py:496-507)."""
  return mesh.IntegrationConfig(dt=0.001, gamma=0.0, k0=0.0, k=0.1, stride=(1, 1),
                                num_iters=1000, max_iters=100000, stop_v_max=0.001,
                                dt_max=100)


def optimize_coarse_mesh(cx, cy, cfg: mesh.IntegrationConfig | None = None,
                         mesh_fn=elastic_tile_mesh) -> np.ndarray:
  """Computes rough initial positions of the tiles (stitch_rigid.py:476-545).

  Args:
    cx: desired XY[Z] offsets between (x, y) and (x+1, y) tiles, [2 or 3, 1, y, x]
    cy: desired XY[Z] offsets between (x, y) and (x, y+1) tiles
    cfg: integration config; None = default_coarse_mesh_config()
    mesh_fn: `elastic_tile_mesh` or `elastic_tile_mesh_3d` of this module

  Returns:
    optimized tile positions relative to the regular no-overlap grid, shaped like cx
  """
  if mesh_fn is elastic_tile_mesh:
    ncomp = 2
  elif mesh_fn is elastic_tile_mesh_3d:
    ncomp = 3
  else:
    raise NotImplementedError(
        'The CUDA backend runs the built-in tile-mesh force fields only '
        '(sofima_b200.stitch_rigid.elastic_tile_mesh / elastic_tile_mesh_3d); arbitrary '
        f'Python callables such as {mesh_fn!r} cannot be traced into the kernel.')
  if cfg is None:
    cfg = default_coarse_mesh_config()
  if cfg.start_cap != cfg.final_cap:  # relax_mesh's own argument checks, mesh.py:556-568
    if not cfg.fire:
      raise NotImplementedError('Adaptive force capping is only supported with FIRE.')
    if cfg.cap_scale <= 1:
      raise ValueError('The scaling factor for the force cap has to be larger '
                       'than 1 when the initial and final cap are different.')

  ctx = _native.Context.get()
  cxd, cyd = _tile_arrays(ctx, cx, cy)
  torch = mesh._torch()
  x = torch.zeros_like(cxd)  # all zeros = the regular grid layout with no overlap
  v = torch.zeros_like(cxd)
  a = torch.empty_like(cxd)
  shape = _tile_shape(tuple(cxd.shape), ncomp)
  # k0, k and stride are unused by the tile force; the stride only has to be well formed
  pod = mesh._config_pod(dataclasses.replace(cfg, stride=(1, 1)), mesh._INPLANE)
  dt, alpha, cap = np.float32(cfg.dt), np.float32(cfg.alpha), np.float32(cfg.start_cap)
  t = 0
  lib = _native.lib()
  while t < cfg.max_iters:
    c_dt, c_alpha, c_cap = (ctypes.c_float(float(dt)), ctypes.c_float(float(alpha)),
                            ctypes.c_float(float(cap)))
    n_pos, e_kin, v_max = ctypes.c_int32(0), ctypes.c_double(0), ctypes.c_float(0)
    ctx.bind_stream()
    rc = lib.sofima_tile_mesh_chunk(
        ctx.handle, x.data_ptr(), v.data_ptr(), a.data_ptr(), cxd.data_ptr(), cyd.data_ptr(),
        ctypes.byref(shape), ctypes.byref(pod), ctypes.byref(c_dt), ctypes.byref(c_alpha),
        ctypes.byref(c_cap), ctypes.byref(n_pos), ctypes.byref(e_kin), ctypes.byref(v_max))
    _native.check(ctx.handle, rc)
    t += cfg.num_iters
    if cfg.fire:
      dt, alpha, cap = (np.float32(c_dt.value), np.float32(c_alpha.value),
                        np.float32(c_cap.value))
    if np.float32(v_max.value) < np.float32(cfg.stop_v_max):
      if np.float32(cap) >= np.float32(cfg.final_cap):
        break
      cap = min(np.float32(cap) * np.float32(cfg.cap_scale), np.float32(cfg.final_cap))
  return x.cpu().numpy()
