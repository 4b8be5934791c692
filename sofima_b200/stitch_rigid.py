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
its own linear spring force; the CPU restatement used by the tests is pinned on the
reference's run (tests/golden/coarse_golden.npz) but the device kernel is not built yet,
so the function raises NotImplementedError here.
"""

from __future__ import annotations

from typing import Mapping, Sequence

import numpy as np
from scipy import ndimage

from . import flow_field

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


def optimize_coarse_mesh(cx, cy, cfg=None, mesh_fn=None):
  """Rough initial tile positions from the coarse offsets (stitch_rigid.py:476-545).

  Not built on the device yet: the one-node-per-tile relaxation uses its own force
  field (`elastic_tile_mesh[_3d]`), which the mesh kernels do not evaluate.
  """
  del cx, cy, cfg, mesh_fn
  raise NotImplementedError(
      'optimize_coarse_mesh: the tile-grid force field is not part of the CUDA backend yet '
      '(DESIGN.md section 7)')
