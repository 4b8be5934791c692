"""NumPy fp32 restatement of the stitching target-mesh path (TEST ORACLE).

Follows /root/reference:
  * map_utils.compose_maps_fast        map_utils.py:616-734
  * stitch_elastic._apply_flow         stitch_elastic.py:456-570
  * stitch_elastic._update_mesh        stitch_elastic.py:573-620
  * stitch_elastic.compute_target_mesh stitch_elastic.py:624-676
and the JAX primitive they rest on, `jax.scipy.ndimage.map_coordinates(order=1)`
(third-party: jax, setup.cfg:27 `jax>=0.2.25`, unpinned, not installable here).
Its published algorithm (jax/_src/scipy/ndimage.py) is restated in
`map_coordinates_linear`: per axis `lower = floor(c)`, weights
`(1 - (c - lower), c - lower)`, corner indices `(lower, lower + 1)`; the corners
are visited in `itertools.product` order (last axis fastest), each contributes
`prod(weights) * value` with the weights multiplied left to right, the
contributions are added left to right.  mode='nearest' clamps the corner indices;
mode='constant' substitutes `cval` for a corner with ANY index out of range, also
when its weight is zero -- so with cval = NaN a query touching the last row/column
is NaN.  JAX runs with x64 disabled: every array op is fp32, integer index
arithmetic is int32, Python scalars are weakly typed.

Pinning: golden vectors from the reference's own source executed on the NumPy
shim (tests/golden/make_golden.py -> tests/golden/stitch_golden.npz) and the
reference KATs tests/map_utils_test.py:266-321.  JAX itself never ran: see the
caveat in DESIGN.md ("Oracle").

Test infrastructure only: never imported from `sofima_b200/`.
"""

from __future__ import annotations

import itertools

import numpy as np

F32 = np.float32


def map_coordinates_linear(inp, coords, mode, cval=np.nan):
  """jax.scipy.ndimage.map_coordinates(inp, coords, order=1, mode, cval), fp32."""
  inp = np.asarray(inp, dtype=F32)
  cval = F32(cval)
  per_axis = []
  with np.errstate(invalid='ignore'):
    for c, size in zip(coords, inp.shape):
      c = np.asarray(c, dtype=F32)
      lower = np.floor(c)
      upper_w = (c - lower).astype(F32)
      lower_w = (F32(1.0) - upper_w).astype(F32)
      # XLA's f32 -> s32 conversion saturates; NaN -> any value (the weights are NaN
      # then and poison the result anyway).
      li = np.where(np.isfinite(lower), np.clip(lower, -2.0**31, 2.0**31 - 1), 0)
      index = li.astype(np.int64)
      nodes = []
      for idx, w in ((index, lower_w), (index + 1, upper_w)):
        if mode == 'constant':
          valid = (idx >= 0) & (idx < size)
          fixed = np.clip(idx, 0, size - 1)
        elif mode == 'nearest':
          valid = True
          fixed = np.clip(idx, 0, size - 1)
        else:
          raise NotImplementedError(mode)
        nodes.append((fixed, valid, w))
      per_axis.append(nodes)

    result = None
    for items in itertools.product(*per_axis):
      idxs = tuple(it[0] for it in items)
      valids = [it[1] for it in items]
      weights = [it[2] for it in items]
      val = inp[idxs]
      if not all(v is True for v in valids):
        allv = valids[0]
        for v in valids[1:]:
          allv = allv & v
        val = np.where(allv, val, cval).astype(F32)
      w = weights[0]
      for ww in weights[1:]:
        w = (w * ww).astype(F32)
      contrib = (w * val).astype(F32)
      result = contrib if result is None else (result + contrib).astype(F32)
  return result


def _scalar_mul(ref_int, stride):
  """int32 grid * stride: int32 for an integer stride, fp32 otherwise."""
  if isinstance(stride, (int, np.integer)):
    return (ref_int * int(stride)).astype(F32)
  return (ref_int.astype(F32) * F32(stride)).astype(F32)


def compose_maps_fast(map1, start1, stride1, map2, start2, stride2, mode='nearest'):
  """map_utils.py:616-734."""
  map1 = np.asarray(map1, dtype=F32)
  map2 = np.asarray(map2, dtype=F32)
  assert map1.shape[0] == map2.shape[0]
  dim = map1.shape[0]
  if not isinstance(stride1, (tuple, list, np.ndarray)):
    stride1 = (stride1,) * dim
  if not isinstance(stride2, (tuple, list, np.ndarray)):
    stride2 = (stride2,) * dim
  start1 = np.asarray(start1).astype(np.int64)
  start2 = np.asarray(start2).astype(np.int64)
  origin = np.minimum(start1, start2)

  def ref_grid(coord_map, start, stride):
    start = (start - origin)[-dim:]
    ranges = [np.arange(coord_map.shape[4 - dim + i]) + start[i] for i in range(dim)]
    ref = np.meshgrid(*ranges, indexing='ij')
    return [_scalar_mul(a, b) for a, b in zip(ref, stride)]

  ref1 = ref_grid(map1, start1, stride1)
  ref2 = ref_grid(map2, start2, stride2)

  def div(a, s):
    return (a / F32(s)).astype(F32)

  if dim == 2:
    ret = np.zeros_like(map1)
    for z in range(map1.shape[1]):
      qx = div((ref1[-1] + map1[0, z]).astype(F32), stride2[-1])
      qy = div((ref1[-2] + map1[1, z]).astype(F32), stride2[-2])
      xx = map_coordinates_linear((map2[0, z] + ref2[-1]).astype(F32), [qy, qx], mode)
      yy = map_coordinates_linear((map2[1, z] + ref2[-2]).astype(F32), [qy, qx], mode)
      ret[0, z] = (xx - ref1[-1]).astype(F32)
      ret[1, z] = (yy - ref1[-2]).astype(F32)
    return ret
  qx = div((ref1[-1] + map1[0]).astype(F32), stride2[-1])
  qy = div((ref1[-2] + map1[1]).astype(F32), stride2[-2])
  qz = div((ref1[-3] + map1[2]).astype(F32), stride2[-3])
  q = [qz, qy, qx]
  xx = map_coordinates_linear((map2[0] + ref2[-1]).astype(F32), q, mode)
  yy = map_coordinates_linear((map2[1] + ref2[-2]).astype(F32), q, mode)
  zz = map_coordinates_linear((map2[2] + ref2[-3]).astype(F32), q, mode)
  return np.stack([(xx - ref1[-1]).astype(F32), (yy - ref1[-2]).astype(F32),
                   (zz - ref1[-3]).astype(F32)])


# NeighborInfo indices (stitch_elastic.py:43-72).
NBOR_IDX, FLOW_IDX, COARSE_ORTHO, FLOW_ORTHO, FLOW_OVERLAP = 0, 1, 2, 3, 4
FINE_X, FINE_Y, DIM, COARSE_Z, FLOW_Z, FINE_Z = 5, 6, 7, 8, 9, 10


def _clamped_start(start, sizes, shape):
  """lax.dynamic_slice / dynamic_update_slice clamp the start so the slice fits."""
  return [int(min(max(int(s), 0), n - int(sz))) for s, sz, n in zip(start, sizes, shape)]


def _apply_flow(base_mesh, nbor_mesh, nbor_flow_all, mult, stride, nd, dim):
  """stitch_elastic.py:456-570.  base_mesh is updated in place and returned."""
  flow_overlap, flow_ortho, offset_ortho = int(nd[FLOW_OVERLAP]), int(nd[FLOW_ORTHO]), int(
      nd[COARSE_ORTHO])
  par_size = nbor_mesh.shape[-dim - 1]
  ortho_size = nbor_mesh.shape[dim - 2]
  start_par = par_size - flow_overlap if mult == 1 else 0
  start_ortho = (ortho_size - flow_ortho) if (
      (mult == 1 and offset_ortho > 0) or (mult == -1 and offset_ortho < 0)) else 0
  start = [start_ortho * (1 - dim) + dim * start_par,
           start_ortho * dim + (1 - dim) * start_par]
  nbor_flow = (F32(mult) * nbor_flow_all[:, int(nd[FLOW_IDX])]).astype(F32)
  three_d = base_mesh.shape[0] == 3
  if three_d:
    offset_z, flow_z = int(nd[COARSE_Z]), int(nd[FLOW_Z])
    start_z = (nbor_mesh.shape[-3] - flow_z) if (
        (mult == 1 and offset_z > 0) or (mult == -1 and offset_z < 0)) else 0
    start = [start_z] + start
    flow3, mesh3 = nbor_flow, nbor_mesh
  else:
    flow3, mesh3 = nbor_flow[:, None], nbor_mesh[:, None]
  update = compose_maps_fast(flow3, start, stride, mesh3, [0] * len(start), stride,
                             mode='constant')
  if not three_d:
    update = update[:, 0]
    off = np.array([nd[FINE_X], nd[FINE_Y]]).reshape(2, 1, 1)
  else:
    off = np.array([nd[FINE_X], nd[FINE_Y], nd[FINE_Z]]).reshape(3, 1, 1, 1)
  update = (update + (mult * off).astype(F32)).astype(F32)

  tg_par = 0 if mult == 1 else par_size - flow_overlap
  tg_ortho = (ortho_size - flow_ortho) if (
      (mult == 1 and offset_ortho < 0) or (mult == -1 and offset_ortho > 0)) else 0
  tg = [0, tg_par * dim + (1 - dim) * tg_ortho, tg_par * (1 - dim) + dim * tg_ortho]
  if three_d:
    tg_z = (nbor_mesh.shape[-3] - flow_z) if (
        (mult == 1 and offset_z < 0) or (mult == -1 and offset_z > 0)) else 0
    tg = [0, tg_z] + tg[1:]
  tg = _clamped_start(tg, nbor_flow.shape, base_mesh.shape)
  sel = tuple(slice(s, s + n) for s, n in zip(tg, nbor_flow.shape))
  previous = base_mesh[sel]
  base_mesh[sel] = np.where(np.isnan(update), previous, update)
  return base_mesh


def compute_target_mesh(nbor_data, x, fx, fy, stride=(20, 20)):
  """stitch_elastic.py:624-676 (one tile: nbor_data is [4, 8 or 11])."""
  x = np.asarray(x, dtype=F32)
  fx = np.asarray(fx, dtype=F32)
  fy = np.asarray(fy, dtype=F32)
  dim = x.shape[0]
  zyx = list(x.shape[-dim:])
  for i in range(dim):
    zyx[i] += max(fy.shape[-dim + i], fx.shape[-dim + i])
  mesh = np.full([dim] + zyx, np.nan, dtype=F32)
  for nd in np.asarray(nbor_data):
    nbor_idx, flow_idx = int(nd[NBOR_IDX]), int(nd[FLOW_IDX])
    if nbor_idx == -1:
      continue
    mult = 1 if nbor_idx == flow_idx else -1
    nbor_mesh = x[:, nbor_idx]
    if int(nd[DIM]) == 0:
      mesh = _apply_flow(mesh, nbor_mesh, fx, mult, stride, nd, 0)
    else:
      mesh = _apply_flow(mesh, nbor_mesh, fy, mult, stride, nd, 1)
  sel = (slice(None),) + tuple(slice(0, n) for n in x.shape[-dim:])
  return mesh[sel]


def target_mesh_all(nbors, x, fx, fy, stride):
  """The notebooks' prev_fn: vmap(compute_target_mesh)(nbors) transposed to
  [dim, n, ...] (notebooks/em_stitching.ipynb:545-549)."""
  out = np.stack([compute_target_mesh(nd, x, fx, fy, stride) for nd in nbors])
  return np.moveaxis(out, 0, 1)


# ------------------------------------------------------------------------------------
# Rigid tile-grid step (stitch_rigid.py:277-545): every tile is one node of a small mesh
# whose springs want the measured coarse offsets between neighbouring tiles.
# ------------------------------------------------------------------------------------
def _tile_links(x, cx, cy, terms):
  """Sum of the spring terms of stitch_rigid.elastic_tile_mesh[_3d] in the reference's
  order.  A term (comp, axis, target) is: f = nan_to_num(x[comp, next] - x[comp, this] -
  target), added to `this` and subtracted from `next` along `axis` (-1: x neighbour,
  -2: y neighbour); the reference adds zero-padded copies of f to the whole array."""
  x = np.asarray(x, dtype=F32)
  tgt = {'cx': np.asarray(cx, dtype=F32), 'cy': np.asarray(cy, dtype=F32)}
  f_tot = np.zeros_like(x)
  for comp, axis, which in terms:
    this = [slice(None)] * 3
    nxt = [slice(None)] * 3
    this[axis] = slice(None, -1)
    nxt[axis] = slice(1, None)
    this, nxt = tuple(this), tuple(nxt)
    d = x[comp][nxt] - x[comp][this]
    f = np.nan_to_num(d - tgt[which][comp][this]).astype(F32)
    f_tot[comp][this] = f_tot[comp][this] + f
    f_tot[comp][nxt] = f_tot[comp][nxt] - f
  return f_tot


def elastic_tile_mesh(x, cx, cy, k=None, stride=None, prefer_orig_order=False, links=None):
  """stitch_rigid.elastic_tile_mesh (stitch_rigid.py:330-391); x: [2, z, y, x]."""
  del k, stride, prefer_orig_order, links
  return _tile_links(x, cx, cy, ((0, -1, 'cx'), (1, -2, 'cy'), (0, -2, 'cy'), (1, -1, 'cx')))


def elastic_tile_mesh_3d(x, cx, cy, k=None, stride=None, prefer_orig_order=False, links=None):
  """stitch_rigid.elastic_tile_mesh_3d (stitch_rigid.py:394-473); x: [3, z, y, x]."""
  del k, stride, prefer_orig_order, links
  return _tile_links(x, cx, cy, ((0, -1, 'cx'), (1, -2, 'cy'), (0, -2, 'cy'), (1, -1, 'cx'),
                                 (2, -1, 'cx'), (2, -2, 'cy')))


def optimize_coarse_mesh(cx, cy, cfg=None, mesh_fn=elastic_tile_mesh):
  """stitch_rigid.optimize_coarse_mesh (stitch_rigid.py:476-545): relaxes the all-zero tile
  mesh under `mesh_fn`; returns the tile positions relative to the regular grid."""
  from oracle import mesh_oracle
  if cfg is None:
    from sofima_b200.mesh import IntegrationConfig  # the dataclass only
    cfg = IntegrationConfig(dt=0.001, gamma=0.0, k0=0.0, k=0.1, stride=(1, 1),
                            num_iters=1000, max_iters=100000, stop_v_max=0.001, dt_max=100)
  force = lambda x, k, stride, prefer_orig_order=False: mesh_fn(x, cx, cy)
  res = mesh_oracle.relax_mesh(np.zeros(np.shape(cx), F32), None, cfg, mesh_force=force)
  return np.array(res[0])


def interpolate_missing_offsets(conn, axis, max_r=4):
  """stitch_rigid.interpolate_missing_offsets (stitch_rigid.py:277-327): an offset marked
  inf is replaced by the mean of its nearest finite neighbours at the smallest distance
  1 <= r < max_r along `axis` (-1: x, -2: y); in place, returns conn."""
  if conn.ndim != 4:
    raise ValueError('conn array must have rank 4')
  n = conn.shape[axis]
  for y, x in zip(*np.where(np.isinf(conn[0, 0]))):
    pos = (y, x)[axis + 2]
    for r in range(1, max_r):
      found = []
      for q in (pos - r, pos + r):
        if 0 <= q < n:
          idx = [0, 0, y, x]
          idx[axis] = q
          if np.isfinite(conn[tuple(idx)]):
            idx[0] = slice(None)
            found.append(conn[tuple(idx)])
      if found:
        conn[:, 0, y, x] = np.mean(found, axis=0)
        break
  return conn
