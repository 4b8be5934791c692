"""BASELINE config 4 in small (3-d EM section alignment, notebooks/em_alignment.ipynb):
section-to-section flow -> clean_flow -> reconcile_flows -> compose with the solved previous
section -> relax_mesh -> ndimage_warp, the whole chain on the CUDA path against the same
chain on the oracle.  Flow vectors are integers, the mesh and warp kernels are bit-faithful,
so the final maps and the warped sections must be IDENTICAL; the two statistics channels
only pass through thresholds (the inputs keep them away from the decision boundaries, which
the CPU-only twin of this test, tests/test_oracle_pipeline.py, asserts)."""

import numpy as np
import pytest
import scipy.ndimage as ndi

from oracle import flow_oracle, mesh_oracle, stitch_oracle, warp_oracle

PATCH, STEP, STRIDE = 160, 40, (40.0, 40.0)
CLEAN = dict(min_peak_ratio=1.6, min_peak_sharpness=1.6, max_magnitude=40, max_deviation=10)


def make_stack(nz=4, size=560, seed=17):
  """Sections cut from one texture with a cumulative integer drift, independent noise and
  one blanked region (no texture -> no peak -> NaN flow vectors)."""
  rng = np.random.default_rng(seed)
  base = ndi.gaussian_filter(rng.standard_normal((size + 64, size + 64)), 2.0)
  base = (base - base.min()) / np.ptp(base) * 255
  drift = [(0, 0), (3, -2), (5, 1), (2, 4)][:nz]
  stack = []
  for z, (dy, dx) in enumerate(drift):
    sec = base[32 + dy:32 + dy + size, 32 + dx:32 + dx + size]
    sec = np.clip(sec + rng.normal(0, 4, sec.shape), 0, 255).astype(np.uint8)
    if z == 2:
      sec[200:400, 120:330] = 0
    stack.append(sec)
  return np.stack(stack)


def align(stack, calc, clean_flow, reconcile_flows, compose, relax, warp_fn, cfg):
  """em_alignment.ipynb in small: returns (flows, cleaned flow, solved meshes, warped)."""
  nz = stack.shape[0]
  flows = [calc.flow_field(stack[z - 1], stack[z], PATCH, STEP, batch_size=64)
           for z in range(1, nz)]
  flow = np.stack(flows, axis=1)  # [4, nz - 1, gy, gx]
  clean = clean_flow(flow, **CLEAN)
  clean = reconcile_flows([clean], max_gradient=0, max_deviation=0, min_patch_size=4)
  solved = [np.zeros((2, 1) + clean.shape[2:], np.float32)]
  warped = [stack[0]]
  for z in range(1, nz):
    prev = compose(clean[:, z - 1:z].astype(np.float32), (z, 0, 0), STRIDE, solved[-1],
                   (z - 1, 0, 0), STRIDE)
    x = relax(np.zeros_like(solved[0]), prev, cfg)[0]
    solved.append(np.asarray(x))
    # The notebook inverts the solved map (map_utils.invert_map: CPU Delaunay resampling,
    # out of scope) before warping; for the near-uniform drift of this stack the inverse
    # is -x to first order.
    warped.append(warp_fn(stack[z], -np.asarray(x)[:, 0].astype(np.float64), STRIDE))
  return flow, clean, np.stack(solved), np.stack(warped)


def config():
  from sofima_b200.mesh import IntegrationConfig  # the dataclass itself needs no GPU
  return IntegrationConfig(
      dt=0.001, gamma=0.0, k0=0.01, k=0.1, stride=STRIDE, num_iters=50, max_iters=300,
      stop_v_max=0.005, dt_max=1000, prefer_orig_order=True, start_cap=0.01, final_cap=10.0)


def oracle_chain(stack):
  from sofima_b200 import flow_utils
  return align(stack, flow_oracle.MaskedXCorrWithStatsCalculator(), flow_utils.clean_flow,
               flow_utils.reconcile_flows, stitch_oracle.compose_maps_fast,
               mesh_oracle.relax_mesh,
               lambda img, m, st: warp_oracle.ndimage_warp(img, m, st), config())


@pytest.mark.gpu
def test_config4_alignment_chain():
  import torch
  if not torch.cuda.is_available():
    pytest.skip('needs a CUDA device')
  from sofima_b200 import flow_field, flow_utils, map_utils, mesh, warp
  stack = make_stack()
  got = align(stack, flow_field.JAXMaskedXCorrWithStatsCalculator(), flow_utils.clean_flow,
              flow_utils.reconcile_flows, map_utils.compose_maps_fast, mesh.relax_mesh,
              lambda img, m, st: warp.ndimage_warp(img, m, st), config())
  want = oracle_chain(stack)
  flow, flow_w = got[0], want[0]
  np.testing.assert_array_equal(np.isnan(flow), np.isnan(flow_w))
  np.testing.assert_array_equal(flow[:2], flow_w[:2])
  np.testing.assert_allclose(flow[2:], flow_w[2:], rtol=2e-3, atol=1e-6)
  np.testing.assert_array_equal(got[1], want[1])   # cleaned + reconciled flow
  np.testing.assert_array_equal(got[2], want[2])   # solved meshes, bit for bit
  np.testing.assert_array_equal(got[3], want[3])   # warped sections
  assert np.isnan(got[1]).any() and np.abs(got[2]).max() > 1.0


# ------------------------------------------------------------------------------------
# BASELINE config 5 in small (LICONN in-plane stitching of 3-d tiles,
# notebooks/liconn_inplane_stitching.ipynb cells 27-35): compute_flow_map3d ->
# clean_flow / reconcile_flows (dim 3) -> aggregate_arrays -> relax_mesh with the
# stitching prev_fn and elastic_mesh_3d -> ndimage_warp of every tile through its mesh.
# ------------------------------------------------------------------------------------
PATCH3, STRIDE3 = (12, 16, 16), (4, 8, 8)
CLEAN3 = dict(min_peak_ratio=1.4, min_peak_sharpness=1.4, max_deviation=5, max_magnitude=0,
              dim=3)


def liconn_inputs():
  import os
  g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'flow3d_golden.npz'))
  vol = g['fm3_vol']
  tz, th, tw = (int(v) for v in g['fm3_tile_zyx'])
  tiles = {(int(a), int(b)): np.ascontiguousarray(vol[z0:z0 + tz, y0:y0 + th, x0:x0 + tw])[None]
           for a, b, z0, y0, x0 in g['fm3_nominal']}
  return tiles, (tw, th, tz), g['fm3_cx'], g['fm3_cy']


def stitch3d(stitch_elastic, relax, target_fn, force, warp_fn):
  """stitch_elastic supplies the host logic (the flow calculator inside it is whatever
  flow_field.JAXMaskedXCorrWithStatsCalculator currently is)."""
  from sofima_b200 import flow_utils
  from sofima_b200.mesh import IntegrationConfig
  tiles, size_xyz, cx, cy = liconn_inputs()
  fx, ox = stitch_elastic.compute_flow_map3d(tiles, size_xyz, cx, 0, PATCH3, STRIDE3, 16)
  fy, oy = stitch_elastic.compute_flow_map3d(tiles, size_xyz, cy, 1, PATCH3, STRIDE3, 16)
  rec = dict(min_patch_size=10, max_gradient=-1, max_deviation=-1)
  fine_x = {k: flow_utils.reconcile_flows([flow_utils.clean_flow(v, **CLEAN3)], **rec)
            for k, v in fx.items()}
  fine_y = {k: flow_utils.reconcile_flows([flow_utils.clean_flow(v, **CLEAN3)], **rec)
            for k, v in fy.items()}
  afx, afy, x, nbors, key_to_idx = stitch_elastic.aggregate_arrays(
      (cx[:, 0], fine_x, ox), (cy[:, 0], fine_y, oy), list(tiles.keys()),
      np.zeros((3, 2, 2)), STRIDE3, size_xyz[::-1])
  afx, afy, x = afx.astype(np.float32), afy.astype(np.float32), x.astype(np.float32)
  cfg = IntegrationConfig(dt=0.001, gamma=0., k0=0.1, k=0.1, stride=STRIDE3[::-1],
                          num_iters=50, max_iters=300, stop_v_max=0.0, dt_max=100,
                          prefer_orig_order=False, start_cap=0.1, final_cap=10.,
                          remove_drift=False)
  xr, _, t = relax(x, None, cfg, prev_fn=target_fn(nbors, afx, afy, STRIDE3), mesh_force=force)
  xr = np.asarray(xr)
  warped = {k: warp_fn(tiles[k][0], xr[:, i].astype(np.float64), STRIDE3)
            for k, i in key_to_idx.items()}
  return (fx, fy), (fine_x, fine_y), (afx, afy, x, nbors), xr, t, warped


def oracle_stitch3d(monkeypatch):
  from sofima_b200 import flow_field, stitch_elastic
  monkeypatch.setattr(flow_field, 'JAXMaskedXCorrWithStatsCalculator',
                      flow_oracle.MaskedXCorrWithStatsCalculator)
  target = lambda nb, a, b, st: (lambda x: stitch_oracle.target_mesh_all(nb, x, a, b, st))
  out = stitch3d(stitch_elastic, mesh_oracle.relax_mesh, target, mesh_oracle.elastic_mesh_3d,
                 lambda img, m, st: warp_oracle.ndimage_warp(img, m, st))
  monkeypatch.undo()
  return out


@pytest.mark.gpu
def test_config5_liconn_chain(monkeypatch):
  import torch
  if not torch.cuda.is_available():
    pytest.skip('needs a CUDA device')
  from sofima_b200 import mesh, stitch_elastic, warp
  got = stitch3d(stitch_elastic, mesh.relax_mesh, stitch_elastic.target_mesh_fn,
                 mesh.elastic_mesh_3d, lambda img, m, st: warp.ndimage_warp(img, m, st))
  want = oracle_stitch3d(monkeypatch)
  for a, b in zip(got[0] + got[1], want[0] + want[1]):  # raw and filtered flow maps
    assert a.keys() == b.keys()
    for k in a:
      np.testing.assert_array_equal(np.isnan(a[k]), np.isnan(b[k]))
      np.testing.assert_array_equal(a[k][:3], b[k][:3])
  for a, b in zip(got[2], want[2]):                     # aggregate_arrays
    np.testing.assert_array_equal(a, b)
  assert got[4] == want[4] == 300
  np.testing.assert_array_equal(got[3], want[3])        # relaxed tile meshes, bit for bit
  for k in want[5]:
    np.testing.assert_array_equal(got[5][k], want[5][k])  # warped tiles
  assert np.abs(got[3]).max() > 0.5
