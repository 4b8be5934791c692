"""Oracle of the stitching target-mesh path against the reference.

* golden vectors: tests/golden/stitch_golden.npz, produced by the reference's own
  map_utils.compose_maps_fast / stitch_elastic.{aggregate_arrays, compute_target_mesh}
  (tests/golden/make_golden.py);
* the reference's KAT tests/map_utils_test.py:266-300.
CPU only.
"""

import os

import numpy as np
import pytest

from oracle import stitch_oracle as so

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'stitch_golden.npz')


@pytest.fixture(scope='module')
def g():
  return np.load(GOLDEN)


@pytest.mark.parametrize('mode', ['nearest', 'constant'])
def test_compose_maps_fast_golden(g, mode):
  got = so.compose_maps_fast(g['cmf_map1'], (3, 2, 1), (20, 25), g['cmf_map2'], (1, 4, 0),
                             (16, 20), mode=mode)
  np.testing.assert_array_equal(got, g[f'cmf_{mode}'])
  got = so.compose_maps_fast(g['cmf3_map1'], (1, 0, 2), (4, 10, 10), g['cmf3_map2'],
                             (0, 1, 1), (4, 10, 10), mode=mode)
  np.testing.assert_array_equal(got, g[f'cmf3_{mode}'])


def test_compose_maps_fast_reference_kat():
  # tests/map_utils_test.py:266-300 (BoundingBox starts are xyz; reversed = zyx)
  coord_map = np.zeros([2, 1, 60, 60])
  flow = np.zeros([2, 1, 50, 50])
  flow[0, 0, :, 10:25] = -5
  flow[0, 0, :, 25:40] = 65
  flow[:, 0, :, 4] = np.nan
  s1, s2 = (64, 58, 42), (64, 50, 40)
  updated = so.compose_maps_fast(flow, s1, 40, coord_map, s2, 40)
  np.testing.assert_array_equal(updated, flow)
  coord_map[0, :, :, 7:] = -10
  updated = so.compose_maps_fast(flow, s1, 40, coord_map, s2, 40)
  flow[0, 0, :, 5:10] = -10
  flow[0, 0, :, 10:25] = -15
  flow[0, 0, :, 25:40] = 55
  flow[0, 0, :, 40:] = -10
  np.testing.assert_array_equal(updated, flow)


def test_target_mesh_golden_2d(g):
  got = so.target_mesh_all(g['st2_nbors'], g['st2_x'], g['st2_fx'], g['st2_fy'],
                           tuple(int(v) for v in g['st2_stride']))
  np.testing.assert_array_equal(got, g['st2_target'])
  assert 0.2 < np.isfinite(got).mean() < 0.8  # targets exist in the overlaps only


def test_target_mesh_golden_3d(g):
  got = so.target_mesh_all(g['st3_nbors'], g['st3_x'], g['st3_fx'], g['st3_fy'],
                           tuple(int(v) for v in g['st3_stride']))
  np.testing.assert_array_equal(got, g['st3_target'])


def _aggregate_inputs(g):
  coords = [tuple(int(v) for v in c) for c in g['st2_coords']]
  fine = {'x': {}, 'y': {}}
  offs = {'x': {}, 'y': {}}
  for key in g.files:
    for nm in ('x', 'y'):
      if key.startswith(f'st2_fine{nm}_'):
        tx, ty = (int(v) for v in key.split('_')[2:])
        fine[nm][tx, ty] = g[key]
        offs[nm][tx, ty] = tuple(int(v) for v in g[f'st2_off{nm}_{tx}_{ty}'])
  return coords, fine, offs


def test_aggregate_arrays_matches_reference(g):
  # host NumPy glue of the product (no GPU needed): same tables as the reference's.
  from sofima_b200 import stitch_elastic
  coords, fine, offs = _aggregate_inputs(g)
  fx, fy, x, nbors, key_to_idx = stitch_elastic.aggregate_arrays(
      (g['st2_cx'], fine['x'], offs['x']), (g['st2_cy'], fine['y'], offs['y']), coords,
      np.zeros((2, 2, 3)), (20, 20), (240, 280))
  np.testing.assert_array_equal(nbors, g['st2_nbors'])
  np.testing.assert_array_equal(fx.astype(np.float32), g['st2_fx'])
  np.testing.assert_array_equal(fy.astype(np.float32), g['st2_fy'])
  np.testing.assert_array_equal(x, g['st2_x0'])
  assert key_to_idx == {c: i for i, c in enumerate(coords)}


@pytest.mark.parametrize('tag,atol', [('st2_relax', 2e-6), ('st2_relax_damped', 0.0)])
def test_relaxation_with_stitching_prev_fn_golden(g, tag, atol):
  # The whole notebooks/em_stitching.ipynb:545-603 solve, run by the reference's own
  # mesh.relax_mesh + compute_target_mesh.  remove_drift takes fp32 means whose
  # summation order is unspecified (oracle: fp64), hence 2e-6 there.
  import ast
  from oracle import mesh_oracle as mo
  from sofima_b200.mesh import IntegrationConfig
  stride = tuple(int(v) for v in g['st2_stride'])
  prev_fn = lambda x: so.target_mesh_all(g['st2_nbors'], x, g['st2_fx'], g['st2_fy'], stride)
  cfg = IntegrationConfig(**ast.literal_eval(str(g[f'{tag}_cfg'])))
  x, e_kin, t = mo.relax_mesh(g['st2_relax_x0'], None, cfg, prev_fn=prev_fn)
  assert t == 36
  np.testing.assert_allclose(x, g[f'{tag}_x'], rtol=0, atol=atol)
  np.testing.assert_allclose(e_kin, g[f'{tag}_ekin'], rtol=1e-6)


@pytest.mark.parametrize('tag,atol', [('st3_relax', 2e-6), ('st3_relax_nodrift', 0.0)])
def test_relaxation_3d_with_stitching_prev_fn_golden(g, tag, atol):
  # notebooks/liconn_inplane_stitching.ipynb:763-783 run by the reference's own code:
  # [3, tiles, z, y, x] meshes, elastic_mesh_3d.  With remove_drift the reference's
  # mean over axes (1, 2, 3) is one mean per x column (mesh.py:496-497).
  import ast
  from oracle import mesh_oracle as mo
  from sofima_b200.mesh import IntegrationConfig
  stride = tuple(int(v) for v in g['st3_stride'])
  prev_fn = lambda x: so.target_mesh_all(g['st3_nbors'], x, g['st3_fx'], g['st3_fy'], stride)
  cfg = IntegrationConfig(**ast.literal_eval(str(g[f'{tag}_cfg'])))
  x, e_kin, t = mo.relax_mesh(g['st3_x'], None, cfg, prev_fn=prev_fn,
                              mesh_force=mo.elastic_mesh_3d)
  assert t == 24
  np.testing.assert_allclose(x, g[f'{tag}_x'], rtol=0, atol=atol)
  np.testing.assert_allclose(e_kin, g[f'{tag}_ekin'], rtol=1e-6)


def test_compute_flow_map_host_logic(g, monkeypatch):
  # stitch_elastic.py:197-282: strip geometry / offsets / NaN padding of the product's
  # compute_flow_map against the reference's own run (golden); the flow calculator is
  # replaced by the oracle here (CPU), tests/test_stitch_gpu.py runs the CUDA one.
  from oracle import flow_oracle
  from sofima_b200 import flow_field, stitch_elastic
  monkeypatch.setattr(flow_field, 'JAXMaskedXCorrWithStatsCalculator',
                      flow_oracle.MaskedXCorrWithStatsCalculator)
  _check_flow_maps(g, stitch_elastic)


def _check_flow_maps(g, stitch_elastic):
  tex = g['fm2_tex']
  tiles = {(int(a), int(b)): np.ascontiguousarray(tex[y0:y0 + 160, x0:x0 + 200])
           for a, b, y0, x0 in g['fm2_nominal']}
  for axis, cm in ((0, g['fm2_cx']), (1, g['fm2_cy'])):
    flows, offsets = stitch_elastic.compute_flow_map(tiles, cm, axis, patch_size=(32, 32),
                                                     stride=(8, 8), batch_size=64)
    assert len(flows) == len([k for k in g.files if k.startswith(f'fm2_flow{axis}_')])
    for k, f in flows.items():
      want = g[f'fm2_flow{axis}_{k[0]}_{k[1]}']
      assert tuple(offsets[k]) == tuple(g[f'fm2_off{axis}_{k[0]}_{k[1]}'])
      assert f.shape == want.shape
      np.testing.assert_array_equal(np.isnan(f), np.isnan(want))
      np.testing.assert_array_equal(f[:2], want[:2])
      np.testing.assert_allclose(f[2:], want[2:], rtol=2e-3, atol=1e-6)


def _check_coarse_offsets(g, stitch_rigid):
  tex = g['fm2_tex']
  tiles = {(int(a), int(b)): np.ascontiguousarray(tex[y0:y0 + 160, x0:x0 + 200])
           for a, b, y0, x0 in g['fm2_nominal']}
  cx, cy = stitch_rigid.compute_coarse_offsets(
      (2, 3), tiles, overlaps_xy=((30, 44), (30, 44)), min_range=(10, 100, 0),
      min_overlap=16, filter_size=5)
  np.testing.assert_array_equal(cx, g['co_conn_x'])
  np.testing.assert_array_equal(cy, g['co_conn_y'])
  np.testing.assert_array_equal(cx[:, 0], g['fm2_cx'])  # = the true tile offsets
  np.testing.assert_array_equal(cy[:, 0], g['fm2_cy'])


def test_compute_coarse_offsets_host_logic(g, monkeypatch):
  # stitch_rigid.py:104-273 against the reference's own run (golden); the correlation
  # itself is the oracle's here, tests/test_stitch_gpu.py runs the CUDA one.
  from oracle import flow_oracle
  from sofima_b200 import flow_field, stitch_rigid
  monkeypatch.setattr(flow_field, 'JAXMaskedXCorrWithStatsCalculator',
                      flow_oracle.MaskedXCorrWithStatsCalculator)
  _check_coarse_offsets(g, stitch_rigid)


GOLDEN3D = os.path.join(os.path.dirname(__file__), 'golden', 'flow3d_golden.npz')


def _check_flow_maps3d(stitch_elastic):
  """stitch_elastic.compute_flow_map3d (stitch_elastic.py:84-193, the fine-flow step of
  notebooks/liconn_inplane_stitching.ipynb) against the reference's own run on 2 x 2
  three-dimensional tiles (tests/golden/make_golden.py flow3d)."""
  g = np.load(GOLDEN3D)
  vol = g['fm3_vol']
  tz, th, tw = (int(v) for v in g['fm3_tile_zyx'])
  tiles = {(int(a), int(b)): np.ascontiguousarray(vol[z0:z0 + tz, y0:y0 + th, x0:x0 + tw])[None]
           for a, b, z0, y0, x0 in g['fm3_nominal']}
  for axis, cm in ((0, g['fm3_cx']), (1, g['fm3_cy'])):
    flows, offsets = stitch_elastic.compute_flow_map3d(
        tiles, (tw, th, tz), cm, axis, patch_size=(12, 16, 16), stride=(4, 8, 8), batch_size=16)
    assert len(flows) == len([k for k in g.files if k.startswith(f'fm3_flow{axis}_')]) == 2
    for k, f in flows.items():
      want = g[f'fm3_flow{axis}_{k[0]}_{k[1]}']
      assert tuple(offsets[k]) == tuple(g[f'fm3_off{axis}_{k[0]}_{k[1]}'])
      assert f.shape == want.shape
      np.testing.assert_array_equal(np.isnan(f), np.isnan(want))
      np.testing.assert_array_equal(f[:3], want[:3])
      np.testing.assert_allclose(f[3:], want[3:], rtol=2e-3, atol=1e-6)


def test_compute_flow_map3d_host_logic(monkeypatch):
  from oracle import flow_oracle
  from sofima_b200 import flow_field, stitch_elastic
  monkeypatch.setattr(flow_field, 'JAXMaskedXCorrWithStatsCalculator',
                      flow_oracle.MaskedXCorrWithStatsCalculator)
  _check_flow_maps3d(stitch_elastic)


# ---- rigid tile-grid step (stitch_rigid.py:277-545) ---------------------------------
COARSE = os.path.join(os.path.dirname(__file__), 'golden', 'coarse_golden.npz')


def test_tile_mesh_forces_golden():
  g = np.load(COARSE)
  np.testing.assert_array_equal(
      so.elastic_tile_mesh(g['cm2_x'], g['cm2_cx'], g['cm2_cy']), g['cm2_force'])
  np.testing.assert_array_equal(
      so.elastic_tile_mesh_3d(g['cm3_x'], g['cm3_cx'], g['cm3_cy']), g['cm3_force'])


def test_optimize_coarse_mesh_golden():
  # the reference's own optimize_coarse_mesh (default config: FIRE until v_max < 1e-3,
  # and a fixed 300 steps) on a 3 x 4 tile grid, 2-d and 3-d offsets
  import ast
  from sofima_b200.mesh import IntegrationConfig
  g = np.load(COARSE)
  cfg = IntegrationConfig(**ast.literal_eval(str(g['cm2_short_cfg'])))
  np.testing.assert_array_equal(
      so.optimize_coarse_mesh(g['cm2_cx'], g['cm2_cy'], cfg), g['cm2_short'])
  np.testing.assert_array_equal(so.optimize_coarse_mesh(g['cm2_cx'], g['cm2_cy']), g['cm2_opt'])
  np.testing.assert_array_equal(
      so.optimize_coarse_mesh(g['cm3_cx'], g['cm3_cy'], mesh_fn=so.elastic_tile_mesh_3d),
      g['cm3_opt'])
  # the solution realises the requested offsets where the grid is consistent with them
  opt = g['cm2_opt']
  dx = opt[0, 0, :, 1:] - opt[0, 0, :, :-1]
  assert np.abs(dx - g['cm2_cx'][0, 0, :, :-1]).max() < 8.0


def test_interpolate_missing_offsets_golden():
  g = np.load(COARSE)
  for key, axis, kw in (('im_x', -1, {}), ('im_y', -2, {}), ('im_y_r2', -2, {'max_r': 2})):
    got = so.interpolate_missing_offsets(g['im_in'].copy(), axis, **kw)
    np.testing.assert_array_equal(got, g[key])
  with pytest.raises(ValueError):
    so.interpolate_missing_offsets(np.zeros((2, 3, 4)), -1)


def test_product_interpolate_missing_offsets_golden():
  # host NumPy in the product too (as in the reference); the relaxation itself runs on the
  # device only (tests/test_tile_mesh_gpu.py) and fails loudly without one
  import torch
  from sofima_b200 import _native, stitch_rigid
  g = np.load(COARSE)
  for key, axis, kw in (('im_x', -1, {}), ('im_y', -2, {}), ('im_y_r2', -2, {'max_r': 2})):
    got = stitch_rigid.interpolate_missing_offsets(g['im_in'].copy(), axis, **kw)
    np.testing.assert_array_equal(got, g[key])
  with pytest.raises(ValueError):
    stitch_rigid.interpolate_missing_offsets(np.zeros((2, 3, 4)), -1)
  with pytest.raises(NotImplementedError):  # only the built-in tile force fields
    stitch_rigid.optimize_coarse_mesh(g['cm2_cx'], g['cm2_cy'], mesh_fn=lambda x, *a: x)
  if not torch.cuda.is_available():
    with pytest.raises(_native.NativeError):
      stitch_rigid.optimize_coarse_mesh(g['cm2_cx'], g['cm2_cy'])
