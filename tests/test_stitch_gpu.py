"""GPU parity tests of the stitching target-mesh path:
sofima_b200.map_utils.compose_maps_fast, sofima_b200.stitch_elastic and
mesh.relax_mesh(prev_fn=...) through the C ABI, against the oracle and the golden
vectors generated from the reference source.

The kernels keep the reference's fp32 association order (compiled with -fmad=false),
so comparisons with the NumPy oracle are exact; 1e-5 abs (north_star) where fp32
reduction order enters (remove_drift).
"""

import ast
import os

import numpy as np
import pytest

from oracle import mesh_oracle as mo
from oracle import stitch_oracle as so

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'stitch_golden.npz')


@pytest.fixture(scope='module')
def g():
  return np.load(GOLDEN)


@pytest.fixture(scope='module')
def mods():
  import torch
  if not torch.cuda.is_available():
    pytest.skip('needs a CUDA device')
  from sofima_b200 import map_utils, mesh, stitch_elastic
  return map_utils, mesh, stitch_elastic


@pytest.mark.parametrize('mode', ['nearest', 'constant'])
def test_compose_maps_fast_golden(mods, g, mode):
  map_utils = mods[0]
  got = map_utils.compose_maps_fast(g['cmf_map1'], (3, 2, 1), (20, 25), g['cmf_map2'],
                                    (1, 4, 0), (16, 20), mode=mode)
  np.testing.assert_array_equal(got, g[f'cmf_{mode}'])
  got = map_utils.compose_maps_fast(g['cmf3_map1'], (1, 0, 2), (4, 10, 10), g['cmf3_map2'],
                                    (0, 1, 1), (4, 10, 10), mode=mode)
  np.testing.assert_array_equal(got, g[f'cmf3_{mode}'])


def test_compose_maps_fast_reference_kat(mods):
  # tests/map_utils_test.py:266-300
  map_utils = mods[0]
  coord_map = np.zeros([2, 1, 60, 60])
  flow = np.zeros([2, 1, 50, 50])
  flow[0, 0, :, 10:25] = -5
  flow[0, 0, :, 25:40] = 65
  flow[:, 0, :, 4] = np.nan
  s1, s2 = (64, 58, 42), (64, 50, 40)
  updated = np.array(map_utils.compose_maps_fast(flow, s1, 40, coord_map, s2, 40))
  np.testing.assert_array_equal(updated, flow)
  coord_map[0, :, :, 7:] = -10
  updated = np.array(map_utils.compose_maps_fast(flow, s1, 40, coord_map, s2, 40))
  flow[0, 0, :, 5:10] = -10
  flow[0, 0, :, 10:25] = -15
  flow[0, 0, :, 25:40] = 55
  flow[0, 0, :, 40:] = -10
  np.testing.assert_array_equal(updated, flow)


@pytest.mark.parametrize('dim,mode', [(2, 'nearest'), (2, 'constant'), (3, 'nearest'),
                                       (3, 'constant')])
def test_compose_maps_fast_random_vs_oracle(mods, dim, mode):
  map_utils = mods[0]
  rng = np.random.default_rng(dim * 10 + len(mode))
  if dim == 2:
    m1 = (rng.standard_normal((2, 3, 70, 90)) * 30).astype(np.float32)
    m2 = (rng.standard_normal((2, 3, 80, 75)) * 10).astype(np.float32)
    args = ((5, 7), (40.0, 32.5), (2, 9), (36.0, 40.0))
  else:
    m1 = (rng.standard_normal((3, 9, 20, 24)) * 12).astype(np.float32)
    m2 = (rng.standard_normal((3, 11, 22, 19)) * 5).astype(np.float32)
    args = ((1, 2, 0), (8.0, 20.0, 20.0), (0, 0, 3), (8.0, 16.0, 20.0))
  m1[:, 1, 5, 6] = np.nan
  m2[0, 2, 7, 7] = np.nan
  m1[0, 0, 0, 0] = np.inf
  m1[1, 0, 0, 1] = 3e12
  want = so.compose_maps_fast(m1, args[0], args[1], m2, args[2], args[3], mode=mode)
  got = map_utils.compose_maps_fast(m1, args[0], args[1], m2, args[2], args[3], mode=mode)
  np.testing.assert_array_equal(got, want)
  import torch
  got_t = map_utils.compose_maps_fast(torch.from_numpy(m1).cuda(), args[0], args[1],
                                      torch.from_numpy(m2).cuda(), args[2], args[3], mode=mode)
  assert got_t.is_cuda
  np.testing.assert_array_equal(got_t.cpu().numpy(), want)


def test_target_mesh_golden(mods, g):
  stitch_elastic = mods[2]
  stride = tuple(int(v) for v in g['st2_stride'])
  fn = stitch_elastic.target_mesh_fn(g['st2_nbors'], g['st2_fx'], g['st2_fy'], stride)
  got = fn(g['st2_x'])
  np.testing.assert_array_equal(got, g['st2_target'])
  # one tile at a time, the reference's own entry point
  for i, nd in enumerate(g['st2_nbors']):
    one = stitch_elastic.compute_target_mesh(nd, g['st2_x'], g['st2_fx'], g['st2_fy'], stride)
    np.testing.assert_array_equal(one, g['st2_target'][:, i])


def test_target_mesh_golden_3d(mods, g):
  stitch_elastic = mods[2]
  stride = tuple(int(v) for v in g['st3_stride'])
  fn = stitch_elastic.target_mesh_fn(g['st3_nbors'], g['st3_fx'], g['st3_fy'], stride)
  np.testing.assert_array_equal(fn(g['st3_x']), g['st3_target'])
  one = stitch_elastic.compute_target_mesh(g['st3_nbors'][2], g['st3_x'], g['st3_fx'],
                                           g['st3_fy'], stride)
  np.testing.assert_array_equal(one, g['st3_target'][:, 2])


@pytest.mark.parametrize('tag,atol', [('st3_relax', 1e-5), ('st3_relax_nodrift', 1e-6)])
def test_relax_mesh_3d_with_prev_fn_golden(mods, g, tag, atol):
  """notebooks/liconn_inplane_stitching.ipynb:763-783 against the reference's own
  trajectory (remove_drift there is one mean per x column, mesh.py:496-497)."""
  _, mesh, stitch_elastic = mods
  stride = tuple(int(v) for v in g['st3_stride'])
  prev_fn = stitch_elastic.target_mesh_fn(g['st3_nbors'], g['st3_fx'], g['st3_fy'], stride)
  cfg = mesh.IntegrationConfig(**ast.literal_eval(str(g[f'{tag}_cfg'])))
  x, e_kin, t = mesh.relax_mesh(g['st3_x'], None, cfg, prev_fn=prev_fn,
                                mesh_force=mesh.elastic_mesh_3d)
  assert t == 24
  np.testing.assert_allclose(x, g[f'{tag}_x'], rtol=0, atol=atol)
  np.testing.assert_allclose(e_kin, g[f'{tag}_ekin'], rtol=1e-5)
  with pytest.raises(ValueError):  # 2-d force with a 3-d target
    mesh.relax_mesh(g['st3_x'][:2, :, 0], None, cfg, prev_fn=prev_fn)


def _big_case(seed=3, nt_x=4, nt_y=3, mesh_shape=(51, 48), stride=(40.0, 40.0)):
  """A 4 x 3 tile grid with jittered overlaps, built through aggregate_arrays."""
  from sofima_b200 import stitch_elastic
  import scipy.ndimage as ndi
  rng = np.random.default_rng(seed)
  my, mx = mesh_shape
  coords = [(tx, ty) for ty in range(nt_y) for tx in range(nt_x)]
  cx = np.full((2, nt_y, nt_x), np.nan)
  cy = np.full((2, nt_y, nt_x), np.nan)
  fine_x, fine_y, off_x, off_y = {}, {}, {}, {}
  def smooth(shape, amp):
    return (ndi.gaussian_filter(rng.standard_normal(shape), (0, 2, 2)) * amp).astype(np.float32)
  for tx, ty in coords:
    if tx + 1 < nt_x:
      oy, ox = my - int(rng.integers(0, 3)), int(rng.integers(3, 6))
      f = np.full((4, oy, ox), 1.0, np.float32)
      f[:2] = smooth((2, oy, ox), 5.0)
      f[:, rng.integers(oy), rng.integers(ox)] = np.nan
      fine_x[tx, ty] = f
      cx[:, ty, tx] = (mx * stride[1] - ox * stride[1] + rng.integers(-9, 9),
                       rng.integers(-50, 50))
      off_x[tx, ty] = (int(rng.integers(-4, 4)), int(rng.integers(-4, 4)))
    if ty + 1 < nt_y:
      oy, ox = int(rng.integers(3, 6)), mx - int(rng.integers(0, 3))
      f = np.full((4, oy, ox), 1.0, np.float32)
      f[:2] = smooth((2, oy, ox), 5.0)
      fine_y[tx, ty] = f
      cy[:, ty, tx] = (rng.integers(-50, 50),
                       my * stride[0] - oy * stride[0] + rng.integers(-9, 9))
      off_y[tx, ty] = (int(rng.integers(-4, 4)), int(rng.integers(-4, 4)))
  coarse = (rng.standard_normal((2, nt_y, nt_x)) * 3)
  fx, fy, x, nbors, _ = stitch_elastic.aggregate_arrays(
      (cx, fine_x, off_x), (cy, fine_y, off_y), coords, coarse, stride,
      (my * int(stride[0]), mx * int(stride[1])))
  x = x + smooth((2 * len(coords), my, mx), 3.0).reshape(2, len(coords), my, mx)
  return fx.astype(np.float32), fy.astype(np.float32), x.astype(np.float32), nbors, stride


def test_target_mesh_grid_vs_oracle(mods):
  stitch_elastic = mods[2]
  fx, fy, x, nbors, stride = _big_case()
  x[:, 5, 10, 11] = np.nan
  want = so.target_mesh_all(nbors, x, fx, fy, stride)
  got = stitch_elastic.target_mesh_fn(nbors, fx, fy, stride)(x)
  np.testing.assert_array_equal(got, want)
  assert 0.05 < np.isfinite(want).mean() < 0.6


@pytest.mark.parametrize('tag,atol', [('st2_relax', 1e-5), ('st2_relax_damped', 0.0)])
def test_relax_mesh_with_prev_fn_golden(mods, g, tag, atol):
  """notebooks/em_stitching.ipynb:545-603 against the reference's own trajectory."""
  _, mesh, stitch_elastic = mods
  stride = tuple(int(v) for v in g['st2_stride'])
  prev_fn = stitch_elastic.target_mesh_fn(g['st2_nbors'], g['st2_fx'], g['st2_fy'], stride)
  cfg = mesh.IntegrationConfig(**ast.literal_eval(str(g[f'{tag}_cfg'])))
  x, e_kin, t = mesh.relax_mesh(g['st2_relax_x0'], None, cfg, prev_fn=prev_fn)
  assert t == int(g['st2_relax_t'])
  np.testing.assert_allclose(x, g[f'{tag}_x'], rtol=0, atol=atol)
  np.testing.assert_allclose(e_kin, g[f'{tag}_ekin'], rtol=1e-5)


@pytest.mark.parametrize('fire,drift', [(True, False), (True, True), (False, False)])
def test_relax_mesh_with_prev_fn_grid_vs_oracle(mods, fire, drift):
  _, mesh, stitch_elastic = mods
  fx, fy, x, nbors, stride = _big_case(seed=11)
  cfg = mesh.IntegrationConfig(
      dt=0.001, gamma=0.0 if fire else 0.4, k0=0.01, k=0.1, stride=stride, num_iters=25,
      max_iters=75, stop_v_max=0.0, dt_max=100, prefer_orig_order=True,
      start_cap=0.1 if fire else 10.0, final_cap=10.0, remove_drift=drift, fire=fire)
  prev_fn = stitch_elastic.target_mesh_fn(nbors, fx, fy, stride)
  want, ek_w, t_w = mo.relax_mesh(
      x, None, cfg, prev_fn=lambda a: so.target_mesh_all(nbors, a, fx, fy, stride))
  got, ek, t = mesh.relax_mesh(x, None, cfg, prev_fn=prev_fn)
  assert t == t_w == 75
  np.testing.assert_allclose(got, want, rtol=0, atol=1e-5 if drift else 0.0)
  np.testing.assert_allclose(ek, ek_w, rtol=1e-5)
  assert np.abs(got - x).max() > 1e-4  # the tiles actually moved


def test_prev_fn_errors(mods):
  _, mesh, stitch_elastic = mods
  fx, fy, x, nbors, stride = _big_case(seed=5, nt_x=2, nt_y=2, mesh_shape=(12, 12))
  cfg = mesh.IntegrationConfig(dt=0.001, gamma=0.0, k0=0.01, k=0.1, stride=stride,
                               num_iters=2, max_iters=2, stop_v_max=0.0)
  prev_fn = stitch_elastic.target_mesh_fn(nbors, fx, fy, stride)
  with pytest.raises(ValueError):  # mesh.py:567-568
    mesh.relax_mesh(x, x, cfg, prev_fn=prev_fn)
  with pytest.raises(NotImplementedError):
    mesh.relax_mesh(x, None, cfg, prev_fn=lambda a: a)
  with pytest.raises(ValueError):
    mesh.relax_mesh(x[:, :2], None, cfg, prev_fn=prev_fn)


def test_compute_flow_map_golden(mods, g):
  """stitch_elastic.compute_flow_map on the CUDA flow path vs the reference's run."""
  from tests.test_oracle_stitch import _check_flow_maps
  _check_flow_maps(g, mods[2])


def test_compute_flow_map3d_golden(mods):
  """stitch_elastic.compute_flow_map3d (LICONN fine flow, 3-d patches) on the CUDA flow
  path vs the reference's own run (tests/golden/flow3d_golden.npz)."""
  from tests.test_oracle_stitch import _check_flow_maps3d
  _check_flow_maps3d(mods[2])


def test_config2_pipeline_vs_oracle(mods, g, monkeypatch):
  """BASELINE config 2 in small: fine flow on the overlap strips -> aggregate_arrays ->
  relaxation with the stitching prev_fn, CUDA vs the same pipeline on the oracle."""
  from oracle import flow_oracle
  from sofima_b200 import flow_field
  _, mesh, stitch_elastic = mods
  tex = g['fm2_tex']
  coords = [(int(a), int(b)) for a, b, _, _ in g['fm2_nominal']]
  tiles = {(int(a), int(b)): np.ascontiguousarray(tex[y0:y0 + 160, x0:x0 + 200])
           for a, b, y0, x0 in g['fm2_nominal']}
  stride = (8, 8)
  cx3, cy3 = g['fm2_cx'], g['fm2_cy']

  def pipeline(relax, target_fn):
    fx, ox = stitch_elastic.compute_flow_map(tiles, cx3, 0, (32, 32), stride, 64)
    fy, oy = stitch_elastic.compute_flow_map(tiles, cy3, 1, (32, 32), stride, 64)
    afx, afy, x, nbors, _ = stitch_elastic.aggregate_arrays(
        (cx3, fx, ox), (cy3, fy, oy), coords, np.zeros((2, 2, 3)), stride, (160, 200))
    cfg = mesh.IntegrationConfig(dt=0.001, gamma=0., k0=0.01, k=0.1, stride=stride,
                                 num_iters=20, max_iters=60, stop_v_max=0.0, dt_max=100,
                                 prefer_orig_order=True, start_cap=0.1, final_cap=10.)
    afx, afy = afx.astype(np.float32), afy.astype(np.float32)
    return relax(x, None, cfg, prev_fn=target_fn(nbors, afx, afy, stride)), nbors, afx

  (got, ek, t), nbors, afx = pipeline(mesh.relax_mesh, stitch_elastic.target_mesh_fn)
  monkeypatch.setattr(flow_field, 'JAXMaskedXCorrWithStatsCalculator',
                      flow_oracle.MaskedXCorrWithStatsCalculator)
  oracle_fn = lambda nb, a, b, st: (lambda x: so.target_mesh_all(nb, x, a, b, st))
  (want, ek_w, t_w), nbors_w, afx_w = pipeline(mo.relax_mesh, oracle_fn)
  np.testing.assert_array_equal(nbors, nbors_w)
  np.testing.assert_array_equal(afx, afx_w)      # integer flow vectors: identical
  assert t == t_w == 60
  np.testing.assert_array_equal(got, want)
  assert np.abs(got).max() > 1e-4


def test_compute_coarse_offsets_golden(mods, g):
  """stitch_rigid.compute_coarse_offsets (whole-strip masked correlations on the GPU)
  vs the reference's run."""
  from sofima_b200 import stitch_rigid
  from tests.test_oracle_stitch import _check_coarse_offsets
  _check_coarse_offsets(g, stitch_rigid)


@pytest.mark.parametrize('shape,shift', [((1500, 100), (7, -3)), ((120, 1600), (-4, 9)),
                                         ((4096, 300), (11, 5))])
def test_whole_strip_masked_correlation(mods, shape, shift):
  """One masked correlation of a whole overlap strip (stitch_rigid.py:39-67): transform
  lengths up to 8192 x 600 -- the long-column / long-row form of the flow kernels."""
  import scipy.ndimage as ndi
  from oracle import flow_oracle
  from sofima_b200 import flow_field
  rng = np.random.default_rng(shape[0])
  h, w = shape
  base = ndi.gaussian_filter(rng.standard_normal((h + 40, w + 40)), 2.0)
  base = ((base - base.min()) / (base.max() - base.min()) * 255).astype(np.uint8)
  a = np.ascontiguousarray(base[20:20 + h, 20:20 + w])
  b = np.ascontiguousarray(base[20 + shift[0]:20 + shift[0] + h, 20 + shift[1]:20 + shift[1] + w])
  am = rng.random(shape) < 0.05
  bm = rng.random(shape) < 0.05
  am[: h // 7, : w // 3] = True
  kw = dict(pre_mask=am, post_mask=bm, patch_size=shape, step=(1, 1), batch_size=1)
  got = flow_field.JAXMaskedXCorrWithStatsCalculator().flow_field(a, b, **kw)
  want = flow_oracle.MaskedXCorrWithStatsCalculator().flow_field(a, b, **kw)
  assert got.shape == want.shape == (4, 1, 1)
  np.testing.assert_array_equal(got[:2], want[:2])
  np.testing.assert_array_equal(got[:2, 0, 0], (shift[1], shift[0]))
  np.testing.assert_allclose(got[2:], want[2:], rtol=5e-3, atol=1e-6)
