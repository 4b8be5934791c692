"""Plugin boundary: processor.flow.EstimateFlow, processor.mesh.RelaxMesh, decorator
chunk functions (reference processor/flow.py:43-275, processor/mesh.py:428-557,
decorators/flow.py:96-105, decorators/flow_test.py:55-88)."""

import dataclasses

import numpy as np
import pytest

from sofima_b200 import compat
from sofima_b200.processor import flow as pflow
from sofima_b200.processor import mesh as pmesh
from sofima_b200.processor.defaults import em_2d


def _cfg(**kw):
  base = dict(patch_size=80, stride=40, z_stride=1, fixed_current=False,
              mask_configs=None, mask_only_for_patch_selection=True,
              selection_mask_configs=None, batch_size=16)
  base.update(kw)
  return pflow.EstimateFlow.Config(**base)


# ---- CPU: geometry contract and host logic -------------------------------------------


def test_estimate_flow_geometry():
  p = pflow.EstimateFlow(_cfg())
  assert p.context() == ((40, 40, 1), (40, 40, 0))
  assert pflow.EstimateFlow(_cfg(z_stride=-2)).context() == ((40, 40, 0), (40, 40, 2))
  assert pflow.EstimateFlow(_cfg(fixed_current=True)).context() == ((40, 40, 0), (40, 40, 1))
  assert pflow.EstimateFlow(_cfg(z_stride=-1, fixed_current=True)).context() == (
      (40, 40, 1), (40, 40, 0))
  assert tuple(p.subvolume_size()) == (640, 640, 16)
  assert p.overlap() == (40, 40, 1)          # context sum minus one stride in x, y
  assert p.num_channels(1) == 4 and p.output_type(np.uint8) == np.float32
  np.testing.assert_array_equal(p.pixelsize(np.array([8, 8, 30])), [320, 320, 30])
  box = compat.BoundingBox(start=(0, 0, 0), size=(640, 640, 16))
  out = p.expected_output_box(box)
  np.testing.assert_array_equal(out.start, [1, 1, 1])
  np.testing.assert_array_equal(out.size, [15, 15, 15])
  with pytest.raises(AssertionError):
    pflow.EstimateFlow(_cfg(patch_size=90))


def test_section_pairs():
  assert pflow.EstimateFlow(_cfg())._section_pairs(4) == [(0, 1), (1, 2), (2, 3)]
  assert pflow.EstimateFlow(_cfg(z_stride=-1))._section_pairs(4) == [(1, 0), (2, 1), (3, 2)]
  assert pflow.EstimateFlow(_cfg(z_stride=2))._section_pairs(5) == [(0, 2), (1, 3), (2, 4)]
  assert pflow.EstimateFlow(_cfg(fixed_current=True))._section_pairs(4) == [
      (0, 3), (1, 3), (2, 3)]
  assert pflow.EstimateFlow(_cfg(z_stride=-1, fixed_current=True))._section_pairs(4) == [
      (1, 0), (2, 0), (3, 0)]


def test_em_2d_defaults():  # processor/defaults/em_2d.py:32-41,139-152
  f = em_2d.estimate_flow_config()
  assert (f.patch_size, f.stride, f.batch_size, f.mask_only_for_patch_selection) == (
      160, 40, 1024, True)
  ic = em_2d.relax_mesh_config().integration_config
  assert (ic.dt, ic.k0, ic.k, ic.stride, ic.num_iters, ic.max_iters) == (
      0.001, 0.01, 0.1, (40, 40), 1000, 100000)
  assert (ic.stop_v_max, ic.dt_max, ic.start_cap, ic.final_cap, ic.prefer_orig_order) == (
      0.005, 1000, 0.01, 10, True)
  assert em_2d.estimate_flow_config({'z_stride': -1}).z_stride == -1


def test_mask_irregular():
  m = np.zeros((2, 6, 7))
  assert not pmesh.mask_irregular(m.copy(), (40, 40), 0.5).any()
  fold = m.copy()
  fold[0, 2, 3] = 30.0      # node 3 moves 30 px right: gap 3->4 = 10 < 20, gap 2->3 = 70 > 60
  bad = pmesh.mask_irregular(fold, (40, 40), 0.5, dilation_iters=0)
  assert bad[2, 3] and bad[2, 2] and bad.sum() == 2 and np.isnan(fold[:, 2, 3]).all()
  stretch = m.copy()
  stretch[1, 4, 1] = 45.0   # gap above grows to 85 > 1.5 * 40; gap below shrinks to -5
  bad = pmesh.mask_irregular(stretch, (40, 40), 0.5, dilation_iters=0)
  assert bad[3, 1] and bad[4, 1] and bad.sum() == 2
  bad = pmesh.mask_irregular(fold.copy() * 0 + m, (40, 40), 0.5, dilation_iters=1)
  assert not bad.any()


def test_bounding_box_compat():
  b = compat.BoundingBox(start=(10, 20, 3), size=(100, 60, 2))
  np.testing.assert_array_equal(b.end, [110, 80, 5])
  s = b.scale([0.25, 0.25, 1])
  np.testing.assert_array_equal(s.start, [2, 5, 3])
  np.testing.assert_array_equal(s.end, [28, 20, 5])
  assert b.to_slice3d() == (slice(3, 5), slice(20, 80), slice(10, 110))


# ---- GPU: the plugins drive the CUDA backend -------------------------------------------

gpu = pytest.mark.gpu


def _need_gpu():
  import torch
  if not torch.cuda.is_available():
    pytest.skip('needs a CUDA device')


def _volume(nz=3, n=360, seed=0):
  import scipy.ndimage as ndi
  rng = np.random.default_rng(seed)
  base = ndi.gaussian_filter(rng.standard_normal((n + 40, n + 40)), 1.5)
  base = ((base - base.min()) / (base.max() - base.min()) * 255).astype(np.uint8)
  secs = [base[10 + 2 * z:10 + 2 * z + n, 12 - 3 * z:12 - 3 * z + n] for z in range(nz)]
  return np.stack(secs)[np.newaxis]  # [1, z, y, x]


@gpu
def test_estimate_flow_process_matches_oracle():
  _need_gpu()
  from oracle import flow_oracle as fo
  vol = _volume()
  box = compat.BoundingBox(start=(0, 0, 10), size=(360, 360, 3))
  for kw in (dict(), dict(z_stride=-1), dict(fixed_current=True)):
    proc = pflow.EstimateFlow(_cfg(**kw))
    out = proc.process(compat.Subvolume(vol, box))
    pairs = proc._section_pairs(3)
    assert out.data.shape == (4, len(pairs), 8, 8) and out.data.dtype == np.float32
    np.testing.assert_array_equal(out.bbox.size, [8, 8, len(pairs)])
    calc = fo.MaskedXCorrWithStatsCalculator()
    for i, (zp, zc) in enumerate(pairs):
      want = calc.flow_field(vol[0, zp], vol[0, zc], 80, 40, batch_size=16)
      np.testing.assert_array_equal(out.data[:2, i], want[:2])
      np.testing.assert_allclose(out.data[2:, i], want[2:], rtol=2e-3, atol=1e-6)
  # section z+1 is section z shifted by (dy, dx) = (2, -3): flow = (-3, 2)
  out = pflow.EstimateFlow(_cfg()).process(compat.Subvolume(vol, box))
  assert (out.data[0] == -3).all() and (out.data[1] == 2).all()


@gpu
def test_relax_mesh_plugin():
  _need_gpu()
  from oracle import mesh_oracle as mo
  import scipy.ndimage as ndi
  rng = np.random.default_rng(1)
  prev = (ndi.gaussian_filter(rng.standard_normal((2, 1, 40, 48)), (0, 0, 3, 3)) * 30)
  ic = dataclasses.replace(em_2d.integration_config(), num_iters=100, max_iters=2000,
                           k0=0.05)

  class Proc(pmesh.RelaxMesh):
    def get_prev_state(self, stride, box):
      return prev.copy()

  proc = Proc(pmesh.RelaxMesh.Config(integration_config=ic))
  x, e_kin, steps, status = proc.relax_mesh(np.zeros_like(prev), prev.copy(), ic, None)
  assert status == pmesh.SolutionStatus.REGULAR and steps % 100 == 0 and len(e_kin) == steps // 100
  want, _, t = mo.relax_mesh(np.zeros_like(prev), prev, ic)
  assert t == steps
  np.testing.assert_allclose(x, want, rtol=0, atol=1e-5)
  # masked nodes stay NaN and the plugin's process() returns the same solution
  mask = np.zeros((1, 40, 48), bool)
  mask[0, :3, :3] = True
  xm, *_ = proc.relax_mesh(np.zeros_like(prev), prev.copy(), ic, mask)
  assert np.isnan(xm[:, 0, :3, :3]).all() and np.isfinite(xm[:, 0, 5:, 5:]).all()
  sub = proc.process(compat.Subvolume(np.zeros((1, 1, 40, 48)),
                                      compat.BoundingBox(start=(0, 0, 7), size=(48, 40, 1))))
  np.testing.assert_array_equal(sub.data, x)
  # first section of a block is not optimised
  proc0 = Proc(pmesh.RelaxMesh.Config(integration_config=ic, block_starts=(7,)))
  sub0 = proc0.process(compat.Subvolume(np.zeros((1, 1, 40, 48)),
                                        compat.BoundingBox(start=(0, 0, 7), size=(48, 40, 1))))
  assert not sub0.data.any()


@gpu
def test_relax_mesh_plugin_fold_retry():
  _need_gpu()
  # A violent local pull folds the mesh: the plugin must detect it and go through
  # the regularisation path (status != REGULAR) without raising.
  prev = np.zeros((2, 1, 24, 24))
  prev[0, 0, 10:14, 10:14] = 150.0
  ic = dataclasses.replace(em_2d.integration_config(), num_iters=200, max_iters=4000,
                           k0=0.5, start_cap=10.0)
  proc = pmesh.RelaxMesh(pmesh.RelaxMesh.Config(integration_config=ic))
  x, e_kin, steps, status = proc.relax_mesh(np.zeros_like(prev), prev, ic, None)
  assert status in (pmesh.SolutionStatus.PREP_FAILED, pmesh.SolutionStatus.REGULARIZED)
  assert x.shape == prev.shape and steps > 0


@gpu
def test_decorator_chunk_functions():  # decorators/flow_test.py:55-88
  _need_gpu()
  from sofima_b200 import mesh
  from sofima_b200.decorators import flow as dflow
  rng = np.random.default_rng(0)
  args = dict(dt=0.001, gamma=0.0, k0=0.01, k=0.1, stride=(1, 1, 1), num_iters=50,
              max_iters=100, stop_v_max=0.001)
  flow3 = rng.standard_normal((3, 1, 6, 7, 8)).astype(np.float32)
  got = dflow.mesh_relax_flow(flow3, **args)
  cfg = mesh.IntegrationConfig(**args)
  want = mesh.relax_mesh(np.zeros_like(flow3.squeeze()), flow3.squeeze(), cfg,
                         mesh_force=mesh.elastic_mesh_3d)[0]
  np.testing.assert_array_equal(got, np.asarray(want).reshape(flow3.shape))
  flow2 = rng.standard_normal((2, 1, 12, 10)).astype(np.float32)
  args2 = dict(args, stride=(1, 1))
  got2 = dflow.mesh_relax_flow(flow2, **args2)
  assert got2.shape == flow2.shape and np.isfinite(got2).all()
  vol = _volume(2, 200)[0].astype(np.float32)
  f = dflow.optim_flow(vol[0], vol[1], (80, 80), (40, 40), batch_size=8)
  assert f.shape == (4, 4, 4) and (f[0] == -3).all() and (f[1] == 2).all()
  # the decorator classes over the in-memory stand-in for TensorStore
  # (decorators/flow_test.py:55-88: MeshRelaxFlowFilter == the direct call; :90-118 OptimFlow)
  from sofima_b200.compat import volume
  data = rng.uniform(size=(3, 3, 3, 3)).astype('float32')
  fargs = {'k0': 0.1, 'k': 0.1, 'dt': 0.001, 'gamma': 0.0, 'stride': (1, 1, 1),
           'num_iters': 1000, 'max_iters': 50_000, 'stop_v_max': 0.001, 'dt_max': 1000}
  store = volume.ArrayStore(data, ['fc', 'fz', 'fy', 'fx'])
  vc = dflow.MeshRelaxFlowFilter(min_chunksize=store.shape, **fargs).decorate(store)
  np.testing.assert_equal(vc[...].read().result(), dflow.mesh_relax_flow(data, **fargs))
  # OptimFlow: (x, y, z) stores, one flow field per section, padded to the image grid
  xyz = np.ascontiguousarray(np.stack([vol[0], vol[0]], axis=-1).transpose(1, 0, 2))
  fixed = np.ascontiguousarray(np.stack([vol[1], vol[1]], axis=-1).transpose(1, 0, 2))
  dec = dflow.OptimFlow(fixed_spec={'array': fixed, 'labels': ['x', 'y', 'z']},
                        image_dims=('x', 'y'), patch_size=(80, 80), step_size=(40, 40),
                        batch_size=8).decorate(volume.ArrayStore(xyz, ['x', 'y', 'z']))
  out = dec[...].read().result()
  assert out.shape == (4, 1, 5, 5, 2)
  np.testing.assert_array_equal(out[:, 0, 1:5, 1:5, 0], f)   # pad_left = 80 // 40 // 2 = 1
  np.testing.assert_array_equal(out[..., 0], out[..., 1])
  assert np.isnan(out[:, :, 0]).all() and np.isnan(out[:, :, :, 0]).all()


def test_masked_xcorr_rejects_unsupported_rank():
  # correlation over 4 axes is refused before the device is touched (no GPU needed here)
  from sofima_b200 import flow_field as ff
  a = np.zeros((6, 6, 6), np.float32)
  with pytest.raises(NotImplementedError):
    ff.masked_xcorr(a, a, dim=4)


# ---- decorator classes (ports of /root/reference/decorators/flow_test.py:25-118 on the
# in-memory stand-in for TensorStore, compat/volume.py) -------------------------------


def test_clean_flow_filter_decorator():
  from sofima_b200.compat import volume
  from sofima_b200.decorators import flow as decorators
  rng = np.random.default_rng(0)
  data = rng.uniform(size=(5, 3, 3, 3)).astype('float32')
  f = volume.ArrayStore(data, ['fc', 'fz', 'fy', 'fx'])
  filter_args = {'min_peak_sharpness': 1.6, 'min_peak_ratio': 1.4, 'max_magnitude': 20,
                 'max_deviation': 2}
  vc = decorators.CleanFlowFilter(min_chunksize=f.shape, **filter_args).decorate(f)
  assert vc.shape == (3, 3, 3, 3) and vc.domain.labels == ('fc', 'fz', 'fy', 'fx')
  res = vc[...].read().result()
  np.testing.assert_equal(res, decorators.clean_flow(data, **filter_args))


def test_reconcile_flow_filter_decorator():
  from sofima_b200.compat import volume
  from sofima_b200.decorators import flow as decorators
  rng = np.random.default_rng(1)
  data = rng.uniform(size=(3, 3, 3, 3)).astype('float32')   # decorators/flow_test.py:120-147
  f = volume.ArrayStore(data, ['fc', 'fz', 'fy', 'fx'])
  filter_args = {'max_gradient': 2.0, 'max_deviation': 2, 'min_patch_size': 20}
  vc = decorators.ReconcileFlowFilter(min_chunksize=f.shape, **filter_args).decorate(f)
  res = vc[...].read().result()
  np.testing.assert_equal(res, decorators.reconcile_flow(data, **filter_args))
  # a sub-domain read is cut out of the same whole-array chunk
  np.testing.assert_equal(vc[:, :, 1:3, 0:2].read().result(), res[:, :, 1:3, 0:2])


def test_optim_flow_decorator_geometry():
  from sofima_b200.compat import volume
  from sofima_b200.decorators import flow as decorators
  img = volume.ArrayStore(np.zeros((3, 3), np.float32), ['x', 'y'])
  dec = decorators.OptimFlow(fixed_spec=img, image_dims=('x', 'y')).decorate(img)
  assert dec.domain.labels == ('fc', 'fz', 'fy', 'fx')
  # 200 x 160 (x, y) images over 3 sections: padded flow grid + trailing non-image dim
  vol = volume.ArrayStore(np.zeros((200, 160, 3), np.float32), ['x', 'y', 'z'])
  dec = decorators.OptimFlow(fixed_spec={'array': np.zeros((200, 160, 3), np.float32),
                                         'labels': ['x', 'y', 'z']},
                             image_dims=('x', 'y'), patch_size=(80, 80), step_size=(40, 40),
                             batch_size=8).decorate(vol)
  assert dec.domain.labels == ('fc', 'fz', 'fy', 'fx', 'z')
  assert dec.shape == (4, 1, 4, 5, 3)   # ceil((160-80+1)/40) + 1 = 4, ceil((200-80+1)/40) + 1 = 5
  with pytest.raises(ValueError):
    decorators.OptimFlow(fixed_spec=img, image_dims=('x', 'y')).decorate(vol)
  with pytest.raises(ValueError):
    decorators.OptimFlow(fixed_spec=vol, image_dims=('x',)).decorate(vol)


def test_get_block_id():  # tests/client_utils_test.py:24-40
  from sofima_b200.processor import client_utils
  from sofima_b200.processor import mesh as pmesh
  assert pmesh.get_block_id is client_utils.get_block_id
  fwd_starts = [0, 50, 100, 150, 200]  # blocks 0..49, 50..99, ...
  assert [client_utils.get_block_id(z, fwd_starts, False) for z in (10, 0, 49, 50)] == [1, 1, 1, 2]
  bwd_starts = [50, 100, 150, 200]  # blocks 0..50, 51..100, ...
  assert [client_utils.get_block_id(z, bwd_starts, True)
          for z in (10, 0, 50, 51, 100)] == [0, 0, 0, 1, 1]
