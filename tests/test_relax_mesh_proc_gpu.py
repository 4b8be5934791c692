"""GPU test of the `RelaxMesh` plugin's section-to-section path: flow volume ->
compose with the solved reference section (map_utils.compose_maps_fast on the GPU) ->
`prev` -> relaxation, against the oracle (processor/mesh.py:248-383, :513-557)."""

import types

import numpy as np
import pytest
import scipy.ndimage as ndi

from oracle import mesh_oracle as mo
from oracle import stitch_oracle as so

pytestmark = pytest.mark.gpu


class _Vol:
  def __init__(self, data, channels):
    self.data = data
    self.meta = types.SimpleNamespace(num_channels=channels)

  def __getitem__(self, sl):
    return self.data[sl].copy()


def test_relax_mesh_sections():
  import torch
  if not torch.cuda.is_available():
    pytest.skip('needs a CUDA device')
  from sofima_b200 import compat, mesh
  from sofima_b200.processor import mesh as pm

  rng = np.random.default_rng(4)
  ny, nx, nz, stride = 40, 44, 3, (40.0, 40.0)
  flow = (ndi.gaussian_filter(rng.standard_normal((2, nz, ny, nx)), (0, 0, 3, 3)) * 25)
  flow = flow.astype(np.float32)
  flow[:, 1, 5, 6] = np.nan
  flow3 = np.concatenate([flow, np.ones((1, nz, ny, nx), np.float32)])  # multi-z: delta_z = 1
  volumes = {'flow1': _Vol(flow, 2), 'flow3': _Vol(flow3, 3)}
  solved = {}

  class Proc(pm.RelaxMesh):
    def _open_volume(self, volume):
      return volumes[volume]

    def _load_stitched_tile(self, output_dir, box):
      return solved.get(int(box.start[2]))

  ic = mesh.IntegrationConfig(dt=0.001, gamma=0.0, k0=0.01, k=0.1, stride=stride,
                              num_iters=50, max_iters=200, stop_v_max=0.005, dt_max=1000,
                              prefer_orig_order=True, start_cap=0.01, final_cap=10.0)
  cfg = pm.RelaxMesh.Config(integration_config=ic, flows=[pm.FlowVolume(1, 'flow1')],
                            block_starts=[0], block_ends=[2])
  proc = Proc(cfg)
  box = lambda z: compat.BoundingBox(start=(0, 0, z), size=(nx, ny, 1))

  x0, _, steps, status = proc.run_relaxation(box(0))
  assert status == pm.SolutionStatus.UNDEFINED and steps == 0 and not x0.any()
  solved[0] = x0

  want_prev = None
  for z in (1, 2):
    prev = proc.get_prev_state(stride, box(z))
    want_prev = so.compose_maps_fast(flow[:, z:z + 1], (z, 0, 0), stride, solved[z - 1],
                                     (z - 1, 0, 0), stride).astype(np.float64)
    pm.mask_irregular(want_prev[:, 0], stride, cfg.mesh_min_frac, cfg.mesh_max_frac,
                      dilation_iters=1)
    np.testing.assert_array_equal(prev, want_prev)
    x, e_kin, steps, status = proc.run_relaxation(box(z))
    assert status == pm.SolutionStatus.REGULAR and steps > 0
    want, _, t_w = mo.relax_mesh(np.zeros_like(prev), prev, ic)
    assert steps == t_w
    np.testing.assert_array_equal(x, want)
    solved[z] = x
    out = proc.process(compat.Subvolume(np.zeros((2, 1, ny, nx)), box(z)))
    np.testing.assert_array_equal(out.data, x)

  # 3-channel flow: the third channel selects the reference section per node
  cfg3 = pm.RelaxMesh.Config(integration_config=ic, flows=[pm.FlowVolume(1, 'flow3')],
                             block_starts=[0], block_ends=[2])
  prev3 = Proc(cfg3).get_prev_state(stride, box(2))
  np.testing.assert_array_equal(prev3, want_prev)
  # skipped reference sections contribute nothing
  cfg_skip = pm.RelaxMesh.Config(integration_config=ic, flows=[pm.FlowVolume(1, 'flow1')],
                                 block_starts=[0], sections_to_skip=[1])
  assert Proc(cfg_skip).get_prev_state(stride, box(2)) is None
