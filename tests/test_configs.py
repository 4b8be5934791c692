"""EM-2D default configs and pipeline configs (ports of /root/reference/pipeline/
flow_config_test.py:54-90 and of the defaults in processor/defaults/em_2d.py)."""

import dataclasses

import pytest

from sofima_b200.compat import config as cfg_lib
from sofima_b200.pipeline import flow_config
from sofima_b200.pipeline import mesh_config
from sofima_b200.processor import flow
from sofima_b200.processor import mesh
from sofima_b200.processor import warp
from sofima_b200.processor.defaults import em_2d


def _expected_default_em_2d(overrides=None) -> flow_config.FlowPipeline:
  default_config = flow_config.FlowPipeline.from_dict({
      'estimate_flow': {
          'config': em_2d.estimate_flow_config(),
          'processing': {'overlap': [160, 160, 1], 'subvolume_size': [3200, 3200, 128]},
          'schedule_batch_size': 16384,
          'corner_whitelist': set(),
          'ignore_existing': False,
          'delete_existing': False,
      },
      'reconcile_flows': em_2d.reconcile_flows_config(),
      'estimate_missing_flow': em_2d.estimate_missing_flow_config(),
      'reconcile_missing_flows': em_2d.reconcile_flows_config({
          'multi_section': True, 'max_magnitude': 0, 'max_deviation': 10, 'max_gradient': 10,
          'min_patch_size': 400, 'base_delta_z': 1}),
  })
  return cfg_lib.update_dataclass(default_config, overrides)


def test_flow_pipeline_default_em_2d():   # pipeline/flow_config_test.py:54-90
  config = cfg_lib.default_config(flow_config.FlowPipeline, cfg_lib.DefaultConfigType.EM_2D)
  assert config == _expected_default_em_2d()
  assert isinstance(config.estimate_flow.processing, cfg_lib.ProcessingConfig)
  config = cfg_lib.default_config(
      flow_config.FlowPipeline, cfg_lib.DefaultConfigType.EM_2D,
      {'estimate_flow': {'config': {'z_stride': 12321},
                         'processing': {'subvolume_size': [1000, 1000, 1]}}})
  assert config == _expected_default_em_2d(
      {'estimate_flow': {'config': {'z_stride': 12321},
                         'processing': {'overlap': [160, 160, 12321],
                                        'subvolume_size': [1000, 1000, 1]}}})


def test_em_2d_builders_match_the_reference_values():
  ef = em_2d.estimate_flow_config()
  assert (ef.patch_size, ef.stride, ef.batch_size, ef.mask_only_for_patch_selection) == (
      160, 40, 1024, True)
  rf = em_2d.reconcile_flows_config()
  assert (rf.min_peak_ratio, rf.min_peak_sharpness, rf.max_magnitude, rf.max_deviation,
          rf.max_gradient, rf.min_patch_size, rf.multi_section) == (1.6, 1.6, 40, 10, 40, 400, False)
  mf = em_2d.estimate_missing_flow_config({'max_delta_z': 2})
  assert (mf.patch_size, mf.max_delta_z, mf.max_attempts, mf.search_radius) == (160, 2, 2, 0)
  rm = em_2d.reconcile_missing_flows_config()
  assert rm.multi_section and rm.max_magnitude == 0 and rm.max_gradient == 10
  rc = em_2d.relax_mesh_config()
  ic = rc.integration_config
  assert (ic.k0, ic.k, ic.stride, ic.num_iters, ic.max_iters, ic.stop_v_max, ic.start_cap,
          ic.final_cap, ic.prefer_orig_order) == (0.01, 0.1, (40, 40), 1000, 100000, 0.005,
                                                  0.01, 10, True)
  assert rc.options.irregular_mask_radius == 5 and rc.mesh_min_frac == 0.5
  assert em_2d.within_block_config() == rc == em_2d.last_section_config()
  cb = em_2d.cross_block_config()
  assert cb.integration_config.k0 == 0.001 and cb.integration_config.stride == (320, 320)
  assert cb.integration_config.stop_v_max == 0.001 and cb.integration_config.k == 0.1
  assert cb.options.init_state == mesh.MeshInitState.PREV_MEDIAN
  assert cb.options.irregular_mask_radius == 5          # nested override keeps the rest
  wc = em_2d.warp_config({'map_volinfo': 'm', 'data_volinfo': 'd'})
  assert (wc.stride, wc.interpolation, wc.downsample, wc.map_volinfo) == (40, 'nearest', 1, 'm')
  rcb = em_2d.default_em_2d_reconcile_config()
  assert (rcb.stride, rcb.xy_overlap, rcb.backward, rcb.z_map) == (40, 128, False, {})
  with pytest.raises(KeyError):
    em_2d.estimate_flow_config({'no_such_field': 1})


def test_registered_defaults_and_mesh_pipeline():
  t = cfg_lib.DefaultConfigType.EM_2D
  assert cfg_lib.default_config(flow.EstimateFlow.Config, t) == em_2d.estimate_flow_config()
  assert cfg_lib.default_config(flow.EstimateMissingFlow.Config, t, {'stride': 20}).stride == 20
  assert cfg_lib.default_config(mesh.RelaxMesh.Config, t) == em_2d.relax_mesh_config()
  assert cfg_lib.default_config(warp.WarpByMap.Config, t) == em_2d.warp_config()
  mp = cfg_lib.default_config(mesh_config.MeshRelaxationConfig, t)
  assert mp.within_block_config == em_2d.within_block_config()
  assert mp.cross_block_config == em_2d.cross_block_config()
  assert mp.reconcile_cross_block_config == em_2d.default_em_2d_reconcile_config()
  mp2 = mesh_config.default_em_2d({'cross_block_config': {'integration_config': {'k0': 0.5}}})
  assert mp2.cross_block_config.integration_config.k0 == 0.5
  assert mp2.cross_block_config.integration_config.stride == (320, 320)
  # JSON round trip of a pipeline config
  d = mp.to_dict()
  assert d['cross_block_config']['integration_config']['stride'] == [320, 320]
  assert dataclasses.is_dataclass(flow_config.default_em_2d().estimate_flow)


def test_warp_pipeline_default_em_2d():   # pipeline/warp_config.py:35-50
  from sofima_b200.pipeline import warp_config
  t = cfg_lib.DefaultConfigType.EM_2D
  wp = cfg_lib.default_config(warp_config.WarpPipelineConfig, t)
  assert wp == warp_config.default_em_2d() and wp.warp == em_2d.warp_config()
  wp2 = warp_config.default_em_2d({'warp': {'interpolation': 'lanczos', 'data_volinfo': 'd'}})
  assert wp2.warp.interpolation == 'lanczos' and wp2.warp.data_volinfo == 'd'
  assert wp2.warp.stride == wp.warp.stride
  assert wp.to_dict()['warp']['stride'] == 40
