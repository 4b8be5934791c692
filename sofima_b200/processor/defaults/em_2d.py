"""EM-2D default configurations (reference processor/defaults/em_2d.py:28-238): the ten
builders, registered like upstream so that `default_config(Config, EM_2D)` finds them."""

from __future__ import annotations

from typing import Any

from ... import mesh as mesh_lib
from ...compat import config as cfg_lib
from .. import flow
from .. import maps
from .. import mesh
from .. import warp

update_dataclass = cfg_lib.update_dataclass
EM_2D = cfg_lib.DefaultConfigType.EM_2D


def estimate_flow_config(overrides: dict[str, Any] | None = None) -> flow.EstimateFlow.Config:
  """em_2d.py:28-45."""
  config = flow.EstimateFlow.Config(
      patch_size=160, stride=40, z_stride=1, fixed_current=False, mask_configs=None,
      mask_only_for_patch_selection=True, selection_mask_configs=None, batch_size=1024)
  return update_dataclass(config, overrides)


def reconcile_flows_config(
    overrides: dict[str, Any] | None = None) -> flow.ReconcileAndFilterFlows.Config:
  """em_2d.py:48-67."""
  config = flow.ReconcileAndFilterFlows.Config(
      flow_volinfos=None, mask_configs=None, min_peak_ratio=1.6, min_peak_sharpness=1.6,
      max_magnitude=40, max_deviation=10, max_gradient=40, min_patch_size=400,
      multi_section=False, base_delta_z=1)
  return update_dataclass(config, overrides)


def estimate_missing_flow_config(
    overrides: dict[str, Any] | None = None) -> flow.EstimateMissingFlow.Config:
  """em_2d.py:70-95."""
  config = flow.EstimateMissingFlow.Config(
      patch_size=160, stride=40, delta_z=1, max_delta_z=4, max_attempts=2, mask_configs=None,
      mask_only_for_patch_selection=True, selection_mask_configs=None, min_peak_ratio=1.6,
      min_peak_sharpness=1.6, max_magnitude=40, batch_size=1024, image_volinfo=None,
      image_cache_bytes=int(1e9), mask_cache_bytes=int(1e9), search_radius=0)
  return update_dataclass(config, overrides)


def reconcile_missing_flows_config(
    overrides: dict[str, Any] | None = None) -> flow.ReconcileAndFilterFlows.Config:
  """em_2d.py:98-116."""
  config = update_dataclass(reconcile_flows_config(), {
      'multi_section': True, 'max_magnitude': 0, 'max_deviation': 10, 'max_gradient': 10,
      'min_patch_size': 400, 'base_delta_z': 1})
  return update_dataclass(config, overrides)


def integration_config(overrides: dict[str, Any] | None = None) -> mesh_lib.IntegrationConfig:
  """The IntegrationConfig inside relax_mesh_config (em_2d.py:143-156)."""
  config = mesh_lib.IntegrationConfig(
      dt=0.001, gamma=0.0, k0=0.01, k=0.1, stride=(40, 40), num_iters=1000,
      max_iters=100000, stop_v_max=0.005, dt_max=1000, start_cap=0.01, final_cap=10,
      prefer_orig_order=True)
  return update_dataclass(config, overrides)


def relax_mesh_config(overrides: dict[str, Any] | None = None) -> mesh.RelaxMesh.Config:
  """em_2d.py:136-177."""
  config = mesh.RelaxMesh.Config(
      output_dir='NONE', integration_config=integration_config(), mesh=None, flows=[],
      sections_to_skip=[], ranges_to_skip=[], mask=None, block_starts=[], block_ends=[],
      backward=False, mesh_min_frac=0.5, mesh_max_frac=2.0, coming_in=[],
      options=mesh.MeshOptions(irregular_mask_radius=5))
  return update_dataclass(config, overrides)


def within_block_config(overrides: dict[str, Any] | None = None) -> mesh.RelaxMesh.Config:
  """em_2d.py:187-194."""
  return update_dataclass(relax_mesh_config(), overrides)


def last_section_config(overrides: dict[str, Any] | None = None) -> mesh.RelaxMesh.Config:
  """em_2d.py:197-204."""
  return update_dataclass(relax_mesh_config(), overrides)


def cross_block_config(overrides: dict[str, Any] | None = None) -> mesh.RelaxMesh.Config:
  """em_2d.py:207-223."""
  config = relax_mesh_config({
      'integration_config': {'k0': 0.001, 'stride': (320, 320), 'stop_v_max': 0.001},
      'options': {'init_state': mesh.MeshInitState.PREV_MEDIAN},
  })
  return update_dataclass(config, overrides)


def default_em_2d_reconcile_config(
    overrides: dict[str, Any] | None = None) -> maps.ReconcileCrossBlockMaps.Config:
  """em_2d.py:226-241."""
  config = maps.ReconcileCrossBlockMaps.Config(
      cross_block='NONE', cross_block_inv='NONE', last_inv='NONE', main_inv='NONE', z_map={},
      stride=40, xy_overlap=128, backward=False)
  return update_dataclass(config, overrides)


def warp_config(overrides: dict[str, Any] | None = None) -> warp.WarpByMap.Config:
  """em_2d.py:244-262."""
  config = warp.WarpByMap.Config(
      stride=40, map_volinfo='UNSET', data_volinfo='UNSET', map_decorator_specs=None,
      data_decorator_specs=None, map_scale=1.0, interpolation='nearest', downsample=1,
      offset=0.0, mask_configs=None, source_cache_bytes=int(1e9))
  return update_dataclass(config, overrides)


cfg_lib.register_default_config(EM_2D, flow.EstimateFlow.Config, estimate_flow_config)
cfg_lib.register_default_config(EM_2D, flow.ReconcileAndFilterFlows.Config,
                                reconcile_flows_config)
cfg_lib.register_default_config(EM_2D, flow.EstimateMissingFlow.Config,
                                estimate_missing_flow_config)
cfg_lib.register_default_config(EM_2D, mesh.RelaxMesh.Config, relax_mesh_config)
cfg_lib.register_default_config(EM_2D, warp.WarpByMap.Config, warp_config)
