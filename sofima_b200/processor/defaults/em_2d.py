"""EM-2D default configurations (reference processor/defaults/em_2d.py:28-45,136-156)."""

from __future__ import annotations

import dataclasses
from typing import Any

from ... import mesh as mesh_lib
from .. import flow
from .. import mesh


def estimate_flow_config(overrides: dict[str, Any] | None = None) -> flow.EstimateFlow.Config:
  cfg = flow.EstimateFlow.Config(
      patch_size=160, stride=40, z_stride=1, fixed_current=False, mask_configs=None,
      mask_only_for_patch_selection=True, selection_mask_configs=None, batch_size=1024)
  return dataclasses.replace(cfg, **overrides) if overrides else cfg


def integration_config(overrides: dict[str, Any] | None = None) -> mesh_lib.IntegrationConfig:
  cfg = mesh_lib.IntegrationConfig(
      dt=0.001, gamma=0.0, k0=0.01, k=0.1, stride=(40, 40), num_iters=1000,
      max_iters=100000, stop_v_max=0.005, dt_max=1000, start_cap=0.01, final_cap=10,
      prefer_orig_order=True)
  return dataclasses.replace(cfg, **overrides) if overrides else cfg


def relax_mesh_config(overrides: dict[str, Any] | None = None) -> mesh.RelaxMesh.Config:
  cfg = mesh.RelaxMesh.Config(integration_config=integration_config())
  return dataclasses.replace(cfg, **overrides) if overrides else cfg
