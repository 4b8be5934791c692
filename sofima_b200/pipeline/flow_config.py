"""Configuration of the end-to-end flow pipeline (reference pipeline/flow_config.py:15-106):
the same dataclasses and EM-2D defaults, so that an existing pipeline config selects the CUDA
backend by importing this package instead of `sofima`."""

from __future__ import annotations

import dataclasses
from typing import Any

from ..compat import config as cfg_lib
from ..processor import flow
from ..processor.defaults import em_2d


@dataclasses.dataclass(frozen=True)
class EstimateFlowStage(cfg_lib.JsonMixin):
  """pipeline/flow_config.py:32-39."""
  config: flow.EstimateFlow.Config
  processing: cfg_lib.ProcessingConfig
  schedule_batch_size: int
  ignore_existing: bool
  delete_existing: bool
  corner_whitelist: set


@dataclasses.dataclass(frozen=True)
class FlowPipeline(cfg_lib.JsonMixin):
  """Configuration for end-to-end SOFIMA flow estimation (pipeline/flow_config.py:42-49)."""
  estimate_flow: EstimateFlowStage
  reconcile_flows: flow.ReconcileAndFilterFlows.Config
  estimate_missing_flow: flow.EstimateMissingFlow.Config
  reconcile_missing_flows: flow.ReconcileAndFilterFlows.Config


def default_em_2d(overrides: dict[str, Any] | None = None) -> FlowPipeline:
  """Default flow pipeline configuration for EM 2D data (pipeline/flow_config.py:52-99)."""
  reconcile_missing_flows = em_2d.reconcile_missing_flows_config()
  estimate_flow_config = em_2d.estimate_flow_config()
  if (overrides is not None and 'estimate_flow' in overrides
      and 'config' in overrides['estimate_flow']):
    estimate_flow_config = cfg_lib.update_dataclass(
        estimate_flow_config, overrides['estimate_flow']['config'])
  config = FlowPipeline(
      estimate_flow=EstimateFlowStage(
          config=estimate_flow_config,
          processing=cfg_lib.ProcessingConfig(
              overlap=[160, 160, estimate_flow_config.z_stride],
              subvolume_size=[3200, 3200, 128]),
          schedule_batch_size=16384, corner_whitelist=set(), ignore_existing=False,
          delete_existing=False),
      reconcile_flows=em_2d.reconcile_flows_config(),
      estimate_missing_flow=em_2d.estimate_missing_flow_config(),
      reconcile_missing_flows=reconcile_missing_flows)
  return cfg_lib.update_dataclass(config, overrides)


cfg_lib.register_default_config(cfg_lib.DefaultConfigType.EM_2D, FlowPipeline, default_em_2d)
