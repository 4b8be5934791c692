"""Configuration of the rendering pipeline (reference pipeline/warp_config.py:15-50)."""

from __future__ import annotations

import dataclasses
from typing import Any

from ..compat import config as cfg_lib
from ..processor import warp
from ..processor.defaults import em_2d


@dataclasses.dataclass(frozen=True)
class WarpPipelineConfig(cfg_lib.JsonMixin):
  """Pipeline configuration for warping a volume (pipeline/warp_config.py:28-32)."""
  warp: warp.WarpByMap.Config


def default_em_2d(overrides: dict[str, Any] | None = None) -> WarpPipelineConfig:
  """Default warp configuration for EM 2D data (pipeline/warp_config.py:35-44)."""
  config = WarpPipelineConfig(warp=em_2d.warp_config())
  return cfg_lib.update_dataclass(config, overrides)


cfg_lib.register_default_config(cfg_lib.DefaultConfigType.EM_2D, WarpPipelineConfig,
                                default_em_2d)
