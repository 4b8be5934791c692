"""Configuration of the mesh relaxation pipeline (reference pipeline/mesh_config.py:15-67)."""

from __future__ import annotations

import dataclasses
from typing import Any

from ..compat import config as cfg_lib
from ..processor import maps
from ..processor import mesh
from ..processor.defaults import em_2d


@dataclasses.dataclass(frozen=True)
class MeshRelaxationConfig(cfg_lib.JsonMixin):
  """Pipeline configuration for mesh relaxation (pipeline/mesh_config.py:33-40)."""
  within_block_config: mesh.RelaxMesh.Config
  last_section_config: mesh.RelaxMesh.Config
  cross_block_config: mesh.RelaxMesh.Config
  reconcile_cross_block_config: maps.ReconcileCrossBlockMaps.Config


def default_em_2d(overrides: dict[str, Any] | None = None) -> MeshRelaxationConfig:
  """Default mesh relaxation configuration for EM 2D data (pipeline/mesh_config.py:43-60)."""
  config = MeshRelaxationConfig(
      within_block_config=em_2d.within_block_config(),
      last_section_config=em_2d.last_section_config(),
      cross_block_config=em_2d.cross_block_config(),
      reconcile_cross_block_config=em_2d.default_em_2d_reconcile_config())
  return cfg_lib.update_dataclass(config, overrides)


cfg_lib.register_default_config(cfg_lib.DefaultConfigType.EM_2D, MeshRelaxationConfig,
                                default_em_2d)
