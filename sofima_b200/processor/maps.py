"""Configuration surface of the reference's processor/maps.py that the EM-2D mesh pipeline
config refers to (pipeline/mesh_config.py:34-43).

`ReconcileCrossBlockMaps` itself -- the host-side reconciliation of block-wise and cross-block
coordinate maps with the reference's CPU map inversion (processor/maps.py:35-330) -- is
outside the hot-path scope of this backend: only its `Config` is
provided so that `pipeline.mesh_config.default_em_2d()` round-trips; constructing the
processor raises."""

from __future__ import annotations

import dataclasses
from typing import Any

from .. import compat
from ..compat import config as cfg_lib


class ReconcileCrossBlockMaps(compat.SubvolumeProcessor):
  """processor/maps.py:35 (configuration only, see the module docstring)."""

  @dataclasses.dataclass(eq=True)
  class Config(cfg_lib.JsonMixin):
    """Same fields as the reference's Config (processor/maps.py:55-84)."""
    cross_block: Any
    cross_block_inv: Any
    last_inv: Any
    main_inv: Any
    z_map: dict[str, int]
    stride: int
    xy_overlap: int = 128
    backward: bool = False

  crop_at_borders = False

  def __init__(self, config: 'ReconcileCrossBlockMaps.Config', input_volinfo=None):
    del config, input_volinfo
    raise NotImplementedError(
        'ReconcileCrossBlockMaps runs on the reference\'s CPU map inversion '
        '(map_utils.invert_map); it is outside the scope of the CUDA backend.')
