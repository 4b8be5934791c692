"""`RelaxMesh` plugin on the B200 backend.

Drop-in for the solver-facing part of the reference's `processor.mesh.RelaxMesh`
(processor/mesh.py:107-557): `SolutionStatus`, `Config` (the fields the solver
needs), `relax_mesh(x, prev, integration_config, mask)` with the reference's
fold-detect / retry logic (processor/mesh.py:428-513), `run_relaxation` and
`process`.  Building `prev` from flow volumes (`get_prev_state`,
processor/mesh.py:279-398) composes coordinate maps on the CPU
(`map_utils.compose_maps_fast`), which is outside this backend: it is
a hook here, exactly like the reference's own `_open_volume` / `_build_mask` hooks.
"""

from __future__ import annotations

import dataclasses
import enum
import logging
from typing import Any, Sequence

import numpy as np
from scipy import ndimage

from .. import compat
from .. import mesh as mesh_lib

Subvolume = compat.Subvolume


class SolutionStatus(enum.IntEnum):
  UNDEFINED = -1
  REGULAR = 0
  PREP_FAILED = 1
  REGULARIZED = 2


class MeshInitState(enum.Enum):
  ZEROS = 0
  PREV_MEDIAN = 1


def apply_mask(flow: np.ndarray, mask: np.ndarray):
  """Sets masked entries of all channels to NaN in place (flow_utils.py:32-34)."""
  for i in range(flow.shape[0]):
    flow[i, ...][mask] = np.nan


def mask_irregular(coord_map: np.ndarray, stride: Sequence[float], frac: float,
                   max_frac: float | None = None, dilation_iters: int = 1) -> np.ndarray:
  """Marks stretched / folded nodes of a [2, y, x] relative map (map_utils.py:737-786).

  A node is bad if the distance to its +x or +y neighbour is below `frac` or above
  `max_frac` (default 2 - frac) times the stride; the bad set is dilated with a full
  3x3 structuring element.  Bad nodes are set to NaN in place; returns the mask.
  """
  assert coord_map.ndim == 3 and coord_map.shape[0] == 2
  sx, sy = float(stride[0]), float(stride[1])
  hi = 2 - frac if max_frac is None else max_frac
  gap_x = np.zeros(coord_map.shape[1:], coord_map.dtype)
  gap_y = np.zeros(coord_map.shape[1:], coord_map.dtype)
  gap_x[:, :-1] = np.diff(coord_map[0], axis=-1)
  gap_y[:-1, :] = np.diff(coord_map[1], axis=-2)
  gap_x += sx
  gap_y += sy
  with np.errstate(invalid='ignore'):
    bad = (gap_x < frac * sx) | (gap_y < frac * sy) | (gap_x > hi * sx) | (gap_y > hi * sy)
  if dilation_iters > 0:
    bad = ndimage.binary_dilation(bad, ndimage.generate_binary_structure(2, 2),
                                  iterations=dilation_iters)
  coord_map[0][bad] = np.nan
  coord_map[1][bad] = np.nan
  return bad


@dataclasses.dataclass(frozen=True)
class MeshOptions:
  init_state: MeshInitState = MeshInitState.ZEROS


class RelaxMesh(compat.SubvolumeProcessor):
  """Finds the equilibrium mesh of one section against its reference section(s)."""

  @dataclasses.dataclass
  class Config:
    """Solver-facing subset of the reference's RelaxMesh.Config (processor/mesh.py:111-161)."""
    integration_config: mesh_lib.IntegrationConfig
    output_dir: str = 'NONE'
    mesh: Any = None
    flows: Sequence[Any] = ()
    mask: Any = None
    block_starts: Sequence[int] = ()
    block_ends: Sequence[int] = ()
    mesh_min_frac: float = 0.5
    mesh_max_frac: float = 2.0
    options: MeshOptions = MeshOptions()

  crop_at_borders = False

  def __init__(self, config: 'RelaxMesh.Config', input_ts_spec=None):
    del input_ts_spec
    self._config = config

  # ---- hooks ---------------------------------------------------------------------------
  def _build_mask(self, mask_configs, box):
    raise NotImplementedError('This function needs to be defined in a subclass.')

  def get_prev_state(self, stride, box):
    raise NotImplementedError(
        'Composing the reference-section maps (map_utils.compose_maps_fast) is not '
        'part of the CUDA backend; provide `prev` by overriding get_prev_state().')

  def maybe_update_init_state(self, state, prev, options: MeshOptions):
    """processor/mesh.py: initial state from the median of `prev` if requested."""
    if options.init_state == MeshInitState.PREV_MEDIAN and prev is not None:
      with np.errstate(all='ignore'):
        med = np.nanmedian(prev, axis=(1, 2, 3))
      state = state + np.nan_to_num(med)[:, None, None, None]
    return state

  def get_mesh_state(self, box, stride, prev):
    del stride
    state = np.zeros((2, 1, int(box.size[1]), int(box.size[0])))
    return self.maybe_update_init_state(state, prev, self._config.options)

  # ---- solver ---------------------------------------------------------------------------
  def relax_mesh(self, x, prev, integration_config, mask):
    """Mesh relaxation with fold detection and one regularised retry.

    Same contract as processor/mesh.py:428-513: returns
    (x [2, 1, y, x], e_kin history, steps simulated, SolutionStatus).
    """
    cfg = self._config
    if mask is not None:
      apply_mask(x, mask)
    logging.info('Starting mesh relaxation with: %r', cfg)

    x, e_kin, steps = mesh_lib.relax_mesh(x, prev, integration_config)
    x = np.array(x)
    first = x.copy()
    folded = mask_irregular(x[:, 0], integration_config.stride, cfg.mesh_min_frac,
                            dilation_iters=5)
    if not folded.any():
      return x, e_kin, steps, SolutionStatus.REGULAR

    logging.info('Attempting relaxation with 10% k0.')
    # Pull a fresh mesh towards the first solution (which now has NaN around the
    # irregular nodes) with weak springs; if that is regular, solve again from it.
    start = self.maybe_update_init_state(np.zeros_like(x), prev, cfg.options)
    soft = dataclasses.replace(integration_config, k0=integration_config.k0 / 10.0)
    x, _, prep_steps = mesh_lib.relax_mesh(start, x, soft)
    x = np.array(x)
    if mask_irregular(x[:, 0], integration_config.stride, cfg.mesh_min_frac).any():
      return first, e_kin, steps + prep_steps, SolutionStatus.PREP_FAILED

    if mask is not None:
      apply_mask(x, mask)
    x, e_kin2, reg_steps = mesh_lib.relax_mesh(x, prev, integration_config)
    return (np.array(x), e_kin2, steps + prep_steps + reg_steps,
            SolutionStatus.REGULARIZED)

  def run_relaxation(self, bbox):
    cfg = self._config
    z = int(bbox.start[2])
    ic = cfg.integration_config
    prev = mask = None
    if z not in cfg.block_starts:  # the first section of a block is not optimised
      if cfg.mask is not None:
        mask = self._build_mask(cfg.mask, bbox)
      prev = self.get_prev_state(ic.stride, bbox)
    x = self.get_mesh_state(bbox, ic.stride, prev)
    e_kin, steps, status = [], 0, SolutionStatus.UNDEFINED
    if (z not in cfg.block_starts and prev is not None and not np.all(np.isnan(x))
        and not np.all(np.isnan(prev))):
      x, e_kin, steps, status = self.relax_mesh(x, prev, ic, mask)
    return x, e_kin, steps, status

  def process(self, subvol):
    x, *_ = self.run_relaxation(subvol.bbox)
    return Subvolume(x, subvol.bbox)
