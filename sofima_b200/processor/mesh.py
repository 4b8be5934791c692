"""`RelaxMesh` plugin on the B200 backend.

Drop-in for the reference's `processor.mesh.RelaxMesh` (processor/mesh.py:107-557):
`SolutionStatus`, `MeshInitState`, `FlowVolume`, `BadSectionRange`, `MeshOptions`,
`ComingIn`, `Config`, the reference-section logic (`compute_ref_mesh`,
`compute_ref_mesh_multiz`, `get_prev_state`, `get_mesh_state`,
processor/mesh.py:169-425), `relax_mesh(x, prev, integration_config, mask)` with the
fold-detect / retry logic (processor/mesh.py:428-513), `run_relaxation`, `process`.

The two device steps -- composing the flow with the solved reference mesh
(`map_utils.compose_maps_fast`) and the relaxation itself (`mesh.relax_mesh`) -- run
on the GPU through the C ABI.  Volume / mask / tile I/O stays behind the same hooks the
reference leaves to its (un-vendored) `connectomics` base class: `_open_volume`,
`_build_mask`, `_load_stitched_tile`.
"""

from __future__ import annotations

import bisect
import dataclasses
import enum
import logging
from typing import Any, Sequence

import numpy as np
from scipy import ndimage

from .. import compat
from .. import map_utils
from .. import mesh as mesh_lib
from . import client_utils

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


# reference processor/mesh.py calls map_utils.mask_irregular; kept importable from here too
mask_irregular = map_utils.mask_irregular


get_block_id = client_utils.get_block_id  # reference: processor/client_utils.py


@dataclasses.dataclass(frozen=True)
class FlowVolume:
  """A flow volume and the section offset it was computed against."""
  delta_z: int
  volume: Any


@dataclasses.dataclass(frozen=True)
class BadSectionRange:
  """Skipped sections [start, end] and the flow bridging them (processor/mesh.py:60-77)."""
  start: int
  end: int
  flow: FlowVolume


@dataclasses.dataclass(frozen=True)
class MeshOptions:
  init_state: MeshInitState = MeshInitState.ZEROS
  irregular_mask_radius: int | None = None


@dataclasses.dataclass(frozen=True)
class ComingIn:
  """First complete section after a coming-in region and its 3-channel flow."""
  z: int
  flow: Any


class RelaxMesh(compat.SubvolumeProcessor):
  """Finds the equilibrium mesh of one section against its reference section(s)."""

  @dataclasses.dataclass(eq=True)
  class Config:
    """Same fields as the reference's RelaxMesh.Config (processor/mesh.py:111-161)."""
    integration_config: mesh_lib.IntegrationConfig
    output_dir: str = 'NONE'
    mesh: Any = None
    flows: Sequence[FlowVolume] = ()
    sections_to_skip: Sequence[int] = ()
    ranges_to_skip: Sequence[BadSectionRange] = ()
    mask: Any = None
    block_starts: Sequence[int] = ()
    block_ends: Sequence[int] = ()
    backward: bool = False
    mesh_min_frac: float = 0.5
    mesh_max_frac: float = 2.0
    coming_in: Sequence[ComingIn] = ()
    options: MeshOptions | None = dataclasses.field(default_factory=MeshOptions)

  crop_at_borders = False

  def __init__(self, config: 'RelaxMesh.Config', input_ts_spec=None):
    del input_ts_spec
    self._config = config

  # ---- I/O hooks (the reference leaves the same ones to its base class) ----------------
  def _build_mask(self, mask_configs, box):
    raise NotImplementedError('This function needs to be defined in a subclass.')

  def _open_volume(self, volume):
    """Returns an object indexable with a 4-d slice, with `.meta.num_channels`."""
    raise NotImplementedError('This function needs to be defined in a subclass.')

  def _load_stitched_tile(self, output_dir, box):
    raise NotImplementedError('This function needs to be defined in a subclass.')

  # ---- reference sections ---------------------------------------------------------------
  def is_skipped_section(self, z: int) -> bool:
    cfg = self._config
    return z in cfg.sections_to_skip or any(r.start <= z <= r.end for r in cfg.ranges_to_skip)

  def _solved_mesh(self, ref_box, allow_missing: bool):
    """Solved mesh of a reference section: this run's output, else the input mesh."""
    cfg = self._config
    ref_mesh = self._load_stitched_tile(cfg.output_dir, ref_box)
    if ref_mesh is None:
      if not allow_missing:
        raise ValueError(f'Missing previous mesh data for {ref_box.start}')
      assert cfg.mesh is not None
      ref_mesh = self._open_volume(cfg.mesh)[ref_box.to_slice4d()]
    return ref_mesh

  def compute_ref_mesh(self, flow: np.ndarray, ref_box, stride: Sequence[float]) -> np.ndarray:
    """Node targets from a 2-channel flow against one solved reference section
    (processor/mesh.py:248-277): flow composed with the reference's mesh on the GPU."""
    cfg = self._config
    ref_mesh = self._solved_mesh(ref_box, allow_missing=True)
    if cfg.mesh is not None:
      apply_mask(ref_mesh, self._build_mask(cfg.mask, ref_box))
    start = ref_box.start[::-1]
    return np.array(map_utils.compose_maps_fast(flow, start, stride, ref_mesh, start, stride))

  def compute_ref_mesh_multiz(self, flow: np.ndarray, box, starts: Sequence[int],
                              stride: Sequence[float], ignore_xblock: bool = True,
                              allow_missing_mesh: bool = True) -> np.ndarray:
    """Node targets from a 3-channel flow whose 3rd channel names, per node, the
    section offset it was estimated against (processor/mesh.py:169-238)."""
    cfg = self._config
    offsets = np.unique(flow[2, 0, :, :])
    offsets = offsets[np.isfinite(offsets) & (offsets != 0)].astype(np.int32).tolist()
    target = np.full([2] + list(flow.shape[1:]), np.nan)
    z = int(box.start[2])
    here = get_block_id(z, starts, cfg.backward)
    for delta_z in sorted(offsets, key=abs):
      if get_block_id(z - delta_z, starts, cfg.backward) != here:
        if ignore_xblock:
          break
        raise ValueError(f'Mesh data needs to be within a single block ({z} vs {z - delta_z}.')
      ref_box = box.translate(-np.array([0, 0, delta_z]))
      logging.info('Attempting to load ref. mesh for %r', ref_box)
      ref_mesh = self._solved_mesh(ref_box, allow_missing=allow_missing_mesh)
      if cfg.mask is not None:
        apply_mask(ref_mesh, self._build_mask(cfg.mask, ref_box))
      chosen = flow[2, ...] == delta_z
      part = flow[:2, ...].copy()
      part[0, ...][~chosen] = np.nan
      part[1, ...][~chosen] = np.nan
      start = box.start[::-1]
      part = np.array(map_utils.compose_maps_fast(part, start, stride, ref_mesh, start, stride))
      target[0, ...][chosen] = part[0, ...][chosen]
      target[1, ...][chosen] = part[1, ...][chosen]
    return target

  def get_prev_state(self, stride: Sequence[float], bbox):
    """Reference node positions of the section in `bbox` (processor/mesh.py:279-383).

    Targets from several reference sections are averaged (Hooke's law is linear);
    nodes whose neighbourhood in the averaged map is folded or over-stretched are
    removed.  Returns None for the first section of a block.
    """
    cfg = self._config
    z = int(bbox.start[2])
    starts = sorted(cfg.block_starts)
    if z in starts:
      return None

    for cin in cfg.coming_in:
      if z == cin.z:
        flow = self._open_volume(cin.flow)[bbox.to_slice4d()]
        return self.compute_ref_mesh_multiz(flow, bbox, starts, stride, ignore_xblock=False,
                                            allow_missing_mesh=False)

    flows = cfg.flows
    before = z + 1 if cfg.backward else z - 1
    for rng in cfg.ranges_to_skip:  # right after a skipped range: its bridging flow
      if before == rng.end:
        flows = [rng.flow]
        break

    here = get_block_id(z, starts, cfg.backward)
    ny, nx = int(bbox.size[1]), int(bbox.size[0])
    total = np.zeros((2, 1, ny, nx))
    count = np.zeros((ny, nx), dtype=np.int32)
    used = 0
    for fv in flows:
      ref_z = z - fv.delta_z
      if self.is_skipped_section(ref_z):
        continue
      if get_block_id(ref_z, starts, cfg.backward) != here:
        continue
      volume = self._open_volume(fv.volume)
      field = volume[bbox.to_slice4d()]
      if volume.meta.num_channels == 2:
        ref_box = bbox.translate(-np.array([0, 0, fv.delta_z]))
        ref_mesh = self.compute_ref_mesh(field, ref_box, stride)
      else:
        ref_mesh = self.compute_ref_mesh_multiz(field, bbox, starts, stride)
      count += np.isfinite(ref_mesh[0, 0, ...]).astype(np.int32)
      total += np.nan_to_num(ref_mesh)
      used += 1
    if used == 0:
      return None

    weight = count.astype(np.float32)
    weight[weight == 0] = np.nan
    prev = total / weight[np.newaxis, np.newaxis, :, :]
    radius = 1
    if cfg.options and cfg.options.irregular_mask_radius is not None:
      radius = cfg.options.irregular_mask_radius
    mask_irregular(prev[:, 0, ...], stride, cfg.mesh_min_frac, cfg.mesh_max_frac,
                   dilation_iters=radius)
    return prev

  def maybe_update_init_state(self, x, prev, options: MeshOptions | None):
    """processor/mesh.py:385-395: start from the median of `prev` if requested."""
    if options is not None and options.init_state == MeshInitState.PREV_MEDIAN and prev is not None:
      with np.errstate(all='ignore'):
        x[0, ...] = np.nanmedian(prev[0, ...])
        x[1, ...] = np.nanmedian(prev[1, ...])
      x = np.nan_to_num(x)
    return x

  def get_mesh_state(self, box, stride, prev):
    """Initial state of the section to optimise (processor/mesh.py:397-425)."""
    cfg = self._config
    shape = (2, 1, int(box.size[1]), int(box.size[0]))
    if cfg.mesh is None:
      return np.zeros(shape)
    state = self._open_volume(cfg.mesh)[box.to_slice4d()]
    bad = mask_irregular(state[:, 0, ...], stride, cfg.mesh_min_frac, cfg.mesh_max_frac,
                         dilation_iters=0)
    if np.any(bad):
      state = self.maybe_update_init_state(np.zeros(shape), prev, cfg.options)
    return state

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

    # The three solves and the irregularity tests between them run on device-resident
    # tensors (mesh solver, csrc/flowfilt.cu mask_irregular): the mesh crosses PCIe once in
    # each direction, not after every step as upstream (processor/mesh.py:462-471).
    import torch
    from .. import _native
    dev = torch.device('cuda', _native.Context.get().device)
    to_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
    xd, pd = to_dev(x), to_dev(prev)
    xd, e_kin, steps = mesh_lib.relax_mesh(xd, pd, integration_config)
    first = xd.clone()
    sec = xd[:, 0].contiguous()
    folded = mask_irregular(sec, integration_config.stride, cfg.mesh_min_frac,
                            dilation_iters=5)
    if not bool(folded.any()):
      return mesh_lib._to_host(first), e_kin, steps, SolutionStatus.REGULAR
    xd[:, 0] = sec

    logging.info('Attempting relaxation with 10% k0.')
    # Pull a fresh mesh towards the first solution (which now has NaN around the
    # irregular nodes) with weak springs; if that is regular, solve again from it.
    start = self.maybe_update_init_state(np.zeros_like(x), prev, cfg.options)
    soft = dataclasses.replace(integration_config, k0=integration_config.k0 / 10.0)
    xd, _, prep_steps = mesh_lib.relax_mesh(to_dev(start), xd, soft)
    sec = xd[:, 0].contiguous()
    if bool(mask_irregular(sec, integration_config.stride, cfg.mesh_min_frac).any()):
      return mesh_lib._to_host(first), e_kin, steps + prep_steps, SolutionStatus.PREP_FAILED
    xd[:, 0] = sec

    if mask is not None:
      xd[:, torch.from_numpy(np.ascontiguousarray(mask)).to(dev)] = float('nan')
    xd, e_kin2, reg_steps = mesh_lib.relax_mesh(xd, pd, integration_config)
    return (mesh_lib._to_host(xd), e_kin2, steps + prep_steps + reg_steps,
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
