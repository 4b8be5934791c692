"""`EstimateFlow` plugin on the B200 backend.

Drop-in for the reference's `processor.flow.EstimateFlow` (processor/flow.py:43-275):
same `Config` fields, same geometry contract (`context`, `subvolume_size`,
`overlap`, `expected_output_box`, `num_channels`, `pixelsize`, `output_type`) and
the same `process(Subvolume) -> Subvolume` result layout ([4, Z', gy, gx] float32,
channels x, y, sharpness, peak ratio).  The per-section-pair arithmetic goes to
`sofima_b200.flow_field` (CUDA).  `EstimateMissingFlow` / `ReconcileAndFilterFlows`
are CPU-side map algebra on top of this path and are not part of this backend.
"""

from __future__ import annotations

import dataclasses
from typing import Any

import numpy as np

from .. import compat
from .. import flow_field

BoundingBox = compat.BoundingBox
Subvolume = compat.Subvolume


class EstimateFlow(compat.SubvolumeProcessor):
  """Section-to-section optical flow: p(z) + f(z) <-> p(z - dz)."""

  @dataclasses.dataclass(eq=True)
  class Config:
    """Same fields as the reference's EstimateFlow.Config (processor/flow.py:64-95)."""
    patch_size: int
    stride: int
    z_stride: int
    fixed_current: bool
    mask_configs: Any
    mask_only_for_patch_selection: bool
    selection_mask_configs: Any
    batch_size: int

    def to_dict(self):
      return dataclasses.asdict(self)

    @classmethod
    def from_dict(cls, d):
      return cls(**d)

  def __init__(self, config: 'EstimateFlow.Config', input_volinfo_or_ts_spec=None):
    del input_volinfo_or_ts_spec
    if config.patch_size % config.stride != 0:
      raise AssertionError('patch_size must be divisible by stride')
    self._config = config
    if isinstance(config.mask_configs, str) and config.mask_configs:
      config.mask_configs = self._get_mask_configs(config.mask_configs)
    if isinstance(config.selection_mask_configs, str) and config.selection_mask_configs:
      config.selection_mask_configs = self._get_mask_configs(
          config.selection_mask_configs)

  # ---- hooks a deployment overrides (as with the reference) ----------------------
  def _get_mask_configs(self, text: str):
    raise NotImplementedError('mask config parsing is provided by the host framework')

  def _build_mask(self, mask_configs, box):
    raise NotImplementedError('This function needs to be defined in a subclass.')

  # ---- geometry contract -------------------------------------------------------------
  def output_type(self, input_type):
    return np.float32

  def subvolume_size(self):
    side = self._config.patch_size * 8
    return compat.SuggestedXyz(side, side, 16)

  def context(self):
    c = self._config
    lo = c.patch_size // 2
    hi = c.patch_size - lo
    dz = abs(c.z_stride)
    # The reference section lies dz before (z_stride > 0) or after the current one;
    # with fixed_current the roles of the two ends are swapped.
    z_before = (c.z_stride > 0) != bool(c.fixed_current)
    if z_before:
      return (lo, lo, dz), (hi, hi, 0)
    return (lo, lo, 0), (hi, hi, dz)

  def num_channels(self, input_channels):
    del input_channels
    return flow_field.JAXMaskedXCorrWithStatsCalculator.non_spatial_flow_channels + 2

  def pixelsize(self, psize):
    out = np.array(psize, dtype=np.float32)
    out[:2] *= self._config.stride
    return out

  def overlap(self):
    ox, oy, oz = super().overlap()
    return ox - self._config.stride, oy - self._config.stride, oz

  def expected_output_box(self, box):
    c = self._config
    inner = self.crop_box(box)
    scaled = inner.scale([1.0 / c.stride, 1.0 / c.stride, 1.0])
    size = scaled.size.copy()
    size[:2] = (np.array(self.subvolume_size()[:2]) - c.patch_size + c.stride) // c.stride
    return BoundingBox(start=scaled.start, size=size)

  # ---- work ----------------------------------------------------------------------------
  def _section_pairs(self, nz: int):
    """(z_prev, z_curr) pairs in output order (processor/flow.py:213-229)."""
    c = self._config
    if c.fixed_current:
      if c.z_stride > 0:
        return [(z, nz - 1) for z in range(0, nz - 1)]
      return [(z, 0) for z in range(1, nz)]
    if c.z_stride > 0:
      return [(z, z + c.z_stride) for z in range(0, nz - c.z_stride)]
    return [(z, z + c.z_stride) for z in range(-c.z_stride, nz)]

  def process(self, subvol):
    c = self._config
    box = subvol.bbox
    data = subvol.data
    compat.counter(self.namespace, 'subvolumes-started').inc()
    assert data.shape[0], 'Input volume should have 1 channel.'
    image = data[0]

    mask = sel_mask = None
    with compat.timer_counter(self.namespace, 'build-mask'):
      if c.mask_configs:
        mask = self._build_mask(c.mask_configs, box)
      if c.selection_mask_configs:
        sel_box = box.scale([1.0 / c.stride, 1.0 / c.stride, 1])
        sel_mask = self._build_mask(c.selection_mask_configs, sel_box)

    calc = flow_field.JAXMaskedXCorrWithStatsCalculator()
    flows = []
    with compat.timer_counter(self.namespace, 'flow'):
      for z_prev, z_curr in self._section_pairs(image.shape[0]):
        flows.append(calc.flow_field(
            image[z_prev], image[z_curr], c.patch_size, c.stride,
            None if mask is None else mask[z_prev],
            None if mask is None else mask[z_curr],
            mask_only_for_patch_selection=c.mask_only_for_patch_selection,
            selection_mask=None if sel_mask is None else sel_mask[z_curr],
            batch_size=c.batch_size))
    ret = np.array(flows)  # [Z', 4, gy, gx]

    inner = self.crop_box(box)
    out_box = BoundingBox(start=inner.start // [c.stride, c.stride, 1],
                          size=[ret.shape[-1], ret.shape[-2], inner.size[2]])
    if ret.shape[0] != out_box.size[2]:
      raise ValueError(f'ret:{ret.shape} vs out:{out_box.size}')
    compat.counter(self.namespace, 'subvolumes-done').inc()
    return Subvolume(np.transpose(ret, (1, 0, 2, 3)), out_box)
