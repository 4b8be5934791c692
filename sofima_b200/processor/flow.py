"""`EstimateFlow` plugin on the B200 backend.

Drop-in for the reference's `processor.flow.EstimateFlow` (processor/flow.py:43-275):
same `Config` fields, same geometry contract (`context`, `subvolume_size`,
`overlap`, `expected_output_box`, `num_channels`, `pixelsize`, `output_type`) and
the same `process(Subvolume) -> Subvolume` result layout ([4, Z', gy, gx] float32,
channels x, y, sharpness, peak ratio).  The per-section-pair arithmetic goes to
`sofima_b200.flow_field` (CUDA).

`EstimateMissingFlow` (processor/flow.py:496-844) re-estimates the invalid vectors of
a flow volume against earlier sections with a larger search patch; its inner call is
the same CUDA flow path.  `ReconcileAndFilterFlows` (processor/flow.py:279-493) is
host-side filtering (`flow_utils`) in the reference too.  Image / mask / metadata I/O
stays behind the hooks the reference leaves to its `connectomics` base class
(`_open_volume`, `_build_mask`, `_get_mask_configs`, `_get_metadata`).
"""

from __future__ import annotations

import dataclasses
import logging
from typing import Any, Sequence

import numpy as np

from .. import compat
from .. import flow_field
from .. import flow_utils

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


class ReconcileAndFilterFlows(compat.SubvolumeProcessor):
  """Filters 4- or 3-channel flow volumes and merges several estimates
  (processor/flow.py:279-493).

  Every input volume is cleaned with `flow_utils.clean_flow`; the cleaned fields are
  merged in order of preference with `flow_utils.reconcile_flows`.  Estimates at a
  coarser pixel size are brought to the base grid by `_upsample_flow`, a hook here:
  the reference does that with its CPU map resampler (`map_utils.resample_map`,
  SciPy), which is outside this backend.
  """

  crop_at_borders = False

  @dataclasses.dataclass(eq=True)
  class Config:
    """Same fields as the reference's Config (processor/flow.py:292-329)."""
    flow_volinfos: Sequence[str] | str | None
    mask_configs: Any
    min_peak_ratio: float
    min_peak_sharpness: float
    max_magnitude: float
    max_deviation: float
    max_gradient: float
    min_patch_size: int
    multi_section: bool
    base_delta_z: int

  def __init__(self, config: 'ReconcileAndFilterFlows.Config', input_path_or_metadata=None):
    self._config = config
    self._metadata = []
    self._scales = []
    if input_path_or_metadata is not None:
      meta = input_path_or_metadata
      if isinstance(meta, str):
        meta = self._get_metadata(meta)
      self._scales.append(None)
      self._metadata.append(meta)
    if isinstance(config.flow_volinfos, str):
      config.flow_volinfos = config.flow_volinfos.split(',')
    if config.flow_volinfos is None:
      config.flow_volinfos = []
    for entry in config.flow_volinfos:
      path, _, scale = entry.partition(':')
      self._scales.append(float(scale) if scale else None)
      self._metadata.append(self._get_metadata(path))
    for fine, coarse in zip(self._metadata, self._metadata[1:]):  # ascending voxel size
      assert fine.pixel_size.x <= coarse.pixel_size.x
      assert fine.pixel_size.y <= coarse.pixel_size.y
      assert fine.pixel_size.x / coarse.pixel_size.x == fine.pixel_size.y / coarse.pixel_size.y
      assert fine.pixel_size.z == coarse.pixel_size.z
    if config.mask_configs and isinstance(config.mask_configs, str):
      config.mask_configs = self._get_mask_configs(config.mask_configs)

  # ---- hooks ---------------------------------------------------------------------------
  def _get_metadata(self, path):
    raise NotImplementedError('This function needs to be defined in a subclass.')

  def _get_mask_configs(self, text: str):
    raise NotImplementedError('mask config parsing is provided by the host framework')

  def _build_mask(self, mask_configs, box):
    raise NotImplementedError('This function needs to be defined in a subclass.')

  def _open_volume(self, path):
    raise NotImplementedError('This function needs to be defined in a subclass.')

  def _upsample_flow(self, flow, read_box, box, scale, mag_scale):
    raise NotImplementedError(
        'Upsampling a coarser flow estimate to the base grid uses the reference\'s CPU '
        'map resampler (map_utils.resample_map); provide it in a subclass.')

  def num_channels(self, input_channels=0):
    del input_channels
    return 3 if self._config.multi_section else 2

  def process(self, subvol):
    cfg = self._config
    box = subvol.bbox
    mask = self._build_mask(cfg.mask_configs, box) if cfg.mask_configs else None
    flows = []
    for i, (meta, mag_scale) in enumerate(zip(self._metadata, self._scales)):
      vol = self._open_volume(meta.path)
      scale = 1 if i == 0 else self._metadata[0].pixel_size.x / meta.pixel_size.x
      read_box = box
      if i > 0:
        assert scale <= 1.0
        read_box = box.scale((scale, scale, 1))
        if scale < 1:
          read_box = read_box.adjusted_by(start=-np.asarray(self._context[0]),
                                          end=np.asarray(self._context[1]))
        read_box = vol.clip_box_to_volume(read_box)
        assert read_box is not None
      flow = flow_utils.clean_flow(vol[read_box.to_slice4d()], cfg.min_peak_ratio,
                                   cfg.min_peak_sharpness, cfg.max_magnitude,
                                   cfg.max_deviation)
      if i == 0 or scale == 1:
        if cfg.multi_section and flow.shape[0] != 3:
          wide = np.full((3,) + flow.shape[1:], np.nan, dtype=flow.dtype)
          wide[:2, ...] = flow[:2, ...]
          wide[2, ...][np.isfinite(wide[0, ...])] = cfg.base_delta_z
          flow = wide
        flows.append(flow)
        continue
      hires = self._upsample_flow(flow, read_box, box, scale,
                                  scale if mag_scale is None else mag_scale)
      if mask is not None:
        flow_utils.apply_mask(hires, mask)
      flows.append(hires)
    merged = flow_utils.reconcile_flows(flows, cfg.max_gradient, cfg.max_deviation,
                                        cfg.min_patch_size)
    return self.crop_box_and_data(box, merged)


class EstimateMissingFlow(compat.SubvolumeProcessor):
  """Fills the invalid (NaN) vectors of a single-section flow volume by estimating
  flow against earlier sections (processor/flow.py:496-844).

  Output channels: flow x, flow y, and the section offset the vector was finally
  estimated against.
  """

  @dataclasses.dataclass(frozen=True)
  class Config:
    """Same fields as the reference's Config (processor/flow.py:503-555)."""
    patch_size: int
    stride: int
    delta_z: int
    max_delta_z: int
    max_attempts: int
    mask_configs: Any
    mask_only_for_patch_selection: bool
    selection_mask_configs: Any
    min_peak_ratio: float
    min_peak_sharpness: float
    max_magnitude: int
    batch_size: int
    image_volinfo: str | None
    image_cache_bytes: int
    mask_cache_bytes: int
    search_radius: int

  def __init__(self, config: 'EstimateMissingFlow.Config', input_volinfo_or_ts_spec=None):
    del input_volinfo_or_ts_spec
    if config.patch_size % config.stride != 0:
      raise ValueError(
          f'patch_size {config.patch_size} not a multiple of stride {config.stride}')
    self._search_patch_size = config.patch_size + 2 * config.search_radius
    if self._search_patch_size % config.stride != 0:
      raise ValueError(f'search_patch_size {self._search_patch_size} not a multiple of'
                       f' stride {config.stride}')
    if config.mask_configs and isinstance(config.mask_configs, str):
      config = dataclasses.replace(
          config, mask_configs=self._get_mask_configs(config.mask_configs))
    if config.selection_mask_configs and isinstance(config.selection_mask_configs, str):
      config = dataclasses.replace(
          config,
          selection_mask_configs=self._get_mask_configs(config.selection_mask_configs))
    self._config = config
    logging.info('EstimateMissingFlow running with config: %r', config)

  # ---- hooks ---------------------------------------------------------------------------
  def _get_mask_configs(self, text: str):
    raise NotImplementedError('mask config parsing is provided by the host framework')

  def _build_mask(self, mask_configs, box):
    """Returns a CZYX-shaped ndarray-like object."""
    raise NotImplementedError('This function needs to be defined in a subclass.')

  def _open_volume(self, path):
    """Image volume: `clip_box_to_volume(box)`, `.asarray` ([c, z, y, x])."""
    raise NotImplementedError('This function needs to be defined in a subclass.')

  def num_channels(self, input_channels):
    del input_channels
    return 3

  def _image_box(self, flow_box, patch: int, depth: int) -> BoundingBox:
    """Image region whose patches of size `patch` are centred on the flow nodes."""
    s = self._config.stride
    return BoundingBox(
        start=(flow_box.start[0] * s - patch // 2, flow_box.start[1] * s - patch // 2,
               flow_box.start[2]),
        size=((flow_box.size[0] - 1) * s + patch, (flow_box.size[1] - 1) * s + patch, depth))

  def process(self, subvol):
    cfg = self._config
    box, field = subvol.bbox, subvol.data
    ns = 'estimate-missing-flow'
    compat.counter(ns, 'subvolumes-started').inc()
    stride, search = cfg.stride, self._search_patch_size
    image_volume = self._open_volume(cfg.image_volinfo)

    # Image region of the 'previous' sections, which carries the search radius.
    wanted = self._image_box(box, search, 1)
    prev_box = image_volume.clip_box_to_volume(wanted)
    assert prev_box is not None
    if np.any(prev_box.size[:2] <= search):  # not enough image context for any node
      return subvol

    # Drop the flow nodes without full image context on the low ...
    lo = prev_box.translate(-wanted.start).start // stride
    out_box = box.adjusted_by(start=lo)
    field = field[:, :, lo[1]:, lo[0]:]
    # ... and high side (ceil division).
    hi = -((prev_box.end - wanted.end) // stride)
    out_box = out_box.adjusted_by(end=-hi)
    field = field[:, :, :out_box.size[1], :out_box.size[0]]

    ret = np.zeros([3] + list(out_box.size[::-1]))
    ret[:2, ...] = field
    ret[2, ...] = cfg.delta_z
    sel_mask = None
    if cfg.selection_mask_configs:
      sel_mask = self._build_mask(cfg.selection_mask_configs, out_box)

    calc = flow_field.JAXMaskedXCorrWithStatsCalculator()
    invalid = np.isnan(field[0, ...])
    curr_box = image_volume.clip_box_to_volume(
        self._image_box(out_box, cfg.patch_size, invalid.shape[0]))
    assert curr_box is not None

    if cfg.delta_z > 0:
      deltas = range(cfg.delta_z + 1, cfg.max_delta_z + 1)
      z_lo, z_hi = out_box.start[2] - cfg.max_delta_z, out_box.end[2]
    else:
      deltas = range(cfg.delta_z - 1, cfg.max_delta_z - 1, -1)
      z_lo, z_hi = out_box.start[2], out_box.end[2] - cfg.max_delta_z  # max_delta_z < 0
    load_box = image_volume.clip_box_to_volume(BoundingBox(
        start=(prev_box.start[0], prev_box.start[1], z_lo),
        size=(prev_box.size[0], prev_box.size[1], z_hi - z_lo)))
    logging.info('Loading image data: %r', load_box)
    stack = image_volume.asarray[load_box.to_slice4d()][0, ...]
    full_mask = self._build_mask(cfg.mask_configs, load_box) if cfg.mask_configs else None

    # The 'curr' image sits centred inside the (larger) 'prev' image.
    rel = curr_box.start - load_box.start
    curr_sel = (slice(rel[1], rel[1] + curr_box.size[1]),
                slice(rel[0], rel[0] + curr_box.size[0]))

    for z in range(invalid.shape[0]):
      if not invalid[z, ...].any():
        compat.counter(ns, 'sections-already-valid').inc()
        continue
      zi = (out_box.start[2] + z) - load_box.start[2]
      assert 0 <= zi < stack.shape[0]
      curr_mask = None
      if full_mask is not None:
        curr_mask = full_mask[zi, ...][curr_sel]
        if np.all(curr_mask):
          compat.counter(ns, 'sections-masked').inc()
          continue
      attempts = np.zeros(ret.shape[2:], dtype=int)
      todo = ~np.isfinite(ret[0, z, ...])
      if sel_mask is not None:
        todo &= sel_mask[z, ...]
      curr = stack[zi, ...][curr_sel]

      for delta_z in deltas:
        pi = zi - delta_z
        if pi < 0 or pi >= stack.shape[0]:
          break
        prev_mask = None
        if full_mask is not None:
          prev_mask = full_mask[pi, ...]
          if np.all(prev_mask):
            continue
        # Attempts only count where both sections are unmasked.
        todo &= attempts <= cfg.max_attempts
        if not todo.any():
          break
        logging.info('delta_z=%d: %d points to evaluate', delta_z, int(todo.sum()))
        flow = calc.flow_field(
            stack[pi, ...], curr, search, stride, prev_mask, curr_mask,
            mask_only_for_patch_selection=cfg.mask_only_for_patch_selection,
            selection_mask=todo, batch_size=cfg.batch_size, post_patch_size=cfg.patch_size)
        got = np.isfinite(flow[0, ...])
        attempts[:got.shape[0], :got.shape[1]][got] += 1
        flow = flow_utils.clean_flow(flow[:, np.newaxis, ...], cfg.min_peak_ratio,
                                     cfg.min_peak_sharpness, cfg.max_magnitude,
                                     max_deviation=0.0)
        sy, sx = flow.shape[2:]
        update = todo[:sy, :sx] & np.isfinite(flow[0, 0, ...])
        todo[:sy, :sx][update] = False
        compat.counter(ns, f'sections-filled-delta{delta_z}').inc(int(update.sum()))
        ret[2, z, :sy, :sx][update] = delta_z
        ret[0, z, :sy, :sx][update] = flow[0, 0, ...][update]
        ret[1, z, :sy, :sx][update] = flow[1, 0, ...][update]
    return Subvolume(ret, out_box)
