"""Flow decorators on the B200 backend (reference decorators/flow.py).

The reference wraps TensorStore `virtual_chunked` views (`OptimFlow` :131,
`MeshRelaxFlowFilter` :108, `CleanFlowFilter` :47, `ReconcileFlowFilter` :358) over the
`Decorator` / `Filter` base classes of the un-vendored connectomics package.  Neither
tensorstore nor gin is installable in this image, so the four classes are built here over
the in-memory stand-ins of compat/volume.py (labelled arrays and a lazy chunked view with the
`read_fn(domain, array, read_params)` contract of `ts.virtual_chunked`): same constructor
arguments, same output domains / labels / chunking, same chunk functions (`optim_flow` and
`mesh_relax_flow` reach the CUDA hot path; `clean_flow`, `reconcile_flow` are host filters).
"""

from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from .. import flow_field
from .. import flow_utils
from .. import mesh
from ..compat import volume


def clean_flow(flow: np.ndarray, **filter_args) -> np.ndarray:
  """Chunk function of `CleanFlowFilter` (decorators/flow.py:38-43): [dim + 2, ...]
  flow with singleton non-spatial axes -> [dim, ...] cleaned vectors."""
  final_shape = list(flow.shape)
  final_shape[0] -= 2
  return flow_utils.clean_flow(flow.squeeze(), dim=flow.shape[0] - 2,
                               **filter_args).reshape(final_shape)


def reconcile_flow(flow: np.ndarray, **filter_args) -> np.ndarray:
  """Chunk function of `ReconcileFlowFilter` (decorators/flow.py:349-354)."""
  return flow_utils.reconcile_flows([flow.squeeze()], **filter_args).reshape(flow.shape)


def mesh_relax_flow(flow: np.ndarray, **filter_args) -> np.ndarray:
  """Chunk function of `MeshRelaxFlowFilter` (decorators/flow.py:96-105).

  `flow` is a [2 or 3, ...] flow field with singleton non-spatial axes; it becomes
  the `prev` target of a mesh that starts at zero and is relaxed with
  `IntegrationConfig(**filter_args)`.
  """
  cfg = mesh.IntegrationConfig(**filter_args)
  target = flow.squeeze()
  x0 = np.zeros_like(target)
  ncomp = flow.shape[0]
  if ncomp == 2:
    if target.ndim == 3:  # [2, y, x] -> [2, 1, y, x]
      res = mesh.relax_mesh(x0[:, None], target[:, None], cfg)
    else:
      res = mesh.relax_mesh(x0, target, cfg)
  elif ncomp == 3:
    res = mesh.relax_mesh(x0, target, cfg, mesh_force=mesh.elastic_mesh_3d)
  else:
    raise ValueError(f'`num_spatial_dim` must be 2 or 3 but is {ncomp}.')
  return np.asarray(res[0]).reshape(flow.shape)


def optim_flow(pre_image: np.ndarray, post_image: np.ndarray, patch_zyx: Sequence[int],
               step_zyx: Sequence[int], batch_size: int = 1024, pre_mask=None,
               post_mask=None, **flow_args) -> np.ndarray:
  """Chunk function of `OptimFlow` (decorators/flow.py:279-290): float32 images in
  ([z,] y, x) order -> flow field from post to pre."""
  calc = flow_field.JAXMaskedXCorrWithStatsCalculator()
  return calc.flow_field(
      pre_image=np.asarray(pre_image, dtype=np.float32),
      post_image=np.asarray(post_image, dtype=np.float32), pre_mask=pre_mask,
      post_mask=post_mask, patch_size=tuple(patch_zyx), step=tuple(step_zyx),
      batch_size=batch_size, **flow_args)


def _flow_shape(o, p, s):
  return np.ceil((o - p + 1) / s).astype(int)


def _padded_flow_shape(o, p, s):
  return _flow_shape(o, p, s) + p // s - 1


class CleanFlowFilter(volume.Filter):
  """Runs `clean_flow` over a flow volume (decorators/flow.py:47-86): the result has the two
  statistics channels removed from the `fc` dimension."""

  def __init__(self, min_chunksize: Optional[Sequence[int]] = None, context_spec=None,
               **filter_args):
    super().__init__(filter_fun=clean_flow, context_spec=context_spec,
                     min_chunksize=min_chunksize, **filter_args)

  def decorate(self, input_ts):
    dom = volume._domain_of(input_ts)
    d0 = dom[0]

    def filt_read(domain, array, unused_read_params):
      read_domain = list(domain)
      read_domain[0] = volume.Dim(0, input_ts.shape[0], d0.label)  # all flow channels
      array[...] = self._filter_fun(np.array(input_ts[volume.IndexDomain(read_domain)]),
                                    **self._filter_args)

    out_dims = list(dom)
    out_dims[0] = volume.Dim(d0.inclusive_min, d0.exclusive_max - 2, d0.label)
    chunk = self._chunk_shape(input_ts)
    chunk[0] = input_ts.shape[0] - 2
    return volume.VirtualChunked(filt_read, volume.IndexDomain(out_dims), chunk, input_ts.dtype)


class MeshRelaxFlowFilter(volume.Filter):
  """Regularises a flow field by relaxing a spring mesh towards it (decorators/flow.py:108-128);
  the relaxation runs on the CUDA solver."""

  def __init__(self, min_chunksize: Optional[Sequence[int]] = None, context_spec=None,
               **filter_args):
    super().__init__(filter_fun=mesh_relax_flow, context_spec=context_spec,
                     min_chunksize=min_chunksize, **filter_args)


class ReconcileFlowFilter(volume.Filter):
  """Runs `reconcile_flows` on a single flow field (decorators/flow.py:358-369)."""

  def __init__(self, min_chunksize: Optional[Sequence[int]] = None, context_spec=None,
               **filter_args):
    super().__init__(filter_fun=reconcile_flow, context_spec=context_spec,
                     min_chunksize=min_chunksize, **filter_args)


class OptimFlow(volume.Decorator):
  """Finds 2-d / 3-d flow for registration via cross-correlation (decorators/flow.py:131-355).

  Same constructor arguments and output layout as the reference: dimensions `fc, fz, fy, fx`
  followed by the input's non-image dimensions, `pad=True` frames the field with NaN so that
  it lines up with the image grid.  `fixed_spec` / `*_mask_spec` may be opened stores or
  `{'array': ..., 'labels': ...}` mappings (compat/volume.py); `jax_device` is accepted and
  ignored -- the computation runs on the current CUDA device.
  """

  def __init__(self, fixed_spec, image_dims: Sequence[str] = ('x', 'y'), context_spec=None,
               patch_size: Sequence[int] = (32, 32), step_size: Sequence[int] = (16, 16),
               batch_size: int = 1, pad: bool = True, input_mask_spec=None,
               fixed_mask_spec=None, invert_masks: bool = False, jax_device=None, **flow_args):
    super().__init__(context_spec)
    self._fixed_spec = fixed_spec
    self._image_dims = tuple(image_dims)
    self._patch_zyx = tuple(patch_size)[::-1]  # [z,]yx
    self._step_zyx = tuple(step_size)[::-1]  # [z,]yx
    self._batch_size = batch_size
    self._pad = pad
    self._input_mask_spec = input_mask_spec
    self._fixed_mask_spec = fixed_mask_spec
    self._invert_masks = invert_masks
    self._jax_device = jax_device
    self._flow_args = flow_args

  def decorate(self, input_ts):
    fixed_ts = volume.open_store(self._fixed_spec)
    in_dom = volume._domain_of(input_ts)

    def check(other, name):
      o = volume._domain_of(other)
      if in_dom.labels != o.labels:
        raise ValueError(f'Input TS and {name} must have same labels, but they are '
                         f'{in_dom.labels} and {o.labels}.')
      if tuple(input_ts.shape) != tuple(other.shape):
        raise ValueError(f'Input TS and {name} must have same shape, but they are '
                         f'{tuple(input_ts.shape)} and {tuple(other.shape)}.')

    check(fixed_ts, 'fixed TS')
    num_image_dims = len(self._image_dims)
    if num_image_dims not in (2, 3):
      raise ValueError(f'2 or 3 image dimensions are required, but got {num_image_dims}.')
    for d in self._image_dims:
      if d not in in_dom.labels:
        raise ValueError(f'image dimension {d} not among labels {in_dom.labels}.')
      elif in_dom[d].size < 2:
        raise ValueError(f'image dimension {d} must at least have size 2 but has size: '
                         f'{in_dom[d].size}.')
    input_mask_ts = fixed_mask_ts = None
    if self._input_mask_spec is not None:
      input_mask_ts = volume.open_store(self._input_mask_spec)
      check(input_mask_ts, 'input mask TS')
    if self._fixed_mask_spec is not None:
      fixed_mask_ts = volume.open_store(self._fixed_mask_spec)
      check(fixed_mask_ts, 'fixed mask TS')

    non_image_dims = [l for l in in_dom.labels if l not in self._image_dims]
    input_domain_dict = {dim.label: dim for dim in in_dom}

    def read_fn(domain, array, unused_read_params):
      domain_dict = {dim.label: dim for dim in domain}
      read_domain = volume.IndexDomain([domain_dict[l] for l in non_image_dims] +
                                       [input_domain_dict[l] for l in self._image_dims])

      def image(store, dtype):
        # indexing by a labelled domain keeps the store's own dimension order (image
        # dimensions x, y[, z] plus singleton non-image ones): squeeze, then xy[z] -> [z]yx
        return np.array(store[read_domain], dtype=dtype).squeeze().T

      pre_mask = post_mask = None
      if input_mask_ts is not None:
        pre_mask = image(input_mask_ts, bool)
        if self._invert_masks:
          pre_mask = ~pre_mask
      if fixed_mask_ts is not None:
        post_mask = image(fixed_mask_ts, bool)
        if self._invert_masks:
          post_mask = ~post_mask
      flow_post_to_pre = optim_flow(image(input_ts, np.float32), image(fixed_ts, np.float32),
                                    self._patch_zyx, self._step_zyx, self._batch_size,
                                    pre_mask=pre_mask, post_mask=post_mask, **self._flow_args)
      if num_image_dims == 2:
        flow_post_to_pre = np.asarray(flow_post_to_pre[:, np.newaxis, ...])
      if self._pad:
        pad_total = np.array(self._patch_zyx) // np.array(self._step_zyx) - 1
        pad_left = np.array(self._patch_zyx) // np.array(self._step_zyx) // 2
        pad_width = [(0, 0)]
        if num_image_dims == 2:
          pad_width.append([0, 0])
        for left, total in zip(pad_left, pad_total):
          pad_width.append([left, total - left])
        array[...] = np.pad(flow_post_to_pre, pad_width,
                            constant_values=np.nan).reshape(array.shape)
      else:
        array[...] = flow_post_to_pre.reshape(array.shape)

    labels = ['fc', 'fz', 'fy', 'fx'] + non_image_dims
    flow_shape = {'fc': num_image_dims + 2}
    if num_image_dims == 2:
      flow_shape['fz'] = 1
    calc_shape = _padded_flow_shape if self._pad else _flow_shape
    for i, l in enumerate(self._image_dims):
      flow_shape[labels[3 - i]] = int(calc_shape(o=input_domain_dict[l].size,
                                                 p=self._patch_zyx[-1 - i],
                                                 s=self._step_zyx[-1 - i]))
    dims, chunk = [], []
    for l in labels:
      if l in non_image_dims:
        d = input_domain_dict[l]
        dims.append(volume.Dim(d.inclusive_min, d.exclusive_max, l))
        chunk.append(1)
      else:
        dims.append(volume.Dim(0, flow_shape[l], l))
        chunk.append(flow_shape[l])
    return volume.VirtualChunked(read_fn, volume.IndexDomain(dims), chunk, np.float32)
