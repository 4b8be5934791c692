"""Drop-in for `sofima.flow_utils` (reference flow_utils.py): flow-field filters.

  apply_mask        flow_utils.py:31-33
  clean_flow        flow_utils.py:36-78
  reconcile_flows   flow_utils.py:81-135

These are small NumPy / SciPy filters in the reference (they sit between the hot-path calls,
SURVEY 8 f-4).  Two residencies, one result: NumPy arrays are filtered on the host exactly as
upstream does; CUDA tensors are filtered by the kernels of csrc/flowfilt.cu and stay on the
device, so that a section loop (EstimateMissingFlow, RelaxMesh) never leaves the GPU between
two hot-path calls.  Both are pinned on golden vectors from the reference module
(tests/golden/flow_utils_golden.npz; tests/test_flowfilt_gpu.py requires them to be equal).
"""

from __future__ import annotations

import ctypes
from typing import Sequence

import numpy as np
from scipy import ndimage


def _is_cuda(a) -> bool:
  return type(a).__module__.startswith('torch') and a.is_cuda


def _device_call(like):
  from . import _native
  ctx = _native.Context.get(like.device.index)
  ctx.bind_stream()
  return _native, ctx


def _clean_flow_device(flow, min_peak_ratio, min_peak_sharpness, max_magnitude, max_deviation,
                       dim):
  import torch
  native, ctx = _device_call(flow)
  f = flow.to(torch.float32).contiguous()
  if f.ndim != 4:
    raise ValueError('device clean_flow takes a [c, z, y, x] tensor')
  out = torch.empty((dim,) + tuple(f.shape[1:]), dtype=torch.float32, device=f.device)
  rc = native.lib().sofima_clean_flow(
      ctx.handle, f.data_ptr(), int(f.shape[0]), int(dim), (ctypes.c_int64 * 3)(*f.shape[1:]),
      float(min_peak_ratio), float(min_peak_sharpness), float(max_magnitude),
      float(max_deviation), out.data_ptr())
  native.check(ctx.handle, rc)
  return out


def _reconcile_flows_device(flows, max_gradient, max_deviation, min_patch_size, min_delta_z):
  import torch
  native, ctx = _device_call(flows[0])
  out = flows[0].to(torch.float32).contiguous().clone()
  if out.ndim != 4 or out.shape[0] not in (2, 3):
    raise ValueError('device reconcile_flows takes [2 or 3, z, y, x] tensors')
  others = [f.to(out.device, torch.float32).contiguous() for f in flows[1:]]
  ptrs = (ctypes.c_void_p * max(1, len(others)))(*[o.data_ptr() for o in others])
  rc = native.lib().sofima_reconcile_flows(
      ctx.handle, out.data_ptr(), ptrs, len(others), int(out.shape[0]),
      (ctypes.c_int64 * 3)(*out.shape[1:]), float(max_gradient), float(max_deviation),
      int(min_patch_size), float(min_delta_z))
  native.check(ctx.handle, rc)
  return out


def apply_mask(flow: np.ndarray, mask: np.ndarray):
  """Sets every channel of `flow` to NaN where `mask` is set (in place)."""
  for c in range(flow.shape[0]):
    flow[c, ...][mask] = np.nan


def _median3(field: np.ndarray, dim: int) -> np.ndarray:
  """Median over the 3x3[x3] spatial window of the NaN-zeroed field (SciPy's default
  'reflect' boundary, as the reference)."""
  size = (1, 1, 3, 3) if dim == 2 else (1, 3, 3, 3)
  return ndimage.median_filter(np.nan_to_num(field), size=size)


def clean_flow(flow: np.ndarray, min_peak_ratio: float, min_peak_sharpness: float,
               max_magnitude: float, max_deviation: float, dim: int = 2) -> np.ndarray:
  """Removes flow vectors that do not fulfil the quality requirements.

  Args:
    flow: [c, z, y, x] flow field with c = dim (vectors only) or dim + 2 (vectors,
      peak sharpness, peak ratio)
    min_peak_ratio: min. peak intensity ratio (a ratio of exactly 0 = "no second peak"
      always passes)
    min_peak_sharpness: min. |sharpness|
    max_magnitude: max. |component|; <= 0 disables the test
    max_deviation: max. |component - 3x3 median|; <= 0 disables the test
    dim: number of spatial dimensions of the flow vectors

  Returns:
    [dim, z, y, x] filtered flow (a copy), rejected vectors set to NaN
  """
  assert dim in (2, 3)
  assert dim <= flow.shape[0] <= dim + 2
  if _is_cuda(flow):
    return _clean_flow_device(flow, min_peak_ratio, min_peak_sharpness, max_magnitude,
                              max_deviation, dim)
  vec = flow[:dim, ...]
  with np.errstate(invalid='ignore'):
    if flow.shape[0] == dim + 2:
      out = vec.copy()
      ratio = np.abs(flow[dim + 1, ...])
      reject = (np.abs(flow[dim, ...]) < min_peak_sharpness) | (
          (ratio > 0.0) & (ratio < min_peak_ratio))
    else:
      out = flow.copy()
      reject = np.zeros(flow.shape[1:], dtype=bool)
    if max_magnitude > 0:
      reject |= np.abs(vec).max(axis=0) > max_magnitude
    if max_deviation > 0:
      reject |= np.abs(_median3(vec, dim) - vec).max(axis=0) > max_deviation
  apply_mask(out, reject)
  return out


def reconcile_flows(flows: Sequence[np.ndarray], max_gradient: float, max_deviation: float,
                    min_patch_size: int, min_delta_z: int = 0) -> np.ndarray:
  """Merges several flow estimates and filters the result.

  Args:
    flows: [c, z, y, x] flow arrays (c = 2 or 3) in order of decreasing preference
    max_gradient: max. |difference to the horizontal (x component) / vertical (y
      component) neighbour|, borders compared against 0; <= 0 disables the test
    max_deviation: max. |component - 3x3 median|; <= 0 disables the test
    min_patch_size: min. size of a connected valid region in pixels; <= 0 disables
    min_delta_z: for 3-channel flows, min. |z offset| for a fill-in value to count

  Returns:
    [c, z, y, x] reconciled flow
  """
  if _is_cuda(flows[0]):
    return _reconcile_flows_device(flows, max_gradient, max_deviation, min_patch_size,
                                   min_delta_z)
  out = flows[0].copy()
  nch = out.shape[0]
  assert nch in (2, 3)
  for other in flows[1:]:  # fill what is still invalid from the next estimate
    fill = np.repeat(np.isnan(out[0:1, ...]), nch, 0)
    if nch == 3:
      fill &= np.repeat(np.abs(other[2:3, ...]) >= min_delta_z, 3, 0)
    out[fill] = other[fill]

  with np.errstate(invalid='ignore'):
    if max_gradient > 0:
      steep = np.abs(np.diff(out[0, ...], axis=-1, prepend=0)) > max_gradient
      steep |= np.abs(np.diff(out[0, ...], axis=-1, append=0)) > max_gradient
      steep |= np.abs(np.diff(out[1, ...], axis=-2, prepend=0)) > max_gradient
      steep |= np.abs(np.diff(out[1, ...], axis=-2, append=0)) > max_gradient
      apply_mask(out, steep)
    if max_deviation > 0:
      med = ndimage.median_filter(np.nan_to_num(out), size=(1, 1, 3, 3))
      apply_mask(out, np.abs(med - out)[:2, ...].max(axis=0) > max_deviation)

  if min_patch_size > 0:
    valid = ~np.any(np.isnan(out), axis=0)
    tiny = np.zeros(valid.shape, dtype=bool)
    for z in range(valid.shape[0]):
      labels, _ = ndimage.label(valid[z, ...])
      ids, sizes = np.unique(labels, return_counts=True)
      tiny[z, ...] = np.isin(labels, ids[sizes < min_patch_size])
    apply_mask(out, tiny)
  return out
