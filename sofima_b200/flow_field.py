"""Drop-in for `sofima.flow_field` (reference flow_field.py) on the B200 backend.

Same public names and signatures as the reference:

  JAXMaskedXCorrWithStatsCalculator(.flow_field)   flow_field.py:449-712
  batched_xcorr_peaks                              flow_field.py:385-441
  masked_xcorr                                     flow_field.py:36-156
  _batched_peaks                                   flow_field.py:205-275

The host driver (`flow_field`) keeps the reference's index arithmetic -- output
geometry, patch selection from masks, 'edge'-padded final batch, targeting
fields -- and hands every batch to the CUDA library (`sofima_xcorr_peaks`).
Unlike the reference it does not block per batch: images are uploaded once, all
batches are queued on the stream and the peak table comes back in one copy.
There is no CPU path here.
"""

from __future__ import annotations

import collections
import collections.abc
import os
import ctypes
import logging
import threading
from typing import Callable, Iterator, Sequence, TypeVar

import numpy as np

from . import _native

T = TypeVar('T')


def _torch():
  import torch  # plumbing: device memory and streams
  return torch


def _is_tensor(a) -> bool:
  return type(a).__module__.startswith('torch')


class _Staging(threading.local):
  """Per-thread pinned arena for small host <-> device transfers.  Pinned allocations are
  expensive (cudaHostAlloc), and pageable copies synchronise the stream; slices of one
  grow-never arena cost nothing.  The arena is handed out front to back; when it is full
  the stream is synchronised once and it starts over."""

  CAP = 192 << 20

  def __init__(self):
    self.buf = None
    self.off = 0

  def take(self, nbytes: int):
    torch = _torch()
    if nbytes > self.CAP // 4:
      return None
    if self.buf is None:
      self.buf = torch.empty(self.CAP, dtype=torch.uint8, pin_memory=True)
    need = (nbytes + 255) & ~255
    if self.off + need > self.CAP:
      torch.cuda.synchronize()  # every transfer that used the arena has completed
      self.off = 0
    view = self.buf[self.off:self.off + nbytes]
    self.off += need
    return view


_STAGING = _Staging()


def _upload(a: np.ndarray, dev):
  """Host array -> device without synchronising the stream: the array is staged in pinned
  memory (one pass that also makes strided views contiguous) and copied asynchronously.  A
  plain `.to(device)` of pageable memory waits for all work queued on the stream first,
  which serialises back-to-back calls on short strips."""
  torch = _torch()
  src = torch.from_numpy(a) if a.flags.writeable else torch.from_numpy(a.copy())
  if src.is_contiguous() and src.is_pinned():
    return src.to(dev, non_blocking=True)
  raw = _STAGING.take(src.numel() * src.element_size())
  if raw is None:  # large image: the synchronisation is amortised
    return src.contiguous().to(dev)
  stage = raw.view(src.dtype).view(src.shape)
  stage.copy_(src)
  return stage.to(dev, non_blocking=True)


def _device_image(img, ctx):
  """uint8 or float32 contiguous CUDA tensor + SOFIMA dtype code."""
  torch = _torch()
  dev = torch.device('cuda', ctx.device)
  if _is_tensor(img):
    t = img.to(dev)
    if t.dtype != torch.uint8:
      t = t.to(torch.float32)
    return t.contiguous()
  a = np.asarray(img)
  if a.dtype != np.uint8:
    a = a.astype(np.float32)  # JAX computes in fp32 whatever the input dtype
  return _upload(a, dev)


def _device_mask(mask, ctx):
  if mask is None:
    return None
  torch = _torch()
  dev = torch.device('cuda', ctx.device)
  if _is_tensor(mask):
    return (mask.to(dev) != 0).to(torch.uint8).contiguous()
  m = np.ascontiguousarray(np.asarray(mask) != 0).view(np.uint8)
  return _upload(m, dev)


def _int3(vals, fill=0):
  vals = [int(v) for v in vals]
  return vals + [fill] * (3 - len(vals))


def _params(ndim, pre, post, pre_mask, post_mask, patch_size, post_patch_size,
            mean, min_distance, threshold_rel, peak_radius):
  torch = _torch()
  if pre.dtype != post.dtype:
    pre, post = pre.to(torch.float32), post.to(torch.float32)
  p = _native.XcorrParams()
  p.ndim = ndim
  p.img_dtype = 0 if pre.dtype == torch.uint8 else 1
  for i, v in enumerate(_int3(pre.shape)):
    p.pre_shape[i] = v
  for i, v in enumerate(_int3(post.shape)):
    p.post_shape[i] = v
  if pre_mask is not None:
    for i, v in enumerate(_int3(pre_mask.shape)):
      p.pre_mask_shape[i] = v
  if post_mask is not None:
    for i, v in enumerate(_int3(post_mask.shape)):
      p.post_mask_shape[i] = v
  for i, v in enumerate(_int3(patch_size, 1)):
    p.pre_patch[i] = v
  for i, v in enumerate(_int3(post_patch_size, 1)):
    p.post_patch[i] = v
  p.has_mean = int(mean is not None)
  p.mean = float(mean) if mean is not None else 0.0
  if isinstance(min_distance, collections.abc.Sequence):
    # The reference itself fails here (unbound `size`, flow_field.py:232-240).
    raise NotImplementedError('min_distance must be a scalar')
  p.min_distance = int(min_distance)
  p.threshold_rel = float(threshold_rel)
  if not isinstance(peak_radius, collections.abc.Sequence):
    peak_radius = (peak_radius,) * ndim
  for i, v in enumerate(_int3(peak_radius)):
    p.peak_radius[i] = v
  return p, pre, post


def _ptr(t):
  return None if t is None else t.data_ptr()


def batched_xcorr_peaks(pre_image, post_image, pre_mask, post_mask,
                        patch_size: Sequence[int], starts, mean: float | None,
                        min_distance: int = 2, threshold_rel: float = 0.5,
                        peak_radius: int | Sequence[int] = 5,
                        post_patch_size: Sequence[int] | None = None,
                        post_starts=None):
  """Computes cross-correlations and identifies their peaks (flow_field.py:385).

  Args and result as in the reference: `starts` / `post_starts` are [b, 2 or 3]
  integer top-left ([z]yx) patch coordinates; returns a [b, ndim + 2] float32
  array (x, y[, z], sharpness, peak ratio).  NumPy inputs give a NumPy result,
  CUDA tensors a CUDA tensor.
  """
  ctx = _native.Context.get(
      pre_image.device.index if _is_tensor(pre_image) and pre_image.is_cuda else None)
  torch = _torch()
  host_result = not _is_tensor(pre_image)
  pre = _device_image(pre_image, ctx)
  post = _device_image(post_image, ctx)
  pre_m = _device_mask(pre_mask, ctx)
  post_m = _device_mask(post_mask, ctx)
  ndim = pre.ndim
  if post_patch_size is None:
    post_patch_size = patch_size
  if post_starts is None:
    post_starts = starts
  p, pre, post = _params(ndim, pre, post, pre_m, post_m, patch_size,
                         post_patch_size, mean, min_distance, threshold_rel,
                         peak_radius)
  dev = torch.device('cuda', ctx.device)

  def starts_dev(s):
    if _is_tensor(s):
      return s.to(dev, dtype=torch.int32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(s, dtype=np.int32)).to(dev)

  st, pst = starts_dev(starts), starts_dev(post_starts)
  b = st.shape[0]
  out = torch.empty((b, ndim + 2), dtype=torch.float32, device=dev)
  ctx.bind_stream()
  rc = _native.lib().sofima_xcorr_peaks(
      ctx.handle, ctypes.byref(p), pre.data_ptr(), post.data_ptr(), _ptr(pre_m),
      _ptr(post_m), st.data_ptr(), pst.data_ptr(), b, out.data_ptr())
  _native.check(ctx.handle, rc)
  return out.cpu().numpy() if host_result else out


def masked_xcorr(prev, curr, prev_mask=None, curr_mask=None, use_jax: bool = False,
                 dim: int = 2):
  """Cross-correlation between two (batches of) masked images (flow_field.py:36).

  Correlation is computed over the last `dim` axes (2 or 3); leading axes are batch.
  The inputs are whole patches (no mean subtraction is applied here, as in the
  reference); `prev` and `curr` may differ in size (the coarse 3-d offset search of
  notebooks/liconn_inplane_stitching.ipynb correlates a small query cuboid with a
  large search volume).  `use_jax` is accepted for signature compatibility.
  """
  del use_jax
  prev = np.asarray(prev, dtype=np.float32)
  curr = np.asarray(curr, dtype=np.float32)
  if dim not in (2, 3):
    raise NotImplementedError(f'correlation over {dim} axes: only 2 and 3 are built')
  lead = prev.shape[:-dim]
  pb = prev.reshape((-1,) + prev.shape[-dim:])
  cb = curr.reshape((-1,) + curr.shape[-dim:])
  nb = pb.shape[0]
  assert cb.shape[0] == nb
  pm = None if prev_mask is None else np.asarray(prev_mask, bool).reshape(pb.shape)
  cm = None if curr_mask is None else np.asarray(curr_mask, bool).reshape(cb.shape)
  ctx = _native.Context.get()
  torch = _torch()
  dev = torch.device('cuda', ctx.device)
  # Lay the batch out as one tall image / volume so that patch b starts at b * size[0]
  # along the first spatial axis.
  tall = lambda a: np.ascontiguousarray(a.reshape((-1,) + a.shape[2:]))
  pre = torch.from_numpy(tall(pb)).to(dev)
  post = torch.from_numpy(tall(cb)).to(dev)
  pre_m = None if pm is None else _device_mask(tall(pm), ctx)
  post_m = None if cm is None else _device_mask(tall(cm), ctx)
  p, pre, post = _params(dim, pre, post, pre_m, post_m, pb.shape[1:], cb.shape[1:],
                         0.0, 2, 0.5, 0)
  st = torch.zeros((nb, dim), dtype=torch.int32, device=dev)
  st[:, 0] = torch.arange(nb, device=dev, dtype=torch.int32) * pb.shape[1]
  pst = torch.zeros((nb, dim), dtype=torch.int32, device=dev)
  pst[:, 0] = torch.arange(nb, device=dev, dtype=torch.int32) * cb.shape[1]
  out_shape = tuple(int(a) + int(b) - 1 for a, b in zip(pb.shape[1:], cb.shape[1:]))
  out = torch.empty((nb,) + out_shape, dtype=torch.float32, device=dev)
  ctx.bind_stream()
  rc = _native.lib().sofima_xcorr_images(
      ctx.handle, ctypes.byref(p), pre.data_ptr(), post.data_ptr(), _ptr(pre_m),
      _ptr(post_m), st.data_ptr(), pst.data_ptr(), nb, out.data_ptr())
  _native.check(ctx.handle, rc)
  return out.cpu().numpy().reshape(lead + out_shape)


def _batched_peaks(img, center_offset, min_distance, threshold_rel,
                   peak_radius: int | Sequence[int] = 5):
  """Peak statistics from a batch of correlation images (flow_field.py:205).

  Args:
    img: [b, y, x] correlation images
    center_offset: (y, x) location of the zero-shift peak
    min_distance: min. distance in pixels between peaks (scalar)
    threshold_rel: fraction of the image max that a peak has to exceed
    peak_radius: radius for the sharpness window

  Returns:
    [b, 4] array: x, y peak offset from center, sharpness, peak ratio
  """
  host_result = not _is_tensor(img)
  ctx = _native.Context.get(
      img.device.index if _is_tensor(img) and img.is_cuda else None)
  torch = _torch()
  dev = torch.device('cuda', ctx.device)
  if _is_tensor(img):
    t = img.to(dev, dtype=torch.float32).contiguous()
  else:
    t = torch.from_numpy(np.ascontiguousarray(img, dtype=np.float32)).to(dev)
  ndim = t.ndim - 1
  if isinstance(min_distance, collections.abc.Sequence):
    raise NotImplementedError('min_distance must be a scalar')
  if not isinstance(peak_radius, collections.abc.Sequence):
    peak_radius = (peak_radius,) * ndim
  shape = (ctypes.c_int64 * 3)(*_int3(t.shape[1:]))
  center = (ctypes.c_int32 * 3)(*_int3(center_offset))
  radius = (ctypes.c_int32 * 3)(*_int3(peak_radius))
  out = torch.empty((t.shape[0], ndim + 2), dtype=torch.float32, device=dev)
  ctx.bind_stream()
  rc = _native.lib().sofima_batched_peaks(
      ctx.handle, ndim, t.data_ptr(), shape, t.shape[0], center, int(min_distance),
      float(threshold_rel), radius, out.data_ptr())
  _native.check(ctx.handle, rc)
  return out.cpu().numpy() if host_result else out


# --- summed-area helpers (connectomics.common.geom_utils stand-ins) ----------------


def _integral_image(mask):
  """Zero-front-padded summed-area table of a mask (flow_field.py:159-175)."""
  if mask is None:
    return None
  m = np.asarray(mask)
  ii = m.astype(np.uint32 if m.size < 2**32 else np.int64)
  for axis in range(m.ndim):
    ii = ii.cumsum(axis=axis, dtype=ii.dtype)
  return np.pad(ii, [(1, 0)] * m.ndim, mode='constant')


def _query_integral_image(summed, diam, stride):
  """Sums over `diam`-sized boxes at every `stride` (VALID mode)."""
  summed = np.asarray(summed).astype(np.int64)
  nd = summed.ndim
  counts = [(n - 1 - d) // s + 1 for n, d, s in zip(summed.shape, diam, stride)]
  out = np.zeros(counts, np.int64)
  for corner in np.ndindex(*([2] * nd)):
    sel = tuple(
        slice(diam[a] * corner[a],
              diam[a] * corner[a] + (counts[a] - 1) * stride[a] + 1, stride[a])
        for a in range(nd))
    out += (-1) ** (nd - sum(corner)) * summed[sel]
  return out


def _batches(seq, n):
  for i in range(0, len(seq), n):
    yield seq[i:i + n]


def _silent_fn(x: list[T]) -> Iterator[T]:
  for item in x:
    yield item


class JAXMaskedXCorrWithStatsCalculator:
  """Estimates optical flow using masked cross-correlation.

  Name kept from the reference (flow_field.py:449) so that callers such as
  processor.flow.EstimateFlow work unchanged; the computation runs in CUDA.
  """

  non_spatial_flow_channels = 2  # peak sharpness, peak ratio
  supports_async = True          # flow_field(..., _async=True) returns a handle

  def __init__(self, mean: float | None = None, peak_min_distance: float = 2,
               peak_radius: float = 5):
    self._mean = mean
    self._min_distance = peak_min_distance
    self._peak_radius = peak_radius

  def flow_field(self, pre_image, post_image, patch_size, step, pre_mask=None,
                 post_mask=None, mask_only_for_patch_selection=False,
                 selection_mask=None, max_masked=0.75, batch_size=4096,
                 post_patch_size=None, pre_targeting_field=None,
                 pre_targeting_step=None, post_targeting_field=None,
                 post_targeting_step=None,
                 progress_fn: Callable[[list[T]], Iterator[T]] = _silent_fn,
                 _async: bool = False):
    """Computes the flow field from post to pre (flow_field.py:474-712).

    Arguments and result are those of the reference.  Returns a float32 array
    [ndim + 2, *out_shape]; channel order x, y[, z], sharpness, peak ratio; NaN
    where no flow was computed.  (`_async=True`, used by stitch_elastic.compute_flow_map:
    returns a handle whose `.result()` gives that array, so that many short calls can be
    queued on the GPU before the first result is read back.)
    """
    assert pre_image.ndim == post_image.ndim
    nd = pre_image.ndim

    if not isinstance(patch_size, collections.abc.Sequence):
      patch_size = (patch_size,) * nd
    if post_patch_size is not None:
      if not isinstance(post_patch_size, collections.abc.Sequence):
        post_patch_size = (post_patch_size,) * nd
    else:
      post_patch_size = patch_size
    if not isinstance(step, collections.abc.Sequence):
      step = (step,) * nd
    if pre_targeting_step is not None and not isinstance(
        pre_targeting_step, collections.abc.Sequence):
      pre_targeting_step = (pre_targeting_step,) * nd

    assert len(patch_size) == nd
    assert len(post_patch_size) == nd
    assert len(step) == nd

    out_shape = (np.array(post_image.shape)
                 - (np.array(post_patch_size) - step)) // step
    out_sel = tuple(slice(0, int(s)) for s in out_shape)
    output = np.full([self.non_spatial_flow_channels + nd] + out_shape.tolist(),
                     np.nan, dtype=np.float32)

    if selection_mask is None:
      selection_mask = np.ones(out_shape, dtype=bool)
    else:
      selection_mask = np.array(selection_mask[out_sel], dtype=bool)

    for mask, psz in ((pre_mask, patch_size), (post_mask, post_patch_size)):
      if mask is not None:
        s = _query_integral_image(_integral_image(np.asarray(mask)), psz, step)
        selection_mask[(s / np.prod(psz) >= max_masked)[out_sel]] = False

    if mask_only_for_patch_selection:
      pre_mask = post_mask = None

    oyx = np.array(np.where(selection_mask)).T
    logging.info('Starting flow estimation for %d patches.', oyx.shape[0])
    if oyx.shape[0] == 0:
      return _PendingFlow(None, None, output, ()) if _async else output

    ctx = _native.Context.get()
    pre_d = _device_image(pre_image, ctx)
    post_d = _device_image(post_image, ctx)
    pre_m = _device_mask(pre_mask, ctx)
    post_m = _device_mask(post_mask, ctx)
    job = _cached_job(ctx, oyx, pre_image.shape, post_image.shape, patch_size,
                      post_patch_size, step, batch_size, pre_targeting_field,
                      pre_targeting_step, post_targeting_field, post_targeting_step)
    peaks_d = job.run(pre_d, post_d, pre_m, post_m, self._mean, self._min_distance,
                      self._peak_radius, progress_fn)
    pending = _PendingFlow(job, peaks_d, output, (pre_d, post_d, pre_m, post_m))
    return pending if _async else pending.result()


class _PendingFlow:
  """Result of a queued flow_field call: the peak table is copied to pinned host memory
  asynchronously; `result()` waits for it and scatters it into the flow field."""

  def __init__(self, job, peaks_d, output, keep_alive):
    self._job, self._output, self._keep = job, output, keep_alive
    self._host = self._event = None
    if job is not None:
      torch = _torch()
      raw = _STAGING.take(peaks_d.numel() * peaks_d.element_size())
      if raw is None:
        raw = torch.empty(peaks_d.numel() * peaks_d.element_size(), dtype=torch.uint8,
                          pin_memory=True)
      self._host = raw.view(peaks_d.dtype).view(peaks_d.shape)
      self._host.copy_(peaks_d, non_blocking=True)
      self._event = torch.cuda.Event()
      self._event.record()

  def result(self) -> np.ndarray:
    if self._job is not None:
      self._event.synchronize()
      self._job.scatter(self._host.numpy(), self._output)
      self._job = self._keep = None
      logging.info('Flow field estimation complete.')
    return self._output


class _FlowJob:
  """Host-side index tables of one flow_field call, uploaded once.

  Holds the per-batch patch start coordinates (flow_field.py:610-680: fixed batch
  size with 'edge' padding, pre/post targeting offsets, clipping) on the device.
  `run` queues one `sofima_xcorr_peaks` call per reference batch -- the
  second-peak rule couples the members of a batch, so the batch composition is
  part of the result -- without any host synchronisation in between.
  """

  def __init__(self, ctx, oyx, pre_shape, post_shape, patch_size, post_patch_size,
               step, batch_size, pre_targeting_field=None, pre_targeting_step=None,
               post_targeting_field=None, post_targeting_step=None):
    torch = _torch()
    self.ctx = ctx
    self.nd = nd = len(pre_shape)
    self.patch_size, self.post_patch_size = tuple(patch_size), tuple(post_patch_size)
    self.batch_size = int(batch_size)
    patch_offset = ((np.array(patch_size) - post_patch_size) // 2)[None, ...]
    patch_offset = patch_offset.astype(int)
    step_arr = np.array(step).reshape((1, -1))

    # Within a reference batch the ORDER of the patch pairs is free (the batch-coupled
    # second-peak rule, flow_field.py:263-265, depends on the set only), so every batch is
    # walked column-major: consecutive pairs then share 3/4 of their image rows at the same
    # x start, which is what the fused kernel's row-spectra reads want to find in L2.
    # `scatter` writes by position, so the flow field is unaffected.
    self.batches = []
    for pos in _batches(oyx, self.batch_size):
      if nd == 2 and pos.shape[0] > 1:
        pos = pos[np.lexsort((pos[:, 0], pos[:, 1]))]
      self.batches.append(pos)
    pre_all, post_all, self.tg_all, self.po_all = [], [], [], []
    for pos_zyx in self.batches:
      real = pos_zyx.shape[0]
      if real < self.batch_size:  # 'edge' padding, flow_field.py:614-618
        proc = np.pad(pos_zyx, ((0, self.batch_size - real), (0, 0)), mode='edge')
      else:
        proc = pos_zyx
      post_starts = proc * step_arr
      pre_starts = np.clip(post_starts - patch_offset, 0, np.inf).astype(int)
      tg = po = None
      if pre_targeting_field is not None and pre_targeting_step is not None:
        tg = _targeting_offsets(pre_targeting_field, pre_targeting_step,
                                pre_starts, patch_size, pre_shape)
        pre_starts = pre_starts + tg
      if post_targeting_field is not None and post_targeting_step is not None:
        po = _targeting_offsets(post_targeting_field, post_targeting_step,
                                post_starts, post_patch_size, post_shape)
        post_starts = post_starts + po
      pre_all.append(np.clip(pre_starts, 0, np.inf).astype(np.int32))
      post_all.append(np.clip(post_starts, 0, np.inf).astype(np.int32))
      self.tg_all.append(tg)
      self.po_all.append(po)

    # NB: np.stack keeps the (Fortran) layout of np.where-derived views, and the
    # library takes plain C-contiguous [batch, nd] tables.
    starts_h = np.ascontiguousarray(
        np.stack([np.stack(pre_all), np.stack(post_all)]), dtype=np.int32)
    dev = torch.device('cuda', ctx.device)
    self.starts_d = _upload(starts_h, dev)  # [2, nb, B, nd]
    self.num_pairs = int(oyx.shape[0])
    # Distinct patch x starts (clamped like the kernels' dynamic_slice) per image, for
    # the shared row spectra (sofima_xcorr_rowcache).
    self._xstarts = None
    if nd == 2:
      xs = []
      for i, (shape, psz) in enumerate(((pre_shape, patch_size), (post_shape, post_patch_size))):
        x = np.clip(starts_h[i, :, :, 1], 0, max(int(shape[1]) - int(psz[1]), 0))
        xs.append(np.unique(x).astype(np.int32))
      self._xstarts = xs
      rows_direct = starts_h.shape[1] * starts_h.shape[2] * (patch_size[0] + post_patch_size[0])
      rows_shared = len(xs[0]) * int(pre_shape[0]) + len(xs[1]) * int(post_shape[0])
      spec_bytes = rows_shared * (sum(patch_size[1:]) + sum(post_patch_size[1:])) * 4 + 1
      self._share_rows = (rows_shared < 0.6 * rows_direct and spec_bytes < (4 << 30))

  def run(self, pre_d, post_d, pre_m=None, post_m=None, mean=None, min_distance=2,
          peak_radius=5, progress_fn=_silent_fn, out=None):
    """Queues every batch on the current stream; returns peaks [nb, B, nd + 2]."""
    torch = _torch()
    nb = len(self.batches)
    params, pre_d, post_d = _params(
        self.nd, pre_d, post_d, pre_m, post_m, self.patch_size, self.post_patch_size,
        mean, min_distance, 0.5, peak_radius)
    if out is None:
      out = torch.empty((nb, self.batch_size, self.nd + 2), dtype=torch.float32,
                        device=pre_d.device)
    self.ctx.bind_stream()
    lib = _native.lib()
    cached = False
    if (self._xstarts is not None and self._share_rows and pre_m is None and post_m is None
        and os.environ.get('SOFIMA_FLOW_ROWCACHE', '1') != '0'):
      # Overlapping patches share their forward row transforms (flow.cu).
      xa, xb = self._xstarts
      i32p = ctypes.POINTER(ctypes.c_int32)
      rc = lib.sofima_xcorr_rowcache(
          self.ctx.handle, ctypes.byref(params), pre_d.data_ptr(), post_d.data_ptr(),
          xa.ctypes.data_as(i32p), len(xa), xb.ctypes.data_as(i32p), len(xb))
      cached = rc == _native.OK  # unsupported size / no memory: plain path
    try:
      for i in progress_fn(list(range(nb))):
        # The 'edge' padding of the last batch (flow_field.py:614-618) repeats its final
        # position: the copies have the same first peak as the original, so they add
        # nothing to the batch's erase set (flow_field.py:263-265) and their rows are
        # discarded -- only the real pairs are computed.  The masked path normalises with
        # batch-wide maxima (flow_field.py:137, :151), which duplicates cannot change either.
        real = self.batches[i].shape[0]
        rc = lib.sofima_xcorr_peaks(
            self.ctx.handle, ctypes.byref(params), pre_d.data_ptr(), post_d.data_ptr(),
            _ptr(pre_m), _ptr(post_m), self.starts_d[0, i].data_ptr(),
            self.starts_d[1, i].data_ptr(), real, out[i].data_ptr())
        _native.check(self.ctx.handle, rc)
    finally:
      if cached:
        lib.sofima_xcorr_rowcache(self.ctx.handle, None, None, None, None, 0, None, 0)
    return out

  def scatter(self, peaks: np.ndarray, output: np.ndarray):
    """Adds the targeting offsets back and writes rows into the flow field
    (flow_field.py:699-709)."""
    nd = self.nd
    for i, pos_zyx in enumerate(self.batches):
      real = pos_zyx.shape[0]
      v = peaks[i, :real]
      if self.tg_all[i] is not None:
        v[:, :nd] = v[:, :nd] + self.tg_all[i][:real, ::-1]  # xy[z]
      if self.po_all[i] is not None:
        v[:, :nd] = v[:, :nd] - self.po_all[i][:real, ::-1]  # xy[z]
      output[(slice(None),) + tuple(pos_zyx.T)] = v.T


# Index tables of recent calls.  A pipeline calls flow_field with the same geometry for
# every section / tile pair (processor/flow.py:163-251), and building + uploading the
# tables costs about as much host time as the H2D copy of the images.  Calls with
# targeting fields (data-dependent starts) are not cached.
_JOB_CACHE: 'collections.OrderedDict[tuple, _FlowJob]' = collections.OrderedDict()
_JOB_CACHE_LOCK = threading.Lock()
_JOB_CACHE_SIZE = 8


def _cached_job(ctx, oyx, pre_shape, post_shape, patch_size, post_patch_size, step,
                batch_size, pre_tf, pre_ts, post_tf, post_ts) -> _FlowJob:
  if pre_tf is not None or post_tf is not None:
    return _FlowJob(ctx, oyx, pre_shape, post_shape, patch_size, post_patch_size, step,
                    batch_size, pre_tf, pre_ts, post_tf, post_ts)
  key = (ctx.device, threading.get_ident(), tuple(pre_shape), tuple(post_shape), tuple(patch_size),
         tuple(post_patch_size), tuple(step), int(batch_size), oyx.shape,
         hash(np.ascontiguousarray(oyx).tobytes()))
  with _JOB_CACHE_LOCK:
    job = _JOB_CACHE.get(key)
    if job is not None and np.array_equal(job.oyx, oyx):
      _JOB_CACHE.move_to_end(key)
      return job
  job = _FlowJob(ctx, oyx, pre_shape, post_shape, patch_size, post_patch_size, step,
                 batch_size)
  job.oyx = oyx.copy()
  with _JOB_CACHE_LOCK:
    _JOB_CACHE[key] = job
    while len(_JOB_CACHE) > _JOB_CACHE_SIZE:
      _JOB_CACHE.popitem(last=False)
  return job


def _targeting_offsets(field, tg_step, starts, patch, img_shape):
  """Start offsets from a targeting field (flow_field.py:626-649, :652-677)."""
  center = (np.array(patch) // 2).reshape((1, -1))  # [z]yx
  tg_step = np.array(tg_step).reshape((1, -1))
  query = np.round((starts + center) / tg_step).astype(int)  # [b, [z]yx]
  q = [np.clip(query[:, i], 0, field.shape[i + 1] - 1)
       for i in range(query.shape[-1])]
  off = np.nan_to_num(field[(slice(None),) + tuple(q)].T)
  off = off.astype(int)[:, ::-1]  # [b, xy[z]] -> [b, [z]yx]
  new_starts = starts + off
  # Clip offsets that would take the patch out of bounds.
  off = off - np.minimum(new_starts, 0)
  shape = np.array(img_shape)[None, ...]
  new_ends = new_starts + np.array(patch)[None, ...]
  return off - (np.maximum(new_ends, shape) - shape)
