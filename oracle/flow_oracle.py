"""NumPy fp32 restatement of the reference patch-flow estimator (TEST ORACLE).

Follows /root/reference/flow_field.py:
  * masked_xcorr                 flow_field.py:36-156  (use_jax=True branch)
  * _batched_xcorr               flow_field.py:278-371
  * _batched_peaks / _peak_stats flow_field.py:205-275 / 178-202
  * batched_xcorr_peaks          flow_field.py:385-441
  * JAXMaskedXCorrWithStatsCalculator.flow_field   flow_field.py:474-712
and the three helpers of the un-vendored `connectomics` package used there
(`geom_utils.integral_image`, `geom_utils.query_integral_image`, `utils.batch`;
flow_field.py:167,576,584,610 -- package unpinned in setup.cfg:23; semantics:
zero-front-padded summed-area table, VALID box sums of `diam` at `stride`,
consecutive chunks of <= n).

JAX primitives are restated from their documented semantics:
  jax.lax.dynamic_slice            start indices clamped so the slice fits
  conv_general_dilated_patches     'same' padding = zero padding
  jnp.argmax                       first (lowest flat index) maximum
  x.at[:, idx].set(v)              NumPy-style column assignment
FFTs are fp32 pocketfft (scipy.fft), the same algorithm family JAX-CPU uses.

Test infrastructure only: never imported from `sofima_b200/`.
"""

from __future__ import annotations

import collections.abc
import os
from typing import Sequence

import numpy as np
import scipy.fft

F32 = np.float32
# pocketfft worker threads (the CPU baseline uses every host core).
WORKERS = int(os.environ.get('SOFIMA_ORACLE_FFT_WORKERS', os.cpu_count() or 1))


def next_fast_len(n: int) -> int:
  """Smallest 5-smooth integer >= n (scipy.fftpack.next_fast_len)."""
  n = int(n)
  while True:
    m = n
    for p in (2, 3, 5):
      while m % p == 0:
        m //= p
    if m == 1:
      return n
    n += 1


def masked_xcorr(prev, curr, prev_mask=None, curr_mask=None, dim=2):
  """flow_field.py:36-156, batch in the leading axes, fp32."""
  prev = np.asarray(prev, dtype=F32)
  curr = np.asarray(curr, dtype=F32)
  shape = np.array(prev.shape[-dim:]) + np.array(curr.shape[-dim:]) - 1
  fast_shape = [next_fast_len(int(s)) for s in shape]
  crop = (Ellipsis,) + tuple(slice(0, int(s)) for s in shape)
  axes = tuple(range(-dim, 0))
  flip = (Ellipsis,) + (slice(None, None, -1),) * dim

  if prev_mask is not None:
    prev = np.where(prev_mask, F32(0), prev)
  if curr_mask is not None:
    curr = np.where(curr_mask, F32(0), curr)
  curr = curr[flip]

  def fwd(a):
    return scipy.fft.rfftn(np.asarray(a, dtype=F32), s=fast_shape, axes=axes,
                           workers=WORKERS)

  def inv(a):
    return scipy.fft.irfftn(a, s=fast_shape, axes=axes,
                            workers=WORKERS).astype(F32)

  p_f = fwd(prev)
  c_f = fwd(curr)
  xcorr = inv(p_f * c_f)
  if prev_mask is None and curr_mask is None:
    return xcorr[crop]                                   # flow_field.py:88-89

  pm = (np.ones(prev.shape, bool) if prev_mask is None
        else np.logical_not(prev_mask))
  cm = (np.ones(curr.shape, bool) if curr_mask is None
        else np.logical_not(curr_mask))
  cm = cm[flip]
  pm_f = fwd(pm)
  cm_f = fwd(cm)

  eps = np.finfo(F32).eps
  overlap = np.fmax(np.round(inv(cm_f * pm_f)), eps)      # :114-115
  overlap_inv = F32(1.0) / overlap
  mc_p = inv(cm_f * p_f)
  mc_c = inv(pm_f * c_f)
  xcorr = xcorr - mc_p * mc_c * overlap_inv              # :121
  p_den = np.fmax(inv(cm_f * fwd(np.square(prev)))
                  - np.square(mc_p) * overlap_inv, F32(0))
  c_den = np.fmax(inv(pm_f * fwd(np.square(curr)))
                  - np.square(mc_c) * overlap_inv, F32(0))
  denom = np.sqrt(p_den * c_den)

  xcorr, denom, overlap = xcorr[crop], denom[crop], overlap[crop]
  tol = F32(1e3) * eps * np.max(np.abs(denom))            # global max, :137
  with np.errstate(all='ignore'):
    out = np.where(denom > tol, xcorr / denom, F32(0))
  out = np.clip(out, -1, 1)
  px_threshold = F32(0.3) * np.max(overlap)               # global max, :151
  return np.where(overlap < px_threshold, F32(0), out).astype(F32)


def _dynamic_slice(img, start, size):
  """jax.lax.dynamic_slice: starts are clamped so that the slice is in-bounds."""
  sel = []
  for st, sz, n in zip(start, size, img.shape):
    st = int(min(max(int(st), 0), n - int(sz)))
    sel.append(slice(st, st + int(sz)))
  return img[tuple(sel)]


def batched_xcorr(pre_image, post_image, pre_mask, post_mask, patch_size,
                  starts, mean, post_patch_size=None, post_starts=None):
  """flow_field.py:278-371.  Returns (center_offset, xcorr [b, ...])."""
  if post_patch_size is None:
    post_patch_size = patch_size
  if post_starts is None:
    post_starts = starts
  dim = len(patch_size)
  axes = tuple(range(-dim, 0))

  def gather(img, sts, size, dtype):
    return np.stack([_dynamic_slice(img, s, size) for s in sts]).astype(dtype)

  pre_b = gather(np.asarray(pre_image), starts, patch_size, F32)
  post_b = gather(np.asarray(post_image), post_starts, post_patch_size, F32)
  pre_m = (None if pre_mask is None
           else gather(np.asarray(pre_mask), starts, patch_size, bool))
  post_m = (None if post_mask is None
            else gather(np.asarray(post_mask), post_starts, post_patch_size,
                        bool))

  def patch_mean(src, m):
    # jnp.mean / jnp.nanmean = fp32 sum / fp32 count.  The sum is accumulated
    # in fp64 here (exact for integer-valued pixels, where any fp32 summation
    # order is exact too), the division is the reference's fp32 division.
    if m is None:
      tot = np.sum(src, axis=axes, keepdims=True, dtype=np.float64)
      return tot.astype(F32) / F32(np.prod(src.shape[-dim:]))
    with np.errstate(all='ignore'):
      cnt = np.sum(~m, axis=axes, keepdims=True)
      tot = np.sum(np.where(m, 0.0, src), axis=axes, keepdims=True,
                   dtype=np.float64)
      return tot.astype(F32) / cnt.astype(F32)  # all-masked patch -> nan

  if mean is None:
    pre_mean, post_mean = patch_mean(pre_b, pre_m), patch_mean(post_b, post_m)
  else:
    pre_mean = post_mean = F32(mean)

  center = (np.array(pre_b.shape[-dim:]) + np.array(post_b.shape[-dim:])) // 2 - 1
  return center, masked_xcorr(pre_b - pre_mean, post_b - post_mean, pre_m,
                              post_m, dim=dim)


def _max_filter_zero_pad(img, axis, width):
  """Running max of odd `width` along `axis`, zero padding (SAME conv patches)."""
  r = width // 2
  pad = [(0, 0)] * img.ndim
  pad[axis] = (r, r)
  p = np.pad(img, pad, mode='constant', constant_values=0)
  out = None
  n = img.shape[axis]
  for o in range(width):
    sl = [slice(None)] * img.ndim
    sl[axis] = slice(o, o + n)
    cur = p[tuple(sl)]
    out = cur if out is None else np.maximum(out, cur)
  return out


def batched_peaks(img, center_offset, min_distance=2, threshold_rel=0.5,
                  peak_radius=5):
  """flow_field.py:205-275 (+ _peak_stats :178-202).  img: [b, [z,] y, x]."""
  img = np.asarray(img, dtype=F32)
  dim = img.ndim - 1
  b = img.shape[0]
  if isinstance(min_distance, collections.abc.Sequence):
    raise NotImplementedError(
        'sequence min_distance hits an unbound variable in the reference '
        '(flow_field.py:232-240)')
  width = 2 * int(min_distance) + 1

  img_max = img
  for ax in range(1, dim + 1):
    img_max = _max_filter_zero_pad(img_max, ax, width)

  thresholds = F32(threshold_rel) * img.max(axis=tuple(range(1, dim + 1)),
                                            keepdims=True)
  peak_mask = (img == img_max) & (img > thresholds)
  flat = np.where(peak_mask, img, -np.inf).astype(F32).reshape(b, -1)

  p1 = np.argmax(flat, axis=-1)
  v1 = flat[np.arange(b), p1]
  flat2 = flat.copy()
  flat2[:, p1] = -np.inf          # every row loses every batch member's peak
  p2 = np.argmax(flat2, axis=-1)
  v2 = flat[np.arange(b), p2]     # NOTE: read from the un-erased array (:266)

  if not isinstance(peak_radius, collections.abc.Sequence):
    peak_radius = (peak_radius,) * dim
  size = 2 * np.array(peak_radius) + 1

  out = np.zeros((b, dim + 2), dtype=F32)
  for i in range(b):
    inds = np.unravel_index(p1[i], img.shape[1:])
    if np.isinf(v1[i]):
      out[i] = np.nan
      continue
    window = _dynamic_slice(img[i], np.array(inds) - size // 2, size)
    with np.errstate(all='ignore'):
      sharp = img[i][inds] / np.min(window)
      ratio = F32(0.0) if np.isinf(v2[i]) else v1[i] / v2[i]
    centered = [F32(x) - F32(o) for x, o in zip(inds, center_offset)]
    out[i, :dim] = centered[::-1]
    out[i, dim] = sharp
    out[i, dim + 1] = ratio
  return out


def batched_xcorr_peaks(pre_image, post_image, pre_mask, post_mask, patch_size,
                        starts, mean, min_distance=2, threshold_rel=0.5,
                        peak_radius=5, post_patch_size=None, post_starts=None):
  """flow_field.py:385-441."""
  center, xc = batched_xcorr(pre_image, post_image, pre_mask, post_mask,
                             patch_size, starts, mean, post_patch_size,
                             post_starts)
  return batched_peaks(xc, center, min_distance, threshold_rel, peak_radius)


# --- connectomics stand-ins ---------------------------------------------------


def integral_image(mask):
  ii = np.asarray(mask).astype(np.uint32 if np.asarray(mask).size < 2**32
                               else np.int64)
  for ax in range(ii.ndim):
    ii = ii.cumsum(axis=ax, dtype=ii.dtype)
  return np.pad(ii, [(1, 0)] * ii.ndim, mode='constant')


def query_integral_image(summed, diam, stride):
  """Box sums of size `diam` at every `stride` (VALID), any rank."""
  summed = np.asarray(summed).astype(np.int64)
  nd = summed.ndim
  out = np.zeros([(n - 1 - d) // s + 1
                  for n, d, s in zip(summed.shape, diam, stride)], np.int64)
  for corner in np.ndindex(*([2] * nd)):
    sel = []
    for ax in range(nd):
      off = diam[ax] if corner[ax] else 0
      cnt = out.shape[ax]
      sel.append(slice(off, off + (cnt - 1) * stride[ax] + 1, stride[ax]))
    sign = (-1) ** (nd - sum(corner))
    out += sign * summed[tuple(sel)]
  return out


def batch(seq, n):
  for i in range(0, len(seq), n):
    yield seq[i:i + n]


# --- host driver ----------------------------------------------------------------


def _as_tuple(v, nd):
  return tuple(v) if isinstance(v, collections.abc.Sequence) else (v,) * nd


def _targeting_offsets(field, tg_step, starts, patch, img_shape):
  """flow_field.py:626-649 / :652-677 (same arithmetic for pre and post)."""
  centre = (np.array(patch) // 2).reshape((1, -1))
  query = np.round((starts + centre) / np.array(tg_step).reshape((1, -1)))
  query = query.astype(int)
  q = [np.clip(query[:, i], 0, field.shape[i + 1] - 1)
       for i in range(query.shape[-1])]
  off = np.nan_to_num(field[(slice(None),) + tuple(q)].T).astype(int)[:, ::-1]
  new_starts = starts + off
  off = off - np.minimum(new_starts, 0)
  shape = np.array(img_shape)[None, ...]
  new_ends = new_starts + np.array(patch)[None, ...]
  off = off - (np.maximum(new_ends, shape) - shape)
  return off


class MaskedXCorrWithStatsCalculator:
  """Oracle twin of JAXMaskedXCorrWithStatsCalculator (flow_field.py:449-712)."""

  non_spatial_flow_channels = 2

  def __init__(self, mean=None, peak_min_distance=2, peak_radius=5):
    self._mean = mean
    self._min_distance = peak_min_distance
    self._peak_radius = peak_radius

  def flow_field(self, pre_image, post_image, patch_size, step, pre_mask=None,
                 post_mask=None, mask_only_for_patch_selection=False,
                 selection_mask=None, max_masked=0.75, batch_size=4096,
                 post_patch_size=None, pre_targeting_field=None,
                 pre_targeting_step=None, post_targeting_field=None,
                 post_targeting_step=None):
    nd = pre_image.ndim
    assert nd == post_image.ndim
    patch_size = _as_tuple(patch_size, nd)
    post_patch_size = (patch_size if post_patch_size is None
                       else _as_tuple(post_patch_size, nd))
    step = _as_tuple(step, nd)
    if pre_targeting_step is not None:
      pre_targeting_step = _as_tuple(pre_targeting_step, nd)
    assert len(patch_size) == nd and len(post_patch_size) == nd
    assert len(step) == nd

    out_shape = (np.array(post_image.shape)
                 - (np.array(post_patch_size) - step)) // step
    out_sel = tuple(slice(0, int(s)) for s in out_shape)
    output = np.full([nd + 2] + out_shape.tolist(), np.nan, dtype=F32)

    sel = (np.ones(out_shape, dtype=bool) if selection_mask is None
           else np.array(selection_mask[out_sel], dtype=bool))
    for mask, psz in ((pre_mask, patch_size), (post_mask, post_patch_size)):
      if mask is not None:
        s = query_integral_image(integral_image(mask), psz, step)
        sel[(s / np.prod(psz) >= max_masked)[out_sel]] = False
    if mask_only_for_patch_selection:
      pre_mask = post_mask = None

    patch_offset = ((np.array(patch_size) - post_patch_size) // 2)[None, ...]
    oyx = np.array(np.where(sel)).T
    for pos in batch(oyx, batch_size):
      real = pos.shape[0]
      proc = (np.pad(pos, ((0, batch_size - real), (0, 0)), mode='edge')
              if real < batch_size else pos)
      post_starts = proc * np.array(step).reshape((1, -1))
      pre_starts = np.clip(post_starts - patch_offset, 0, np.inf).astype(int)

      tg = po = None
      if pre_targeting_field is not None and pre_targeting_step is not None:
        tg = _targeting_offsets(pre_targeting_field, pre_targeting_step,
                                pre_starts, patch_size, pre_image.shape)
        pre_starts = pre_starts + tg
      if post_targeting_field is not None and post_targeting_step is not None:
        po = _targeting_offsets(post_targeting_field, post_targeting_step,
                                post_starts, post_patch_size, post_image.shape)
        post_starts = post_starts + po
      pre_starts = np.clip(pre_starts, 0, np.inf).astype(int)
      post_starts = np.clip(post_starts, 0, np.inf).astype(int)

      peaks = batched_xcorr_peaks(
          pre_image, post_image, pre_mask, post_mask, patch_size, pre_starts,
          self._mean, post_patch_size=post_patch_size,
          min_distance=self._min_distance, peak_radius=self._peak_radius,
          post_starts=post_starts)
      for i, coord in enumerate(pos):
        vec = peaks[i]
        if tg is not None:
          vec[:nd] = vec[:nd] + tg[i, ::-1]
        if po is not None:
          vec[:nd] = vec[:nd] - po[i, ::-1]
        output[(slice(None),) + tuple(coord)] = vec
    return output
