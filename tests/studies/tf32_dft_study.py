"""Numerical study for DESIGN.md section 8 (item 3): if the per-patch correlation were
computed as DFT-as-GEMM on the tensor cores, how often would the integer flow vectors and
the peak statistics differ from the fp32 FFT the reference computes?

Emulation (NumPy, CPU): every 320-point transform is done as the two GEMM passes a
tcgen05 kernel would run (16-point and 20-point DFT matrices with the twiddle in between),
with the OPERANDS of every GEMM rounded to the tensor-core input format and fp32
accumulation:
  tf32     one pass, operands rounded to 10 mantissa bits
  tf32x3   the 3-term split (hi*hi + hi*lo + lo*hi), i.e. fp32-grade products
The correlation images then go through the oracle's peak search (oracle/flow_oracle.py);
the study is test infrastructure (it lives under tests/ because it uses the oracle).

  python tests/studies/tf32_dft_study.py            # prints one JSON line per workload
"""
import json
import os
import sys

import numpy as np
import scipy.ndimage as ndi

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import flow_oracle as fo  # noqa: E402

N1, N2, L = 16, 20, 320


def tf32(a):
  """Round-to-nearest-even to 10 explicit mantissa bits (fp32 container)."""
  a = np.ascontiguousarray(a, dtype=np.float32)
  u = a.view(np.uint32).astype(np.uint64)
  u = (u + 0xFFF + ((u >> 13) & 1)) & ~np.uint64(0x1FFF)
  return u.astype(np.uint32).view(np.float32).reshape(a.shape)


def gemm(a, b, mode):
  """Real GEMM a @ b with tensor-core operand rounding, fp32 accumulate (emulated in
  float64 then rounded once -- the accumulation error of a 16/20-term fp32 sum is far
  below the operand rounding studied here)."""
  if mode == 'fp32':
    return (a.astype(np.float64) @ b.astype(np.float64)).astype(np.float32)
  ah, bh = tf32(a), tf32(b)
  acc = ah.astype(np.float64) @ bh.astype(np.float64)
  if mode == 'tf32x3':
    al, bl = tf32(a - ah), tf32(b - bh)
    acc += ah.astype(np.float64) @ bl.astype(np.float64) + al.astype(np.float64) @ bh.astype(np.float64)
  return acc.astype(np.float32)


def cgemm(ar, ai, br, bi, mode):
  return (gemm(ar, br, mode) - gemm(ai, bi, mode), gemm(ar, bi, mode) + gemm(ai, br, mode))


def dft_mats(n, sign):
  k = np.arange(n)
  w = np.exp(sign * 2j * np.pi * np.outer(k, k) / n)
  return w.real.astype(np.float32), w.imag.astype(np.float32)


def fft_last_axis(xr, xi, sign, mode):
  """Length-320 DFT along the last axis as two GEMM passes: x[N2 n1 + n2] -> X[k1 + 16 k2]."""
  shp = xr.shape[:-1]
  xr = xr.reshape(-1, N1, N2)
  xi = xi.reshape(-1, N1, N2)
  w1r, w1i = dft_mats(N1, sign)
  # pass 1: 16-point DFT over n1 for every n2:  A[k1, n2] = sum_n1 W16[k1, n1] x[n1, n2]
  ar, ai = cgemm(w1r[None], w1i[None], xr, xi, mode)
  # twiddle W_L^(n2 k1) in fp32 (CUDA cores / epilogue)
  t = np.exp(sign * 2j * np.pi * np.outer(np.arange(N1), np.arange(N2)) / L)
  tr, ti = t.real.astype(np.float32), t.imag.astype(np.float32)
  ar, ai = ar * tr - ai * ti, ar * ti + ai * tr
  # pass 2: 20-point DFT over n2:  X[k1, k2] = sum_n2 A[k1, n2] W20[n2, k2]
  w2r, w2i = dft_mats(N2, sign)
  yr, yi = cgemm(ar, ai, w2r[None], w2i[None], mode)
  # X[k1 + 16 k2]
  yr = yr.transpose(0, 2, 1).reshape(shp + (L,))
  yi = yi.transpose(0, 2, 1).reshape(shp + (L,))
  return yr, yi


def fft2(xr, xi, sign, mode):
  yr, yi = fft_last_axis(xr, xi, sign, mode)
  yr, yi = fft_last_axis(yr.swapaxes(-1, -2).copy(), yi.swapaxes(-1, -2).copy(), sign, mode)
  return yr.swapaxes(-1, -2), yi.swapaxes(-1, -2)


def xcorr(pre, post, mode):
  """flow_field.py:66-89 (no masks) for [b, 160, 160] mean-subtracted patches."""
  b, p = pre.shape[0], pre.shape[1]
  a = np.zeros((b, L, L), np.float32)
  c = np.zeros((b, L, L), np.float32)
  a[:, :p, :p] = pre
  c[:, :p, :p] = post[:, ::-1, ::-1]
  z = np.zeros_like(a)
  ar, ai = fft2(a, z, -1, mode)
  cr, ci = fft2(c, z, -1, mode)
  pr, pi = ar * cr - ai * ci, ar * ci + ai * cr     # fp32 product (epilogue)
  xr, _ = fft2(pr, pi, +1, mode)
  return (xr / np.float32(L * L))[:, :2 * p - 1, :2 * p - 1]


def patches(pre_img, post_img, patch=160, step=40):
  g = (pre_img.shape[0] - (patch - step)) // step
  pre, post = [], []
  for y in range(g):
    for x in range(g):
      a = pre_img[y * step:y * step + patch, x * step:x * step + patch].astype(np.float32)
      c = post_img[y * step:y * step + patch, x * step:x * step + patch].astype(np.float32)
      pre.append(a - a.mean(dtype=np.float32))
      post.append(c - c.mean(dtype=np.float32))
  return np.stack(pre), np.stack(post)


def workload(name, size, noise, sigma, seed, blank=False):
  rng = np.random.default_rng(seed)
  base = ndi.gaussian_filter(rng.standard_normal((size + 128, size + 128)), sigma)
  base = ((base - base.min()) / np.ptp(base) * 255)
  pre = np.clip(base[64:64 + size, 64:64 + size], 0, 255).astype(np.uint8)
  post = np.clip(base[69:69 + size, 61:61 + size] + rng.normal(0, noise, (size, size)), 0, 255)
  post = post.astype(np.uint8)
  if blank:
    post[size // 3:2 * size // 3, size // 4:3 * size // 4] = 0
  return name, pre, post


def main():
  center = np.array([159, 159])
  for name, pre_img, post_img in (
      workload('config-1 texture (sigma 2, noise 5)', 512, 5.0, 2.0, 0),
      workload('heavy noise (sigma 2, noise 60)', 512, 60.0, 2.0, 1),
      workload('smooth low-contrast (sigma 8, noise 20)', 512, 20.0, 8.0, 2),
      workload('partly blank post tile', 512, 5.0, 2.0, 3, blank=True)):
    pre, post = patches(pre_img, post_img)
    ref = fo.batched_peaks(xcorr(pre, post, 'fp32'), center, 2, 0.5, 5)
    line = {'workload': name, 'pairs': int(pre.shape[0])}
    for mode in ('tf32', 'tf32x3'):
      xc = xcorr(pre, post, mode)
      got = fo.batched_peaks(xc, center, 2, 0.5, 5)
      ok = ~np.isnan(ref[:, 0])
      same_nan = bool(np.array_equal(np.isnan(ref), np.isnan(got)))
      diff_xy = int(np.sum(np.any(ref[ok, :2] != got[ok, :2], axis=1)))
      with np.errstate(invalid='ignore', divide='ignore'):
        rel = np.abs(got[ok, 2:] - ref[ok, 2:]) / np.maximum(np.abs(ref[ok, 2:]), 1e-6)
      line[mode] = {'same_nan_pattern': same_nan, 'pairs_with_different_xy': diff_xy,
                    'max_rel_err_sharpness': float(np.nanmax(rel[:, 0])) if ok.any() else None,
                    'max_rel_err_ratio': float(np.nanmax(rel[:, 1])) if ok.any() else None}
    print(json.dumps(line))


if __name__ == '__main__':
  main()
