"""NumPy transcription of the data flow of csrc/flow3d_masked.cuh (the masked 3-d correlation
that has not run on hardware yet): the same slots, the same six products in the same
pairing, the same crop and the same Padfield term / normalisation formulas as the kernels,
with numpy.fft in place of the device transforms.  It must reproduce the oracle -- which is
pinned on the reference's NumPy branch -- so a wrong pairing or a mask / flip inconsistency in
the kernel design shows up here, on CPU."""
import numpy as np

from oracle import flow_oracle as fo

F32 = np.float32


def model(prev, curr, pm, cm):
  """prev / curr: [b, d, h, w] fp32 whole patches (mean 0 given, as masked_xcorr passes
  has_mean = 1, mean = 0); pm / cm: bool masks or None."""
  b = prev.shape[0]
  sz = tuple(int(p + c - 1) for p, c in zip(prev.shape[1:], curr.shape[1:]))
  L = tuple(fo.next_fast_len(s) for s in sz)

  def pack(img, mask, which, kind):  # pack3m_kernel
    z = np.zeros((b,) + L, np.complex64)
    valid = np.ones(img.shape, bool) if mask is None else ~mask
    if kind == 1:
      v = valid.astype(F32)
    else:
      v = np.where(valid, img - F32(0.0), F32(0.0)).astype(F32)
      if kind == 2:
        v = v * v
    if which == 1:
      v = v[:, ::-1, ::-1, ::-1]
    z[:, :img.shape[1], :img.shape[2], :img.shape[3]] = v
    return z

  slots = [pack(prev, pm, 0, 0), pack(curr, cm, 1, 0), pack(prev, pm, 0, 1), pack(curr, cm, 1, 1),
           pack(prev, pm, 0, 2), pack(curr, cm, 1, 2)]
  Z = [np.fft.fftn(s, axes=(1, 2, 3)).astype(np.complex64) for s in slots]
  P, C, MP, MC, P2, C2 = Z
  W = [P * C, MC * MP, MC * P, MP * C, MC * P2, MP * C2]          # multiply3m_kernel
  crop = lambda w: np.fft.ifftn(w, axes=(1, 2, 3)).real.astype(F32)[:, :sz[0], :sz[1], :sz[2]]
  xc, ov, mcp, mcc, psq, csq = (crop(w) for w in W)                # crop3m_kernel (scale in ifftn)
  eps = F32(1.1920929e-07)                                         # padfield_terms_kernel
  o = np.maximum(np.rint(ov), eps)
  oi = F32(1.0) / o
  num = xc - mcp * mcc * oi
  pd = np.maximum(psq - mcp * mcp * oi, F32(0))
  cd = np.maximum(csq - mcc * mcc * oi, F32(0))
  den = np.sqrt(pd * cd)
  tol = F32(1e3) * eps * np.abs(den).max()                          # padfield_normalise_kernel
  thr = F32(0.3) * o.max()
  with np.errstate(all='ignore'):
    out = np.where(den > tol, num / den, F32(0))
  out = np.clip(out, -1, 1)
  return np.where(o < thr, F32(0), out).astype(F32)


def test_kernel_data_flow_reproduces_the_oracle():
  rng = np.random.default_rng(12)
  prev = (rng.standard_normal((2, 9, 11, 12)) * 10).astype(F32)
  curr = (rng.standard_normal((2, 9, 11, 12)) * 10).astype(F32)
  pm, cm = rng.random(prev.shape) > 0.8, rng.random(curr.shape) > 0.75
  want = fo.masked_xcorr(prev, curr, pm, cm, dim=3)
  np.testing.assert_allclose(model(prev, curr, pm, cm), want, rtol=0, atol=2e-5)
  # unequal sizes, one-sided mask
  prev = (rng.standard_normal((1, 10, 12, 9)) * 7).astype(F32)
  curr = (rng.standard_normal((1, 5, 6, 7)) * 7).astype(F32)
  pm = rng.random(prev.shape) > 0.7
  want = fo.masked_xcorr(prev, curr, pm, None, dim=3)
  np.testing.assert_allclose(model(prev, curr, pm, None), want, rtol=0, atol=2e-5)
