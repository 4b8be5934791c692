"""Generates the committed golden vectors by running the REFERENCE'S OWN source.

  python tests/golden/make_golden.py            # needs /root/reference

/root/reference/mesh.py and /root/reference/flow_field.py are imported unmodified
through tests/golden/_jax_shim.py (NumPy stand-ins for the JAX primitives; JAX
cannot be installed in this image) and executed on seeded inputs.  Inputs and
outputs are written to tests/golden/*.npz.  /root/reference does not exist on the
GPU box, so tests only ever read the .npz files.
"""

from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import scipy.ndimage as ndi

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, '..', '..'))
import _jax_shim as shim  # pylint: disable=g-import-not-at-top

warnings.filterwarnings('ignore')


def texture(seed, shape, sigma=2.0):
  rng = np.random.default_rng(seed)
  base = ndi.gaussian_filter(rng.standard_normal(shape), sigma)
  base = (base - base.min()) / (base.max() - base.min()) * 255
  return base.astype(np.uint8)


def shifted_pair(seed, size, dy, dx, noise=5.0, margin=48):
  rng = np.random.default_rng(seed + 1000)
  base = texture(seed, (size + 2 * margin, size + 2 * margin))
  pre = base[margin:margin + size, margin:margin + size]
  post = base[margin + dy:margin + dy + size, margin + dx:margin + dx + size]
  post = np.clip(post.astype(float) + rng.normal(0, noise, post.shape), 0, 255)
  return np.ascontiguousarray(pre), post.astype(np.uint8)


def mesh_cases(mesh):
  out = {}
  rng = np.random.default_rng(42)

  # -- force fields -------------------------------------------------------------
  x2 = (rng.standard_normal((2, 2, 17, 19)) * 3).astype(np.float32)
  x2[0, 0, 3, 4] = np.nan
  x2[1, 1, 9, 0] = np.nan
  out['force2d_x'] = x2
  for poo in (0, 1):
    out[f'force2d_poo{poo}'] = np.asarray(
        mesh.inplane_force(shim.asjax(x2), 0.1, (40.0, 30.0), bool(poo)))
  x3 = (rng.standard_normal((3, 2, 6, 7, 8)) * 3).astype(np.float32)
  out['force3d_x'] = x3
  for poo in (0, 1):
    out[f'force3d_poo{poo}'] = np.asarray(
        mesh.elastic_mesh_3d(shim.asjax(x3), 0.1, (40.0, 40.0, 14.0), bool(poo)))

  # -- chunked relaxations (state after every velocity_verlet call) -------------
  def run(tag, x0, prev, cfg_kwargs, force='inplane', max_chunks=4):
    cfg = mesh.IntegrationConfig(**cfg_kwargs)
    fn = mesh.inplane_force if force == 'inplane' else mesh.elastic_mesh_3d
    x = shim.asjax(x0)
    v = shim.asjax(np.zeros_like(x0))
    p = None if prev is None else shim.asjax(prev)
    dt, alpha, cap = cfg.dt, cfg.alpha, cfg.start_cap
    xs, vs, scal = [], [], []
    for _ in range(max_chunks):
      state = mesh.velocity_verlet(
          shim.asjax(x), shim.asjax(v), p, cfg, force_cap=cap, fire_dt=dt,
          fire_alpha=alpha, mesh_force=fn)
      x, v = state[:2]
      n_pos = -1
      if cfg.fire:
        dt, alpha, n_pos, cap = state[-4:]
      xs.append(np.asarray(x).copy())
      vs.append(np.asarray(v).copy())
      scal.append([float(dt), float(alpha), float(n_pos), float(cap)])
    out[f'{tag}_x0'] = np.asarray(x0, np.float32)
    if prev is not None:
      out[f'{tag}_prev'] = np.asarray(prev, np.float32)
    out[f'{tag}_xs'] = np.stack(xs)
    out[f'{tag}_vs'] = np.stack(vs)
    out[f'{tag}_scalars'] = np.array(scal, np.float64)
    out[f'{tag}_cfg'] = np.array(repr(cfg_kwargs))

  shape = (2, 2, 40, 56)
  prev = ndi.gaussian_filter(rng.standard_normal(shape), (0, 0, 4, 4)) * 30
  prev = prev.astype(np.float32)
  prev[:, 0, 5:8, 10:12] = np.nan
  prev[0, 1, 20, 30] = np.nan
  base = dict(dt=0.001, gamma=0.0, k0=0.1, k=0.1, stride=(40.0, 40.0),
              num_iters=60, max_iters=240, stop_v_max=0.0, fire=True,
              dt_max=1000.0)
  run('fire_poo', np.zeros(shape, np.float32), prev,
      dict(base, prefer_orig_order=True))
  run('fire_cap', np.zeros(shape, np.float32), prev,
      dict(base, start_cap=0.01, final_cap=10.0, cap_upscale_every=7,
           prefer_orig_order=True))
  # remove_drift: the mean's fp32 summation order is implementation-defined, so this
  # case keeps dt <= 0.1 (well inside the Verlet stability limit) where rounding
  # differences are damped instead of amplified.
  run('fire_drift', (rng.standard_normal(shape) * 0.5).astype(np.float32), prev,
      dict(base, k0=0.02, dt_max=100.0, remove_drift=True))
  run('damped', np.zeros(shape, np.float32), prev,
      dict(base, fire=False, dt=0.05, gamma=0.5, num_iters=40))
  x0n = np.zeros(shape, np.float32)
  x0n[:, 1, 0:3, 0:3] = np.nan  # invalid nodes stay invalid and exert no force
  run('fire_nan_x', x0n, prev, dict(base, prefer_orig_order=False))
  shape3 = (3, 6, 10, 12)
  prev3 = (ndi.gaussian_filter(rng.standard_normal(shape3), (0, 2, 2, 2)) * 20
           ).astype(np.float32)
  run('fire_3d', np.zeros(shape3, np.float32), prev3,
      dict(base, stride=(40.0, 40.0, 30.0), prefer_orig_order=True, num_iters=30),
      force='3d', max_chunks=3)

  # -- the reference's own relax KATs (tests/mesh_test.py:25-65), full run --------
  xk = np.zeros((2, 1, 50, 50), np.float32)
  xk[0, 0, 20:30, 10] = 3
  xk[0, 0, 20:30, 40] = -4
  xk[1, 0, 30, 10:20] = 2
  for tag, kw in (('kat_fire', dict(gamma=0.0, fire=True)),
                  ('kat_damped', dict(gamma=0.9 * np.sqrt(4 * 0.1), fire=False))):
    cfg = mesh.IntegrationConfig(dt=0.01, k0=0.1, k=0.1, stride=(10, 10),
                                 num_iters=100, max_iters=10000,
                                 stop_v_max=0.001, **kw)
    xr, ek, t = mesh.relax_mesh(shim.asjax(xk), shim.asjax(np.zeros_like(xk)), cfg)
    out[f'{tag}_x0'] = xk
    out[f'{tag}_x'] = np.asarray(xr)
    out[f'{tag}_ekin'] = np.array(ek, np.float64)
    out[f'{tag}_t'] = np.array(t)
  return out


def flow_cases(ff):
  out = {}
  calc = ff.JAXMaskedXCorrWithStatsCalculator()

  def put(tag, result, **inputs):
    for k, v in inputs.items():
      if v is not None:
        out[f'{tag}_{k}'] = np.asarray(v)
    out[f'{tag}_flow'] = np.asarray(result, np.float32)

  # reference KATs, tests/flow_field_test.py:24-56
  pre = np.zeros((120, 120), np.uint8)
  post = np.zeros((120, 120), np.uint8)
  pre[60, 60] = 255
  post[70, 53] = 255
  put('kat_delta', calc.flow_field(pre, post, patch_size=80, step=40,
                                   batch_size=4), pre=pre, post=post)
  post2 = post.copy()
  post2[54, 68] = 255
  pmask = np.zeros((128, 128), bool)
  pmask[:55, :70] = 1
  put('kat_delta_mask', calc.flow_field(pre, post2, patch_size=80, step=40,
                                        post_mask=pmask, batch_size=4),
      pre=pre, post=post2, post_mask=pmask)
  # tests/flow_field_test.py:58-72
  pre3 = np.zeros((50, 100, 100), np.uint8)
  post3 = np.zeros((50, 100, 100), np.uint8)
  pre3[25, 50, 50] = 255
  post3[22, 45, 54] = 255
  put('kat_3d', calc.flow_field(pre3, post3, patch_size=(40, 80, 80), step=10,
                                batch_size=1))
  # tests/flow_field_test.py:96-125
  pre = np.zeros((120, 120), np.uint8)
  post = np.zeros((120, 120), np.uint8)
  pre[50, 55] = 255
  post[100, 100] = 255
  put('kat_notarget', calc.flow_field(pre, post, patch_size=80, step=40,
                                      batch_size=4), pre=pre, post=post)
  tgt = np.full((2, 2, 2), 40.0, np.float32)
  put('kat_target', calc.flow_field(pre, post, patch_size=80, step=40,
                                    batch_size=4, post_targeting_field=tgt,
                                    post_targeting_step=40), tgt=tgt)

  # textured pairs: EM-2D shape (patch 160 / step 40), several batch sizes so that
  # the batch-coupled 2nd-peak rule and the 'edge' padding are exercised.
  pre, post = shifted_pair(0, 360, 5, -3)
  for bs in (4, 7, 64):
    put(f'tex160_b{bs}', calc.flow_field(pre, post, 160, 40, batch_size=bs),
        pre=pre if bs == 4 else None, post=post if bs == 4 else None)

  # weak texture with a periodic component -> several competing peaks.
  yy, xx = np.mgrid[:280, :280]
  rng = np.random.default_rng(7)
  per = 127 + 60 * np.sin(2 * np.pi * xx / 23.0) * np.sin(2 * np.pi * yy / 31.0)
  pre_p = np.clip(per + rng.normal(0, 25, per.shape), 0, 255).astype(np.uint8)
  post_p = np.clip(np.roll(per, (4, -6), (0, 1)) + rng.normal(0, 25, per.shape),
                   0, 255).astype(np.uint8)
  put('periodic', calc.flow_field(pre_p, post_p, 120, 40, batch_size=5),
      pre=pre_p, post=post_p)

  # masked (Padfield) path with both masks, rectangular patches, non-square step.
  pre, post = shifted_pair(3, 300, -7, 4)
  rng = np.random.default_rng(11)
  m_pre = ndi.gaussian_filter(rng.standard_normal(pre.shape), 12) > 0.012
  m_post = ndi.gaussian_filter(rng.standard_normal(post.shape), 12) > 0.012
  put('masked', calc.flow_field(pre, post, (96, 128), (32, 40), pre_mask=m_pre,
                                post_mask=m_post, batch_size=6),
      pre=pre, post=post, pre_mask=m_pre, post_mask=m_post)
  put('masked_selonly', calc.flow_field(
      pre, post, (96, 128), (32, 40), pre_mask=m_pre, post_mask=m_post,
      mask_only_for_patch_selection=True, max_masked=0.4, batch_size=6))

  # post patch smaller than pre patch (EstimateMissingFlow regime), selection
  # mask, float32 images, constant mean.
  sel = np.ones((9, 9), bool)
  sel[::2, 1::2] = False
  pre_f = pre.astype(np.float32) / 3.0
  post_f = post.astype(np.float32) / 3.0
  calc_m = ff.JAXMaskedXCorrWithStatsCalculator(mean=40.0, peak_radius=(3, 4))
  put('postpatch', calc_m.flow_field(pre_f, post_f, 128, 24, post_patch_size=96,
                                     selection_mask=sel, batch_size=16),
      pre=pre_f, post=post_f, sel=sel)

  # pre-targeting field with NaNs.
  tg = np.zeros((2, 8, 8), np.float32)
  tg[0], tg[1] = 6.0, -9.0
  tg[:, 2, 3] = np.nan
  put('pretarget', calc.flow_field(pre, post, 128, 24, pre_targeting_field=tg,
                                   pre_targeting_step=24, batch_size=32), tg=tg)

  # _batched_peaks on the analytic bump, tests/flow_field_test.py:74-94
  hy, hx = np.mgrid[:50, :50]
  r = np.sqrt(2 * (28 - hx) ** 2 + (20 - hy) ** 2)
  bump = (10 * np.exp(-r / 4)).astype(np.float32)
  out['peaks_bump_img'] = bump
  out['peaks_bump'] = np.asarray(ff._batched_peaks(
      shim.asjax(bump[np.newaxis]), (25, 25), min_distance=2, threshold_rel=0.5,
      peak_radius=(2, 3)))
  return out


def main():
  mesh = shim.load_reference('mesh')
  ff = shim.load_reference('flow_field')
  np.savez_compressed(os.path.join(HERE, 'mesh_golden.npz'), **mesh_cases(mesh))
  np.savez_compressed(os.path.join(HERE, 'flow_golden.npz'), **flow_cases(ff))
  for f in ('mesh_golden.npz', 'flow_golden.npz'):
    print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, 'KiB')


if __name__ == '__main__':
  main()
