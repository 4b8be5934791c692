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


def stitch_cases(mu, se):
  """compose_maps_fast (map_utils.py:616-734) and the stitching target mesh
  (stitch_elastic.py:285-676: aggregate_arrays is plain NumPy in the reference and
  builds the neighbour table; compute_target_mesh runs on the shim)."""
  out = {}
  rng = np.random.default_rng(7)

  # -- compose_maps_fast, both modes, fractional strides / starts ---------------
  m1 = (rng.standard_normal((2, 2, 9, 11)) * 12).astype(np.float32)
  m1[:, 0, 3, 4] = np.nan
  m2 = (rng.standard_normal((2, 2, 13, 10)) * 6).astype(np.float32)
  m2[0, 1, 5, 5] = np.nan
  out['cmf_map1'], out['cmf_map2'] = m1, m2
  out['cmf_args'] = np.array([3, 2, 1, 1, 4, 0], np.int64)  # start1 zyx, start2 zyx
  for mode in ('nearest', 'constant'):
    out[f'cmf_{mode}'] = np.asarray(mu.compose_maps_fast(
        shim.asjax(m1), (3, 2, 1), (20, 25), shim.asjax(m2), (1, 4, 0), (16, 20),
        mode=mode))
  m13 = (rng.standard_normal((3, 4, 5, 6)) * 5).astype(np.float32)
  m23 = (rng.standard_normal((3, 5, 6, 7)) * 3).astype(np.float32)
  m13[:, 1, 2, 3] = np.nan
  out['cmf3_map1'], out['cmf3_map2'] = m13, m23
  for mode in ('nearest', 'constant'):
    out[f'cmf3_{mode}'] = np.asarray(mu.compose_maps_fast(
        shim.asjax(m13), (1, 0, 2), (4, 10, 10), shim.asjax(m23), (0, 1, 1), (4, 10, 10),
        mode=mode))

  # -- 2-d stitching: 3 x 2 tiles, mesh 12 x 14 nodes, stride 20 -----------------
  def smooth(shape, amp):
    return (ndi.gaussian_filter(rng.standard_normal(shape), (0,) * (len(shape) - 2) + (2, 2))
            * amp).astype(np.float32)

  nx_t, ny_t, stride = 3, 2, (20, 20)
  tile_shape = (240, 280)
  coords = [(tx, ty) for ty in range(ny_t) for tx in range(nx_t) if (tx, ty) != (2, 1)]
  cx = np.full((2, ny_t, nx_t), np.nan)
  cy = np.full((2, ny_t, nx_t), np.nan)
  fine_x, fine_y, off_x, off_y = {}, {}, {}, {}
  sizes_x = {(0, 0): (12, 4), (1, 0): (11, 3), (0, 1): (10, 4)}
  sizes_y = {(0, 0): (3, 14), (1, 0): (4, 13), (2, 0): (3, 12)}
  for k, (oy, ox) in sizes_x.items():
    f = np.full((4, oy, ox), np.nan, np.float32)
    f[:2] = smooth((2, oy, ox), 6.0)
    f[2:] = 1.0
    f[:, rng.integers(oy), rng.integers(ox)] = np.nan
    fine_x[k] = f
    cx[:, k[1], k[0]] = (260 + rng.integers(-5, 5), rng.integers(-30, 30))
    off_x[k] = (int(rng.integers(-3, 3)), int(rng.integers(-3, 3)))
  for k, (oy, ox) in sizes_y.items():
    if (k[0], k[1] + 1) not in coords:
      continue
    f = np.full((4, oy, ox), np.nan, np.float32)
    f[:2] = smooth((2, oy, ox), 6.0)
    f[2:] = 1.0
    fine_y[k] = f
    cy[:, k[1], k[0]] = (rng.integers(-30, 30), 220 + rng.integers(-5, 5))
    off_y[k] = (int(rng.integers(-3, 3)), int(rng.integers(-3, 3)))
  coarse_mesh = np.zeros((2, ny_t, nx_t))
  fx, fy, x, nbors, key_to_idx = se.aggregate_arrays(
      (cx, fine_x, off_x), (cy, fine_y, off_y), coords, coarse_mesh, stride, tile_shape)
  # inputs of aggregate_arrays, for the product's own aggregate_arrays
  out['st2_coords'] = np.array(coords)
  out['st2_cx'], out['st2_cy'] = cx, cy
  for nm, fine, offs in (('x', fine_x, off_x), ('y', fine_y, off_y)):
    for k, f in fine.items():
      out[f'st2_fine{nm}_{k[0]}_{k[1]}'] = f
      out[f'st2_off{nm}_{k[0]}_{k[1]}'] = np.array(offs[k])
  out['st2_x0'] = x.astype(np.float32)
  x = x + smooth(x.shape, 8.0)
  x[:, 1, 4, 5] = np.nan
  fx, fy, x = fx.astype(np.float32), fy.astype(np.float32), x.astype(np.float32)
  out['st2_fx'], out['st2_fy'], out['st2_x'], out['st2_nbors'] = fx, fy, x, nbors
  out['st2_stride'] = np.array(stride)
  res = [np.asarray(se.compute_target_mesh(shim.asjax(nb), shim.asjax(x), shim.asjax(fx),
                                           shim.asjax(fy), stride)) for nb in nbors]
  out['st2_target'] = np.transpose(np.stack(res), [1, 0, 2, 3])

  # -- the whole stitching relaxation of notebooks/em_stitching.ipynb:545-603 run by
  # the reference's own mesh.relax_mesh with the notebook's prev_fn --------------
  import functools as ft
  import jax
  import jax.numpy as jnp
  mesh = shim.load_reference('mesh')
  nb_j, fx_j, fy_j = shim.asjax(nbors), shim.asjax(fx), shim.asjax(fy)

  def prev_fn(xx):
    target_fn = ft.partial(se.compute_target_mesh, x=xx, fx=fx_j, fy=fy_j, stride=stride)
    res = jax.vmap(target_fn)(nb_j)
    return jnp.transpose(res, [1, 0, 2, 3])

  cfg = dict(dt=0.001, gamma=0., k0=0.01, k=0.1, stride=stride, num_iters=12, max_iters=36,
             stop_v_max=0.0, dt_max=100, prefer_orig_order=True, start_cap=0.1,
             final_cap=10., remove_drift=True)
  x_start = np.nan_to_num(x)  # the notebook starts from the (NaN-free) coarse mesh
  out['st2_relax_cfg'] = np.array(repr(cfg))
  out['st2_relax_x0'] = x_start
  xr, ekin, t = mesh.relax_mesh(shim.asjax(x_start), None, mesh.IntegrationConfig(**cfg),
                                prev_fn=prev_fn)
  out['st2_relax_x'], out['st2_relax_ekin'], out['st2_relax_t'] = (
      np.asarray(xr), np.asarray(ekin, dtype=np.float64), np.array(t))
  cfg_nd = dict(cfg, remove_drift=False, fire=False, gamma=0.3, start_cap=10., final_cap=10.)
  out['st2_relax_damped_cfg'] = np.array(repr(cfg_nd))
  xr, ekin, t = mesh.relax_mesh(shim.asjax(x_start), None, mesh.IntegrationConfig(**cfg_nd),
                                prev_fn=prev_fn)
  out['st2_relax_damped_x'], out['st2_relax_damped_ekin'] = (
      np.asarray(xr), np.asarray(ekin, dtype=np.float64))

  # -- compute_flow_map (stitch_elastic.py:197-282): 3 x 2 tiles cut from one texture ----
  tex = texture(21, (420, 560), sigma=1.5)
  th, tw = 160, 200
  nominal = {(tx, ty): (ty * 130 + [0, 4, -3][tx] * (ty > 0), tx * 170 + [0, -5][ty] * (tx > 0))
             for tx in range(3) for ty in range(2)}
  tiles = {k: np.ascontiguousarray(tex[y0:y0 + th, x0:x0 + tw]) for k, (y0, x0) in nominal.items()}
  cxm = np.full((2, 2, 3), np.nan)
  cym = np.full((2, 2, 3), np.nan)
  for (tx, ty), (y0, x0) in nominal.items():
    if (tx + 1, ty) in nominal:
      y1, x1 = nominal[tx + 1, ty]
      cxm[:, ty, tx] = (x1 - x0 - tw, y1 - y0)
    if (tx, ty + 1) in nominal:
      y1, x1 = nominal[tx, ty + 1]
      cym[:, ty, tx] = (x1 - x0, y1 - y0 - th)
  out['fm2_tex'] = tex
  out['fm2_nominal'] = np.array([[tx, ty, y0, x0] for (tx, ty), (y0, x0) in nominal.items()])
  out['fm2_cx'], out['fm2_cy'] = cxm, cym
  for axis, cm in ((0, cxm), (1, cym)):
    fl, of = se.compute_flow_map(tiles, cm, axis, patch_size=(32, 32), stride=(8, 8),
                                 batch_size=64)
    for k in fl:
      out[f'fm2_flow{axis}_{k[0]}_{k[1]}'] = np.asarray(fl[k])
      out[f'fm2_off{axis}_{k[0]}_{k[1]}'] = np.array(of[k])

  # -- compute_coarse_offsets (stitch_rigid.py:104-273) on the same tile grid ----------
  sr = shim.load_reference('stitch_rigid')
  cox, coy = sr.compute_coarse_offsets((2, 3), tiles, overlaps_xy=((30, 44), (30, 44)),
                                       min_range=(10, 100, 0), min_overlap=16, filter_size=5)
  out['co_conn_x'], out['co_conn_y'] = np.asarray(cox), np.asarray(coy)

  # -- 3-d stitching (LICONN): 2 x 2 tiles, mesh 4 x 6 x 7 nodes, stride (8, 20, 20) --
  stride3 = (8, 20, 20)
  coords3 = [(0, 0), (1, 0), (0, 1), (1, 1)]
  cx3 = np.full((3, 2, 2), np.nan)
  cy3 = np.full((3, 2, 2), np.nan)
  fine_x3, fine_y3, off_x3, off_y3 = {}, {}, {}, {}
  for k, shp in {(0, 0): (4, 6, 3), (0, 1): (3, 5, 2)}.items():
    f = np.full((5,) + shp, np.nan, np.float32)
    f[:3] = (rng.standard_normal((3,) + shp) * 3).astype(np.float32)
    fine_x3[k] = f
    cx3[:, k[1], k[0]] = (120 + rng.integers(-3, 3), rng.integers(-15, 15), rng.integers(-6, 6))
    off_x3[k] = tuple(int(v) for v in rng.integers(-2, 3, 3))
  for k, shp in {(0, 0): (4, 2, 7), (1, 0): (3, 3, 6)}.items():
    f = np.full((5,) + shp, np.nan, np.float32)
    f[:3] = (rng.standard_normal((3,) + shp) * 3).astype(np.float32)
    fine_y3[k] = f
    cy3[:, k[1], k[0]] = (rng.integers(-15, 15), 100 + rng.integers(-3, 3), rng.integers(-6, 6))
    off_y3[k] = tuple(int(v) for v in rng.integers(-2, 3, 3))
  fx3, fy3, x3, nbors3, _ = se.aggregate_arrays(
      (cx3, fine_x3, off_x3), (cy3, fine_y3, off_y3), coords3, np.zeros((3, 2, 2)), stride3,
      (32, 120, 140))
  x3 = x3 + (rng.standard_normal(x3.shape) * 4)
  fx3, fy3, x3 = fx3.astype(np.float32), fy3.astype(np.float32), x3.astype(np.float32)
  out['st3_fx'], out['st3_fy'], out['st3_x'], out['st3_nbors'] = fx3, fy3, x3, nbors3
  out['st3_stride'] = np.array(stride3)
  res = [np.asarray(se.compute_target_mesh(shim.asjax(nb), shim.asjax(x3), shim.asjax(fx3),
                                           shim.asjax(fy3), stride3)) for nb in nbors3]
  out['st3_target'] = np.transpose(np.stack(res), [1, 0, 2, 3, 4])

  # notebooks/liconn_inplane_stitching.ipynb:763-783: 3-d relaxation with prev_fn.
  nb3_j, fx3_j, fy3_j = shim.asjax(nbors3), shim.asjax(fx3), shim.asjax(fy3)

  def prev_fn3(xx):
    target_fn = ft.partial(se.compute_target_mesh, x=xx, fx=fx3_j, fy=fy3_j, stride=stride3)
    res = jax.vmap(target_fn)(nb3_j)
    return jnp.transpose(res, [1, 0, 2, 3, 4])

  # mesh.elastic_mesh_3d takes the stride in xyz order; the target in zyx.
  cfg3 = dict(dt=0.001, gamma=0., k0=0.01, k=0.1, stride=stride3[::-1], num_iters=8,
              max_iters=24, stop_v_max=0.0, dt_max=100, prefer_orig_order=False,
              start_cap=0.1, final_cap=10., remove_drift=True)
  out['st3_relax_cfg'] = np.array(repr(cfg3))
  for tag, kw in (('st3_relax', {}), ('st3_relax_nodrift', dict(remove_drift=False))):
    c = dict(cfg3, **kw)
    out[f'{tag}_cfg'] = np.array(repr(c))
    xr, ekin, t = mesh.relax_mesh(shim.asjax(x3), None, mesh.IntegrationConfig(**c),
                                  prev_fn=prev_fn3, mesh_force=mesh.elastic_mesh_3d)
    out[f'{tag}_x'], out[f'{tag}_ekin'] = np.asarray(xr), np.asarray(ekin, dtype=np.float64)
  return out


def flow3d_cases(se):
  """stitch_elastic.compute_flow_map3d (stitch_elastic.py:84-193), the fine-flow step of
  notebooks/liconn_inplane_stitching.ipynb: 2 x 2 tiles of [1, 24, 56, 64] voxels cut from
  one 3-d texture with a jittered nominal grid, 3-d patches."""
  out = {}
  rng = np.random.default_rng(31)
  vol = ndi.gaussian_filter(rng.standard_normal((40, 120, 140)), (1.0, 1.5, 1.5))
  vol = ((vol - vol.min()) / (vol.max() - vol.min()) * 255).astype(np.uint8)
  tz, th, tw = 24, 56, 64
  # nominal (z0, y0, x0) of tile (tx, ty): ~25 % overlap plus jitter in all three axes
  nominal = {(0, 0): (6, 4, 5), (1, 0): (8, 7, 5 + 46), (0, 1): (5, 4 + 41, 8),
             (1, 1): (7, 6 + 41, 4 + 47)}
  tiles = {k: np.ascontiguousarray(vol[z0:z0 + tz, y0:y0 + th, x0:x0 + tw])[None]
           for k, (z0, y0, x0) in nominal.items()}
  cxm = np.full((3, 1, 2, 2), np.nan)
  cym = np.full((3, 1, 2, 2), np.nan)
  for (tx, ty), (z0, y0, x0) in nominal.items():
    if (tx + 1, ty) in nominal:
      z1, y1, x1 = nominal[tx + 1, ty]
      cxm[:, 0, ty, tx] = (x1 - x0 - tw, y1 - y0, z1 - z0)
    if (tx, ty + 1) in nominal:
      z1, y1, x1 = nominal[tx, ty + 1]
      cym[:, 0, ty, tx] = (x1 - x0, y1 - y0 - th, z1 - z0)
  out['fm3_vol'] = vol
  out['fm3_nominal'] = np.array([[tx, ty, z0, y0, x0] for (tx, ty), (z0, y0, x0) in nominal.items()])
  out['fm3_tile_zyx'] = np.array([tz, th, tw])
  out['fm3_cx'], out['fm3_cy'] = cxm, cym
  for axis, cm in ((0, cxm), (1, cym)):
    fl, of = se.compute_flow_map3d(tiles, (tw, th, tz), cm, axis, patch_size=(12, 16, 16),
                                   stride=(4, 8, 8), batch_size=16)
    for k in fl:
      out[f'fm3_flow{axis}_{k[0]}_{k[1]}'] = np.asarray(fl[k])
      out[f'fm3_off{axis}_{k[0]}_{k[1]}'] = np.array(of[k])
  return out


def coarse_cases(sr, mesh):
  """stitch_rigid.{interpolate_missing_offsets, elastic_tile_mesh[_3d], optimize_coarse_mesh}
  (stitch_rigid.py:277-545): the rigid tile-grid step between the coarse offsets and the
  fine flow in both stitching notebooks."""
  out = {}
  rng = np.random.default_rng(41)
  ny, nx = 3, 4
  # offsets as compute_coarse_offsets returns them: [2, 1, y, x], NaN where there is no
  # neighbour, a jittered nominal overlap elsewhere
  cx = np.full((2, 1, ny, nx), np.nan)
  cy = np.full((2, 1, ny, nx), np.nan)
  cx[0, 0, :, :-1] = -40 + rng.integers(-6, 7, (ny, nx - 1))
  cx[1, 0, :, :-1] = rng.integers(-8, 9, (ny, nx - 1))
  cy[0, 0, :-1, :] = rng.integers(-8, 9, (ny - 1, nx))
  cy[1, 0, :-1, :] = -30 + rng.integers(-6, 7, (ny - 1, nx))
  out['cm2_cx'], out['cm2_cy'] = cx, cy
  x = (rng.standard_normal((2, 1, ny, nx)) * 5).astype(np.float32)
  out['cm2_x'] = x
  out['cm2_force'] = np.asarray(sr.elastic_tile_mesh(shim.asjax(x), shim.asjax(cx), shim.asjax(cy)))
  # The reference starts from np.zeros_like(cx); real JAX (x64 disabled) turns float64 NumPy
  # inputs into float32 at the jit boundary, the shim would keep them in float64 -- so the
  # offsets are handed over as float32 here.
  f32 = lambda a: a.astype(np.float32)
  out['cm2_opt'] = np.asarray(sr.optimize_coarse_mesh(f32(cx), f32(cy)))
  cfg = dict(dt=0.001, gamma=0.0, k0=0.0, k=0.1, stride=(1, 1), num_iters=100, max_iters=300,
             stop_v_max=0.0, dt_max=100)
  out['cm2_short_cfg'] = np.array(repr(cfg))
  out['cm2_short'] = np.asarray(sr.optimize_coarse_mesh(f32(cx), f32(cy),
                                                      cfg=mesh.IntegrationConfig(**cfg)))
  # 3-d tiles (LICONN): XYZ offsets, elastic_tile_mesh_3d
  cx3 = np.full((3, 1, ny, nx), np.nan)
  cy3 = np.full((3, 1, ny, nx), np.nan)
  cx3[0, 0, :, :-1] = -40 + rng.integers(-6, 7, (ny, nx - 1))
  cx3[1, 0, :, :-1] = rng.integers(-8, 9, (ny, nx - 1))
  cx3[2, 0, :, :-1] = rng.integers(-4, 5, (ny, nx - 1))
  cy3[0, 0, :-1, :] = rng.integers(-8, 9, (ny - 1, nx))
  cy3[1, 0, :-1, :] = -30 + rng.integers(-6, 7, (ny - 1, nx))
  cy3[2, 0, :-1, :] = rng.integers(-4, 5, (ny - 1, nx))
  out['cm3_cx'], out['cm3_cy'] = cx3, cy3
  x3 = (rng.standard_normal((3, 1, ny, nx)) * 5).astype(np.float32)
  out['cm3_x'] = x3
  out['cm3_force'] = np.asarray(sr.elastic_tile_mesh_3d(shim.asjax(x3), shim.asjax(cx3),
                                                        shim.asjax(cy3)))
  out['cm3_opt'] = np.asarray(sr.optimize_coarse_mesh(f32(cx3), f32(cy3),
                                                    mesh_fn=sr.elastic_tile_mesh_3d))
  # interpolate_missing_offsets: inf = "no acceptable estimate"
  conn = cx.copy()
  conn[:, 0, 1, 1] = np.inf
  conn[:, 0, 0, 0] = np.inf
  conn[:, 0, 2, 2] = np.inf
  conn[:, 0, 2, 1] = np.inf
  out['im_in'] = conn.copy()
  out['im_x'] = sr.interpolate_missing_offsets(conn.copy(), -1)
  out['im_y'] = sr.interpolate_missing_offsets(conn.copy(), -2)
  out['im_y_r2'] = sr.interpolate_missing_offsets(conn.copy(), -2, max_r=2)
  return out


def xcorr_numpy_branch_cases(ff):
  """flow_field.masked_xcorr(use_jax=False) (flow_field.py:36-156): the reference's OWN NumPy
  branch, i.e. real numpy.fft in float64 -- no stand-in for a JAX primitive is involved, so
  these vectors pin the correlation / Padfield arithmetic independently of the shim."""
  out = {}
  rng = np.random.default_rng(51)
  for tag, pshape, cshape in (('a', (2, 24, 30), (2, 24, 30)), ('b', (1, 17, 9), (1, 8, 13))):
    prev = rng.standard_normal(pshape).astype(np.float32) * 20
    curr = rng.standard_normal(cshape).astype(np.float32) * 20
    pm = rng.random(pshape) > 0.8
    cm = rng.random(cshape) > 0.75
    out[f'xn_{tag}_prev'], out[f'xn_{tag}_curr'] = prev, curr
    out[f'xn_{tag}_pm'], out[f'xn_{tag}_cm'] = pm, cm
    out[f'xn_{tag}_plain'] = np.asarray(ff.masked_xcorr(prev, curr, use_jax=False))
    out[f'xn_{tag}_masked'] = np.asarray(ff.masked_xcorr(prev, curr, pm, cm, use_jax=False))
  prev = rng.standard_normal((9, 11, 12)).astype(np.float32)
  curr = rng.standard_normal((4, 6, 5)).astype(np.float32)
  out['xn_3d_prev'], out['xn_3d_curr'] = prev, curr
  out['xn_3d_plain'] = np.asarray(ff.masked_xcorr(prev, curr, use_jax=False, dim=3))
  prev = (rng.standard_normal((2, 9, 11, 12)) * 10).astype(np.float32)
  curr = (rng.standard_normal((2, 9, 11, 12)) * 10).astype(np.float32)
  pm, cm = rng.random(prev.shape) > 0.8, rng.random(curr.shape) > 0.75
  out['xn_3dm_prev'], out['xn_3dm_curr'], out['xn_3dm_pm'], out['xn_3dm_cm'] = prev, curr, pm, cm
  out['xn_3dm_masked'] = np.asarray(ff.masked_xcorr(prev, curr, pm, cm, use_jax=False, dim=3))
  return out


def main():
  if 'xcorr_numpy' in sys.argv[1:]:
    ff = shim.load_reference('flow_field')
    path = os.path.join(HERE, 'xcorr_numpy_golden.npz')
    np.savez_compressed(path, **xcorr_numpy_branch_cases(ff))
    print('xcorr_numpy_golden.npz', os.path.getsize(path) // 1024, 'KiB')
    return
  if 'coarse' in sys.argv[1:]:
    sr = shim.load_reference('stitch_rigid')
    mesh = shim.load_reference('mesh')
    path = os.path.join(HERE, 'coarse_golden.npz')
    np.savez_compressed(path, **coarse_cases(sr, mesh))
    print('coarse_golden.npz', os.path.getsize(path) // 1024, 'KiB')
    return
  if 'flow3d' in sys.argv[1:]:
    se = shim.load_reference('stitch_elastic')
    path = os.path.join(HERE, 'flow3d_golden.npz')
    np.savez_compressed(path, **flow3d_cases(se))
    print('flow3d_golden.npz', os.path.getsize(path) // 1024, 'KiB')
    return
  if 'stitch' in sys.argv[1:] or len(sys.argv) == 1:
    mu = shim.load_reference('map_utils')
    se = shim.load_reference('stitch_elastic')
    np.savez_compressed(os.path.join(HERE, 'stitch_golden.npz'), **stitch_cases(mu, se))
    print('stitch_golden.npz', os.path.getsize(os.path.join(HERE, 'stitch_golden.npz')) // 1024,
          'KiB')
    if len(sys.argv) > 1:
      return
  mesh = shim.load_reference('mesh')
  ff = shim.load_reference('flow_field')
  np.savez_compressed(os.path.join(HERE, 'mesh_golden.npz'), **mesh_cases(mesh))
  np.savez_compressed(os.path.join(HERE, 'flow_golden.npz'), **flow_cases(ff))
  for f in ('mesh_golden.npz', 'flow_golden.npz'):
    print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, 'KiB')


if __name__ == '__main__':
  main()
