"""NumPy fp32 restatement of the reference spring-mesh solver (TEST ORACLE).

Follows /root/reference/mesh.py:
  * link forces            inplane_force      mesh.py:42-169
                           elastic_mesh_3d    mesh.py:192-279
  * integrator             velocity_verlet    mesh.py:371-521
  * host convergence loop  relax_mesh         mesh.py:524-608

The reference runs under JAX with x64 disabled, i.e. every array op is fp32 and
every Python/NumPy scalar constant is rounded to fp32 before it meets an array.
This file reproduces that: constants are rounded exactly where JAX would round
them and every elementwise operation is carried out in fp32 in the reference's
association order, so that a bit-faithful CUDA kernel can be compared with it.
Deliberate, documented deviation: global reductions (`vdot`, `mean`, `sum`) are
accumulated in fp64 -- XLA's fp32 reduction order is unspecified, fp64 is the
neutral choice (only the SIGN of `power` feeds back into the trajectory).

Test infrastructure only: never imported from `sofima_b200/`.
"""

from __future__ import annotations

import math
from typing import Callable, Sequence

import numpy as np

F32 = np.float32

# xyz link directions of the 3-d mesh (same set as mesh.py:172-189).
LINKS_3D = (
    (1, 0, 0), (0, 1, 0), (0, 0, 1),
    (1, 1, 0), (-1, 1, 0), (1, 0, 1), (-1, 0, 1), (0, 1, 1), (0, -1, 1),
    (1, 1, 1), (1, 1, -1), (1, -1, 1), (-1, 1, 1),
)


def _axis_slices(step: int):
  """(slice of the 'to' node, slice of the 'from' node) along one axis."""
  if step == 1:
    return slice(1, None), slice(None, -1)
  if step == -1:
    return slice(None, -1), slice(1, None)
  if step == 0:
    return slice(None), slice(None)
  raise ValueError('Only |v| <= 1 values supported within links.')


def _zero_nonfinite(f: np.ndarray) -> np.ndarray:
  # jnp.nan_to_num(f, posinf=0.0, neginf=0.0): nan, +inf, -inf -> 0.
  return np.where(np.isfinite(f), f, F32(0.0)).astype(F32)


def _link_force(x, direction, l0_vec, l0_len, k_eff, prefer_orig_order):
  """Force field of one link family.

  Returns (f, to_sel, from_sel): `f` lives on the links; +f acts on the node
  selected by `to_sel`, -f on the node selected by `from_sel`.
  """
  ncomp = x.shape[0]
  nsp = len(direction)  # number of spatial axes that carry links
  lead = [slice(None)] * (x.ndim - nsp)
  to_sel, from_sel = list(lead), list(lead)
  for step in direction[::-1]:  # spatial axes are stored z, y, x
    a, b = _axis_slices(step)
    to_sel.append(a)
    from_sel.append(b)
  to_sel, from_sel = tuple(to_sel), tuple(from_sel)

  l0_b = np.asarray(l0_vec, dtype=F32).reshape([ncomp] + [1] * (x.ndim - 1))
  dx = (x[to_sel] - x[from_sel]) + l0_b
  sq = dx[0] * dx[0]
  for c in range(1, ncomp):
    sq = sq + dx[c] * dx[c]
  with np.errstate(all='ignore'):
    length = np.sqrt(sq)
    if prefer_orig_order:
      factor = np.ones_like(dx)
      for c in range(len(direction)):
        if direction[c] != 0:
          factor[c] = F32(direction[c]) * np.sign(dx[c])
      f = (F32(-k_eff) * (F32(1.0) - F32(l0_len) * factor / length)) * dx
    else:
      f = (F32(-k_eff) * (F32(1.0) - F32(l0_len) / length)) * dx
  return _zero_nonfinite(f), to_sel, from_sel


def _placed(f, sel, shape):
  out = np.zeros(shape, dtype=F32)
  out[sel] = f
  return out


def inplane_force(x, k, stride, prefer_orig_order=False):
  """2-d in-plane Hookean forces; mesh.py:42-169.  x: [2, z, y, x]."""
  if len(stride) != 2:
    raise ValueError('stride must be 2D.')
  x = np.asarray(x, dtype=F32)
  sx, sy = float(stride[0]), float(stride[1])
  l0_diag = float(np.linalg.norm(np.array(stride, dtype=np.float64)))
  k_ax = F32(k)
  k_diag = F32(k) / F32(np.sqrt(F32(2.0)))  # mesh.py:137, fp32 division
  families = (
      ((1, 0), (sx, 0.0), sx, k_ax),          # -   mesh.py:107-119
      ((0, 1), (0.0, sy), sy, k_ax),          # |   mesh.py:122-134
      ((1, 1), (sx, sy), l0_diag, k_diag),    # \   mesh.py:140-152
      ((-1, 1), (-sx, sy), l0_diag, k_diag),  # /   mesh.py:155-167
  )
  plus, minus = [], []
  for direction, l0_vec, l0_len, k_eff in families:
    f, to_sel, from_sel = _link_force(x, direction, l0_vec, l0_len, k_eff,
                                      prefer_orig_order)
    plus.append(_placed(f, to_sel, x.shape))
    minus.append(_placed(f, from_sel, x.shape))
  # mesh.py:169 -- f1p + f2p + f3p + f4p - f1n - f2n - f3n - f4n, left to right.
  total = plus[0]
  for p in plus[1:]:
    total = total + p
  for m in minus:
    total = total - m
  return total


def elastic_mesh_3d(x, k, stride, prefer_orig_order=False, links=LINKS_3D):
  """3-d Hookean forces; mesh.py:192-279.  x: [3, [batch..], z, y, x]."""
  x = np.asarray(x, dtype=F32)
  assert x.shape[0] == 3
  if not isinstance(stride, (tuple, list, np.ndarray)):
    stride = (stride,) * 3
  stride = np.array(stride, dtype=np.float64)
  total = None
  for direction in links:
    l0_vec = np.array(stride * direction, dtype=F32)         # mesh.py:249
    l0_len = F32(np.sqrt((l0_vec[0] * l0_vec[0] + l0_vec[1] * l0_vec[1])
                         + l0_vec[2] * l0_vec[2]))           # mesh.py:253
    k_eff = F32(float(k) * stride[0] / np.float64(l0_len))   # mesh.py:259
    f, to_sel, from_sel = _link_force(x, direction, l0_vec, l0_len, k_eff,
                                      prefer_orig_order)
    fp = _placed(f, to_sel, x.shape)
    total = fp if total is None else total + fp              # mesh.py:271-275
    total = total - _placed(f, from_sel, x.shape)            # mesh.py:276-277
  return total


def _nan_to_num_default(d: np.ndarray) -> np.ndarray:
  # jnp.nan_to_num defaults: nan -> 0, +/-inf -> +/-FLT_MAX (mesh.py:433).
  fmax = np.finfo(F32).max
  d = np.where(np.isnan(d), F32(0.0), d)
  d = np.where(d == np.inf, fmax, d)
  d = np.where(d == -np.inf, -fmax, d)
  return d.astype(F32)


def total_force(x, prev, cap, config, mesh_force, prev_fn=None):
  """`_force` of mesh.py:427-434."""
  a = mesh_force(x, config.k, config.stride, config.prefer_orig_order)
  if prev_fn is not None:
    prev = prev_fn(x)
  if prev is not None:
    with np.errstate(all='ignore'):
      pull = F32(-config.k0) * _nan_to_num_default(x - np.asarray(prev, F32))
    a = a + np.clip(pull, -F32(cap), F32(cap)).astype(F32)
  return a.astype(F32)


def _norm0(a):
  sq = a[0] * a[0]
  for c in range(1, a.shape[0]):
    sq = sq + a[c] * a[c]
  return np.sqrt(sq)[np.newaxis]


def power_f64(a, v) -> float:
  """vdot(a, v) of mesh.py:455 with exact products and fp64 accumulation."""
  return float(np.sum(a.astype(np.float64) * v.astype(np.float64)))


def velocity_verlet(x, v, prev, config, force_cap, fire_dt=None,
                    fire_alpha=None, mesh_force=inplane_force, prev_fn=None,
                    trace: list | None = None):
  """One chunk of `config.num_iters` integration steps; mesh.py:371-521.

  Returns (x, v, a) or, with FIRE, (x, v, a, dt, alpha, n_pos, cap).  Inputs
  are not modified.  If `trace` is a list, one `(power, dt, alpha, n_pos, cap)`
  tuple is appended per FIRE step (test instrumentation).
  """
  x = np.array(x, dtype=F32)
  v = np.array(v, dtype=F32)
  gamma = config.gamma
  with np.errstate(all='ignore'):
    a = total_force(x, prev, force_cap, config, mesh_force, prev_fn)

    if not config.fire:
      # Python-float constants folded in double, rounded once (mesh.py:439-445
      # with dt = config.dt a Python float).
      dt = float(config.dt)
      c_dt = F32(dt)
      c_hdt2 = F32(0.5 * dt**2)
      fact0 = F32(1.0 / (1.0 + 0.5 * dt * gamma))
      fact1 = F32(1.0 - 0.5 * dt * gamma)
      c_hdt = F32(0.5 * dt)
      for _ in range(config.num_iters):
        x = x + (c_dt * v + c_hdt2 * a)
        a_prev = a
        a = total_force(x, prev, force_cap, config, mesh_force, prev_fn)
        v = fact0 * (v * fact1 + c_hdt * (a_prev + a))
      return x, v, a

    dt = F32(config.dt if fire_dt is None else fire_dt)
    alpha = F32(config.alpha if fire_alpha is None else fire_alpha)
    cap = F32(force_cap)
    n_pos = 0
    dt_ceiling = F32(config.dt_max * config.dt)
    for _ in range(config.num_iters):
      # vv_step, mesh.py:436-446, all scalars fp32.
      hdt2 = F32(0.5) * (dt * dt)
      x = x + (dt * v + hdt2 * a)
      a_prev = a
      a = total_force(x, prev, cap, config, mesh_force, prev_fn)
      hdtg = (F32(0.5) * dt) * F32(gamma)
      fact0 = F32(1.0) / (F32(1.0) + hdtg)
      fact1 = F32(1.0) - hdtg
      v = fact0 * (v * fact1 + (F32(0.5) * dt) * (a_prev + a))

      # fire_step, mesh.py:448-499.
      a_norm = _norm0(a) + F32(1e-6)
      v_norm = _norm0(v)
      power = power_f64(a, v)
      v = v + alpha * (a / a_norm * v_norm - v)
      pos = power >= 0
      n_pos = n_pos + 1 if pos else 0
      if pos:
        if n_pos > config.n_min:
          dt = min(dt * F32(config.f_inc), dt_ceiling)
          alpha = alpha * F32(config.f_alpha)
        if n_pos > 0 and n_pos % config.cap_upscale_every == 0:
          cap = F32(config.cap_scale) * cap
      else:
        dt = dt * F32(config.f_dec)
        alpha = F32(config.alpha)
      cap = min(cap, F32(config.final_cap))
      v = v * F32(1.0 if pos else 0.0)
      if config.remove_drift:
        # mesh.py:496-497 averages over axes (1, 2, 3) LITERALLY: all of (z, y, x) for
        # [N, z, y, x] meshes, but only (batch, z, y) -- one mean per x column -- for
        # the 5-d [3, tiles, z, y, x] meshes of 3-d stitching.
        axes = (1, 2, 3)
        x = x - np.mean(x, axis=axes, keepdims=True, dtype=np.float64).astype(F32)
        v = v - np.mean(v, axis=axes, keepdims=True, dtype=np.float64).astype(F32)
      if trace is not None:
        trace.append((power, float(dt), float(alpha), n_pos, float(cap)))
  return x, v, a, F32(dt), F32(alpha), n_pos, F32(cap)


def chunk_stats(v):
  """(e_kin, v_max) as relax_mesh computes them; mesh.py:584-586."""
  v_mag = _norm0(np.asarray(v, F32))[0]
  e_kin = float(np.sum((v_mag * v_mag).astype(np.float64)))
  with np.errstate(all='ignore'):
    v_max = F32(np.max(v_mag)) if v_mag.size else F32(0)
  return e_kin, v_max


def relax_mesh(x, prev, config, mesh_force=inplane_force, prev_fn=None,
               log: list | None = None):
  """Host convergence loop; mesh.py:524-608.  Returns (x, e_kin list, t)."""
  t = 0
  x = np.array(x, dtype=F32)
  v = np.zeros_like(x)
  dt, alpha, cap = config.dt, config.alpha, config.start_cap
  e_kin = []

  if config.start_cap != config.final_cap:
    if not config.fire:
      raise NotImplementedError(
          'Adaptive force capping is only supported with FIRE.')
    if config.cap_scale <= 1:
      raise ValueError(
          'The scaling factor for the force cap has to be larger '
          'than 1 when the initial and final cap are different.')
  if prev is not None and prev_fn is not None:
    raise ValueError('Only one of: "prev" and "prev_fn" can be specified.')

  while t < config.max_iters:
    state = velocity_verlet(x, v, prev, config, force_cap=cap, fire_dt=dt,
                            fire_alpha=alpha, mesh_force=mesh_force,
                            prev_fn=prev_fn)
    t += config.num_iters
    x, v = state[:2]
    ek, v_max = chunk_stats(v)
    e_kin.append(ek)
    n_pos = None
    if config.fire:
      dt, alpha, n_pos, cap = state[-4:]
    if log is not None:
      log.append(dict(t=t, dt=float(dt), alpha=float(alpha), n_pos=n_pos,
                      cap=float(cap), v_max=float(v_max), e_kin=ek))
    if v_max < F32(config.stop_v_max):
      if F32(cap) >= F32(config.final_cap):
        break
      cap = min(F32(cap) * F32(config.cap_scale), F32(config.final_cap))
  return x, e_kin, t
