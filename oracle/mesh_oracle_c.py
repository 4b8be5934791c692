"""ctypes wrapper of oracle/mesh_oracle.c (TEST ORACLE / CPU baseline, multi-threaded).

Follows /root/reference/mesh.py:42-169 (inplane_force) and :371-521
(velocity_verlet); validated bit-for-bit against oracle/mesh_oracle.py.
Test infrastructure only: never imported from `sofima_b200/`.
"""

from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, '_build', 'libmesh_oracle.so')


class _Config(ctypes.Structure):
  _fields_ = [
      ('dt', ctypes.c_double), ('gamma', ctypes.c_double),
      ('k0', ctypes.c_double), ('k', ctypes.c_double),
      ('stride', ctypes.c_double * 3),
      ('num_iters', ctypes.c_int32), ('fire', ctypes.c_int32),
      ('f_alpha', ctypes.c_double), ('f_inc', ctypes.c_double),
      ('f_dec', ctypes.c_double), ('alpha', ctypes.c_double),
      ('n_min', ctypes.c_int32), ('dt_max', ctypes.c_double),
      ('start_cap', ctypes.c_double), ('final_cap', ctypes.c_double),
      ('cap_scale', ctypes.c_double),
      ('cap_upscale_every', ctypes.c_int32),
      ('prefer_orig_order', ctypes.c_int32),
      ('remove_drift', ctypes.c_int32),
  ]


def build(force: bool = False) -> str:
  src = os.path.join(HERE, 'mesh_oracle.c')
  if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
    subprocess.run(['make', '-C', HERE, '-B', '_build/libmesh_oracle.so'], check=True,
                   stdout=subprocess.DEVNULL)
  return LIB


_lib = None


def _load():
  global _lib
  if _lib is None:
    build()
    _lib = ctypes.CDLL(LIB)
    _lib.oracle_mesh_chunk.restype = ctypes.c_int
    _lib.oracle_num_threads.restype = ctypes.c_int
  return _lib


def num_threads() -> int:
  return int(_load().oracle_num_threads())


def _pod(config) -> _Config:
  c = _Config()
  for name in ('dt', 'gamma', 'k0', 'k', 'f_alpha', 'f_inc', 'f_dec', 'alpha',
               'dt_max', 'start_cap', 'final_cap', 'cap_scale'):
    setattr(c, name, float(getattr(config, name)))
  c.stride[0], c.stride[1], c.stride[2] = float(config.stride[0]), float(config.stride[1]), 0.0
  c.num_iters, c.fire, c.n_min = int(config.num_iters), int(config.fire), int(config.n_min)
  c.cap_upscale_every = int(config.cap_upscale_every)
  c.prefer_orig_order = int(config.prefer_orig_order)
  c.remove_drift = int(config.remove_drift)
  return c


def velocity_verlet(x, v, prev, config, force_cap, fire_dt=None, fire_alpha=None):
  """One chunk; same results as mesh_oracle.velocity_verlet (2-d inplane only).

  Returns (x, v, a, dt, alpha, n_pos, cap, e_kin, v_max).
  """
  lib = _load()
  x = np.array(x, dtype=np.float32, order='C')
  v = np.array(v, dtype=np.float32, order='C')
  a = np.empty_like(x)
  assert x.ndim == 4 and x.shape[0] == 2
  pv = None
  if prev is not None:
    pv = np.ascontiguousarray(prev, dtype=np.float32)
  dt = ctypes.c_float(config.dt if fire_dt is None else fire_dt)
  alpha = ctypes.c_float(config.alpha if fire_alpha is None else fire_alpha)
  cap = ctypes.c_float(force_cap)
  n_pos, e_kin, v_max = ctypes.c_int32(0), ctypes.c_double(0), ctypes.c_float(0)
  pod = _pod(config)
  fp = ctypes.POINTER(ctypes.c_float)
  rc = lib.oracle_mesh_chunk(
      x.ctypes.data_as(fp), v.ctypes.data_as(fp), a.ctypes.data_as(fp),
      None if pv is None else pv.ctypes.data_as(fp), ctypes.c_long(x.shape[1]),
      ctypes.c_long(x.shape[2]), ctypes.c_long(x.shape[3]), ctypes.byref(pod),
      ctypes.byref(dt), ctypes.byref(alpha), ctypes.byref(cap), ctypes.byref(n_pos),
      ctypes.byref(e_kin), ctypes.byref(v_max))
  if rc:
    raise MemoryError('oracle_mesh_chunk failed')
  return (x, v, a, np.float32(dt.value), np.float32(alpha.value), int(n_pos.value),
          np.float32(cap.value), float(e_kin.value), np.float32(v_max.value))


def relax_mesh(x, prev, config):
  """mesh.relax_mesh (mesh.py:524-608) on the C integrator."""
  t = 0
  x = np.array(x, dtype=np.float32)
  v = np.zeros_like(x)
  dt, alpha, cap = config.dt, config.alpha, config.start_cap
  e_kin = []
  while t < config.max_iters:
    x, v, _, dt_n, alpha_n, _, cap_n, ek, v_max = velocity_verlet(
        x, v, prev, config, cap, dt, alpha)
    t += config.num_iters
    e_kin.append(ek)
    if config.fire:
      dt, alpha, cap = dt_n, alpha_n, cap_n
    if v_max < np.float32(config.stop_v_max):
      if np.float32(cap) >= np.float32(config.final_cap):
        break
      cap = min(np.float32(cap) * np.float32(config.cap_scale),
                np.float32(config.final_cap))
  return x, e_kin, t
