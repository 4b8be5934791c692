"""The rigid tile-grid relaxation (sofima_b200/csrc/tile_mesh_core.cuh) checked WITHOUT a GPU:
the header's __host__ __device__ functions -- the same ones the CUDA kernel of tile_mesh.cu
runs -- are compiled for the host (tests/host/tile_mesh_host.cpp, g++ -ffp-contract=off) and
must reproduce the reference's own run (tests/golden/coarse_golden.npz) bit for bit."""
import ast
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden', 'coarse_golden.npz')
F32P = ctypes.POINTER(ctypes.c_float)


@pytest.fixture(scope='module')
def host(tmp_path_factory):
  from sofima_b200 import _native
  so = tmp_path_factory.mktemp('tile_host') / 'libtile_mesh_host.so'
  subprocess.run(['g++', '-O2', '-ffp-contract=off', '-std=c++17', '-shared', '-fPIC', '-o',
                  str(so), os.path.join(ROOT, 'tests', 'host', 'tile_mesh_host.cpp')], check=True)
  lib = ctypes.CDLL(str(so))
  lib.tile_mesh_force_host.argtypes = [F32P, F32P, F32P] + [ctypes.c_int] * 4 + [F32P]
  lib.tile_mesh_chunk_host.argtypes = (
      [F32P] * 5 + [ctypes.c_int] * 4 + [ctypes.POINTER(_native.IntegrationConfigPod)] +
      [F32P, F32P, F32P, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_double), F32P])
  return lib, _native


def _p(a):
  return a.ctypes.data_as(F32P)


def _force(lib, x, cx, cy):
  x, cx, cy = (np.ascontiguousarray(a, np.float32) for a in (x, cx, cy))
  out = np.empty_like(x)
  lib.tile_mesh_force_host(_p(x), _p(cx), _p(cy), *x.shape, _p(out))
  return out


def _relax(lib, native, cx, cy, cfg):
  """Control flow of mesh.relax_mesh (mesh.py:524-608) around the host-built chunk."""
  from sofima_b200 import mesh
  cx, cy = np.ascontiguousarray(cx, np.float32), np.ascontiguousarray(cy, np.float32)
  x = np.zeros_like(cx)
  v, a = np.zeros_like(x), np.zeros_like(x)
  import dataclasses
  pod = mesh._config_pod(dataclasses.replace(cfg, stride=(1, 1)), mesh._INPLANE)
  dt, alpha, cap = (ctypes.c_float(float(s)) for s in (cfg.dt, cfg.alpha, cfg.start_cap))
  n_pos, e_kin, v_max = ctypes.c_int32(), ctypes.c_double(), ctypes.c_float()
  t = 0
  while t < cfg.max_iters:
    lib.tile_mesh_chunk_host(_p(x), _p(v), _p(a), _p(cx), _p(cy), *x.shape, ctypes.byref(pod),
                             ctypes.byref(dt), ctypes.byref(alpha), ctypes.byref(cap),
                             ctypes.byref(n_pos), ctypes.byref(e_kin), ctypes.byref(v_max))
    t += cfg.num_iters
    if np.float32(v_max.value) < np.float32(cfg.stop_v_max):
      if np.float32(cap.value) >= np.float32(cfg.final_cap):
        break
      cap = ctypes.c_float(float(min(np.float32(cap.value) * np.float32(cfg.cap_scale),
                                     np.float32(cfg.final_cap))))
  return x, t


def test_forces(host):
  lib, _ = host
  g = np.load(GOLDEN)
  np.testing.assert_array_equal(_force(lib, g['cm2_x'], g['cm2_cx'], g['cm2_cy']), g['cm2_force'])
  np.testing.assert_array_equal(_force(lib, g['cm3_x'], g['cm3_cx'], g['cm3_cy']), g['cm3_force'])


def test_relaxations(host):
  from sofima_b200 import mesh, stitch_rigid
  lib, native = host
  g = np.load(GOLDEN)
  short = mesh.IntegrationConfig(**ast.literal_eval(str(g['cm2_short_cfg'])))
  x, t = _relax(lib, native, g['cm2_cx'], g['cm2_cy'], short)
  assert t == 300
  np.testing.assert_array_equal(x, g['cm2_short'])
  default = stitch_rigid.default_coarse_mesh_config()
  x, t = _relax(lib, native, g['cm2_cx'], g['cm2_cy'], default)
  np.testing.assert_array_equal(x, g['cm2_opt'])
  x3, _ = _relax(lib, native, g['cm3_cx'], g['cm3_cy'], default)
  np.testing.assert_array_equal(x3, g['cm3_opt'])
  # plain velocity Verlet with damping (no FIRE): against the NumPy restatement
  from oracle import stitch_oracle as so
  damped = mesh.IntegrationConfig(dt=0.05, gamma=0.5, k0=0.0, k=0.1, stride=(1, 1),
                                  num_iters=50, max_iters=200, stop_v_max=0.0, fire=False)
  x, t = _relax(lib, native, g['cm2_cx'], g['cm2_cy'], damped)
  assert t == 200
  np.testing.assert_array_equal(x, so.optimize_coarse_mesh(g['cm2_cx'], g['cm2_cy'], damped))


# ---- the kernel source itself, one emulated thread block on the host ---------------------
EMU_SRC = os.path.join(ROOT, 'tests', 'host', 'tile_mesh_block_emu.cpp')
EMU_FLAGS = ['-O1', '-g', '-ffp-contract=off', '-std=c++20', '-pthread']


@pytest.fixture(scope='module')
def emu(tmp_path_factory):
  from sofima_b200 import _native
  so = tmp_path_factory.mktemp('tile_emu') / 'libtile_mesh_emu.so'
  subprocess.run(['g++'] + EMU_FLAGS + ['-shared', '-fPIC', '-o', str(so), EMU_SRC], check=True)
  lib = ctypes.CDLL(str(so))
  lib.tile_mesh_chunk_emu.argtypes = (
      [F32P] * 5 + [ctypes.c_int] * 4 + [ctypes.POINTER(_native.IntegrationConfigPod)] +
      [F32P, F32P, F32P, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_double), F32P])
  lib.tile_mesh_chunk_host = lib.tile_mesh_chunk_emu  # same signature: reuse _relax
  return lib, _native


def test_kernel_source_in_an_emulated_block(emu):
  """tile_chunk_kernel compiled for the host (256 OS threads, __syncthreads = std::barrier,
  __shared__ = static): the thread-block version of the chunk equals the reference's run."""
  from sofima_b200 import mesh, stitch_rigid
  from oracle import stitch_oracle as so
  lib, native = emu
  g = np.load(GOLDEN)
  short = mesh.IntegrationConfig(**ast.literal_eval(str(g['cm2_short_cfg'])))
  x, t = _relax(lib, native, g['cm2_cx'], g['cm2_cy'], short)
  assert t == 300
  np.testing.assert_array_equal(x, g['cm2_short'])
  x, _ = _relax(lib, native, g['cm3_cx'], g['cm3_cy'], stitch_rigid.default_coarse_mesh_config())
  np.testing.assert_array_equal(x, g['cm3_opt'])
  # more tiles than threads, plain integrator
  rng = np.random.default_rng(5)
  cx = np.full((2, 1, 20, 30), np.nan, np.float32)
  cy = cx.copy()
  cx[0, 0, :, :-1] = -40 + rng.integers(-6, 7, (20, 29))
  cx[1, 0, :, :-1] = rng.integers(-8, 9, (20, 29))
  cy[0, 0, :-1, :] = rng.integers(-8, 9, (19, 30))
  cy[1, 0, :-1, :] = -30 + rng.integers(-6, 7, (19, 30))
  damped = mesh.IntegrationConfig(dt=0.05, gamma=0.5, k0=0.0, k=0.1, stride=(1, 1),
                                  num_iters=50, max_iters=100, stop_v_max=0.0, fire=False)
  x, _ = _relax(lib, native, cx, cy, damped)
  np.testing.assert_array_equal(x, so.optimize_coarse_mesh(cx, cy, damped))
  x, _ = _relax(lib, native, cx, cy, short)
  np.testing.assert_array_equal(x, so.optimize_coarse_mesh(cx, cy, short))


def test_kernel_source_is_race_free_under_tsan(tmp_path):
  """The same emulation under ThreadSanitizer: a missing barrier or a node written by two
  threads would be reported as a data race (checked by hand: dropping every third barrier
  produces ten reports)."""
  exe = tmp_path / 'tile_mesh_emu_tsan'
  build = subprocess.run(['g++'] + EMU_FLAGS + ['-fsanitize=thread', '-DEMU_MAIN', '-o', str(exe),
                                                EMU_SRC], capture_output=True, text=True)
  if build.returncode != 0:
    pytest.skip('ThreadSanitizer runtime not available: ' + build.stderr[-200:])
  run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
  assert 'ThreadSanitizer' not in run.stderr, run.stderr[:2000]
  assert run.returncode == 0 and run.stdout.count('fire=') == 4
