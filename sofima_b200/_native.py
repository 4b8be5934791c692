"""Build + ctypes bindings of the sofima_b200 C-ABI library (include/sofima_b200.h).

The library is built in-tree (sofima_b200/_lib/libsofima_b200.so) by
`build()` with nvcc for sm_100a.  There is no CPU fallback: if the library is
missing or no B200 is present, every product entry point raises.
"""

from __future__ import annotations

import ctypes
import os
import subprocess
import threading
from typing import Sequence

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_DIR = os.path.join(HERE, '_lib')
BUILD_LIB_PATH = os.path.join(LIB_DIR, 'libsofima_b200.so')  # what build() writes
# SOFIMA_B200_LIB: load another build of the same C ABI (kernel A/B measurements,
# tools/ab_flow.py); build() never writes there.
LIB_PATH = os.environ.get('SOFIMA_B200_LIB') or BUILD_LIB_PATH
INCLUDE = os.path.join(os.path.dirname(HERE), 'include')

NVCC_ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
NVCC_COMMON = ['-lineinfo', '-O3', '-std=c++17', '-Xcompiler', '-fPIC']
# Per-file flags.  mesh.cu is bit-faithful to the fp32 reference arithmetic and
# must not contract a*b+c into FMA.
SOURCES = {
    'ctx.cu': [],
    'mesh.cu': ['-fmad=false'],
    'tile_mesh.cu': ['-fmad=false'],  # fp32 operation order of mesh.py / stitch_rigid.py
    'flow.cu': [],
    'flowfilt.cu': ['-fmad=false'],  # fp32 comparisons exactly as NumPy evaluates them
    'warp.cu': ['-fmad=false'],  # float64 arithmetic identical to SciPy's
    'warp_cv.cu': ['-fmad=false'],  # float arithmetic identical to SciPy's / OpenCV's
}

OK, EINVAL, EUNSUPPORTED, ECUDA, ENOMEM = 0, 1, 2, 3, 4


class NativeError(RuntimeError):
  pass


def _nvcc() -> str:
  for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
    if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
      return cand
  return 'nvcc'


def _stale(target: str, deps: Sequence[str]) -> bool:
  if not os.path.exists(target):
    return True
  t = os.path.getmtime(target)
  return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
  """Compiles csrc/*.cu for sm_100a into sofima_b200/_lib/libsofima_b200.so."""
  os.makedirs(LIB_DIR, exist_ok=True)
  headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC)
             if f.endswith(('.cuh', '.h'))]
  headers.append(os.path.join(INCLUDE, 'sofima_b200.h'))
  objs = []
  procs = []
  for src, extra in SOURCES.items():
    path = os.path.join(CSRC, src)
    if not os.path.exists(path):
      continue
    obj = os.path.join(LIB_DIR, src.replace('.cu', '.o'))
    objs.append(obj)
    if force or _stale(obj, [path] + headers):
      cmd = [_nvcc()] + NVCC_ARCH + NVCC_COMMON + extra + ['-c', path, '-o', obj]
      if verbose:
        print(' '.join(cmd))
      procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE,
                                          stderr=subprocess.STDOUT)))
  for cmd, proc in procs:
    out, _ = proc.communicate()
    if proc.returncode != 0:
      raise NativeError('nvcc failed: %s\n%s' % (' '.join(cmd), out.decode()))
  if force or procs or _stale(BUILD_LIB_PATH, objs):
    cmd = [_nvcc()] + NVCC_ARCH + ['-shared', '-o', BUILD_LIB_PATH] + objs
    if verbose:
      print(' '.join(cmd))
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if res.returncode != 0:
      raise NativeError('link failed: %s\n%s' % (' '.join(cmd), res.stdout.decode()))
  return BUILD_LIB_PATH


# --- ctypes mirror of include/sofima_b200.h ---------------------------------------


class IntegrationConfigPod(ctypes.Structure):
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


class MeshShape(ctypes.Structure):
  _fields_ = [('ncomp', ctypes.c_int32), ('nb', ctypes.c_int64),
              ('nz', ctypes.c_int64), ('ny', ctypes.c_int64),
              ('nx', ctypes.c_int64), ('batch_rank', ctypes.c_int32)]


class MeshState(ctypes.Structure):
  _fields_ = [('dt', ctypes.c_float), ('alpha', ctypes.c_float),
              ('cap', ctypes.c_float), ('gate', ctypes.c_float),
              ('n_pos', ctypes.c_int32), ('ticket', ctypes.c_uint32),
              ('mean_x', ctypes.c_float * 3), ('mean_v', ctypes.c_float * 3),
              ('power', ctypes.c_double), ('e_kin', ctypes.c_double),
              ('v_max', ctypes.c_float), ('pad', ctypes.c_int32)]


class StitchTargetPod(ctypes.Structure):
  _fields_ = [('fx', ctypes.c_void_p), ('fy', ctypes.c_void_p),
              ('nbors', ctypes.c_void_p), ('ndim', ctypes.c_int32),
              ('fx_shape', ctypes.c_int64 * 3), ('fy_shape', ctypes.c_int64 * 3),
              ('stride', ctypes.c_double * 3)]


class XcorrParams(ctypes.Structure):
  _fields_ = [
      ('ndim', ctypes.c_int32), ('img_dtype', ctypes.c_int32),
      ('pre_shape', ctypes.c_int64 * 3), ('post_shape', ctypes.c_int64 * 3),
      ('pre_mask_shape', ctypes.c_int64 * 3),
      ('post_mask_shape', ctypes.c_int64 * 3),
      ('pre_patch', ctypes.c_int32 * 3), ('post_patch', ctypes.c_int32 * 3),
      ('has_mean', ctypes.c_int32), ('mean', ctypes.c_float),
      ('min_distance', ctypes.c_int32), ('threshold_rel', ctypes.c_float),
      ('peak_radius', ctypes.c_int32 * 3),
  ]


_vp = ctypes.c_void_p
_PROTOS = {
    'sofima_abi_version': (ctypes.c_int, []),
    'sofima_ctx_create': (ctypes.c_int, [ctypes.c_int, _vp, ctypes.POINTER(_vp)]),
    'sofima_ctx_destroy': (ctypes.c_int, [_vp]),
    'sofima_ctx_set_stream': (ctypes.c_int, [_vp, _vp]),
    'sofima_last_error': (ctypes.c_char_p, [_vp]),
    'sofima_ctx_launch_count': (ctypes.c_int64, [_vp]),
    'sofima_ctx_set_timing': (ctypes.c_int, [_vp, ctypes.c_int]),
    'sofima_ctx_timing_report': (ctypes.c_int, [_vp, ctypes.c_char_p, ctypes.c_int64]),
    'sofima_ctx_trim': (ctypes.c_int, [_vp, ctypes.c_int64]),
    'sofima_mesh_force': (ctypes.c_int, [
        _vp, ctypes.c_int, _vp, ctypes.POINTER(MeshShape), ctypes.c_double,
        ctypes.POINTER(ctypes.c_double), ctypes.c_int, _vp]),
    'sofima_mesh_force_links': (ctypes.c_int, [
        _vp, ctypes.c_int, _vp, ctypes.POINTER(MeshShape), ctypes.c_double,
        ctypes.POINTER(ctypes.c_double), ctypes.c_int,
        ctypes.POINTER(ctypes.c_int32), ctypes.c_int, _vp]),
    'sofima_mesh_chunk': (ctypes.c_int, [
        _vp, ctypes.c_int, _vp, _vp, _vp, _vp, ctypes.POINTER(MeshShape),
        ctypes.POINTER(IntegrationConfigPod), ctypes.POINTER(ctypes.c_float),
        ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float),
        ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_double),
        ctypes.POINTER(ctypes.c_float)]),
    'sofima_mesh_chunk_stitch': (ctypes.c_int, [
        _vp, ctypes.c_int, _vp, _vp, _vp, ctypes.POINTER(StitchTargetPod), ctypes.POINTER(MeshShape),
        ctypes.POINTER(IntegrationConfigPod), ctypes.POINTER(ctypes.c_float),
        ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float),
        ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_double),
        ctypes.POINTER(ctypes.c_float)]),
    'sofima_tile_mesh_force': (ctypes.c_int, [
        _vp, _vp, _vp, _vp, ctypes.POINTER(MeshShape), _vp]),
    'sofima_tile_mesh_chunk': (ctypes.c_int, [
        _vp, _vp, _vp, _vp, _vp, _vp, ctypes.POINTER(MeshShape),
        ctypes.POINTER(IntegrationConfigPod), ctypes.POINTER(ctypes.c_float),
        ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float),
        ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_double),
        ctypes.POINTER(ctypes.c_float)]),
    'sofima_stitch_target_mesh': (ctypes.c_int, [
        _vp, _vp, ctypes.POINTER(MeshShape), ctypes.POINTER(StitchTargetPod), _vp]),
    'sofima_compose_maps': (ctypes.c_int, [
        _vp, ctypes.c_int, _vp, ctypes.POINTER(ctypes.c_int64),
        ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_double), _vp,
        ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64),
        ctypes.POINTER(ctypes.c_double), ctypes.c_int, _vp]),
    'sofima_mesh_chunk_async': (ctypes.c_int, [
        _vp, ctypes.c_int, _vp, _vp, _vp, _vp, ctypes.POINTER(MeshShape),
        ctypes.POINTER(IntegrationConfigPod), ctypes.c_float, ctypes.c_float,
        ctypes.c_float, _vp]),
    'sofima_shard_create': (ctypes.c_int, [
        _vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(MeshShape), ctypes.POINTER(_vp)]),
    'sofima_shard_export': (ctypes.c_int, [_vp, ctypes.c_char_p]),
    'sofima_shard_connect': (ctypes.c_int, [_vp, ctypes.c_char_p]),
    'sofima_shard_set_state': (ctypes.c_int, [_vp, _vp, _vp, _vp]),
    'sofima_shard_chunk': (ctypes.c_int, [
        _vp, ctypes.POINTER(IntegrationConfigPod), ctypes.c_float, ctypes.c_float,
        ctypes.c_float, ctypes.c_int64, ctypes.POINTER(MeshState)]),
    'sofima_shard_get_state': (ctypes.c_int, [_vp, _vp, _vp, _vp]),
    'sofima_shard_destroy': (ctypes.c_int, [_vp]),
    'sofima_xcorr_peaks': (ctypes.c_int, [
        _vp, ctypes.POINTER(XcorrParams), _vp, _vp, _vp, _vp, _vp, _vp,
        ctypes.c_int64, _vp]),
    'sofima_warp_image': (ctypes.c_int, [
        _vp, ctypes.c_int, _vp, ctypes.c_int, ctypes.POINTER(ctypes.c_int64), _vp,
        ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_double),
        ctypes.POINTER(ctypes.c_double), ctypes.c_int, _vp, ctypes.POINTER(ctypes.c_int64)]),
    'sofima_warp_subvolume': (ctypes.c_int, [
        _vp, _vp, ctypes.c_int, ctypes.POINTER(ctypes.c_int64), _vp, ctypes.c_int, _vp, _vp,
        ctypes.c_int64, ctypes.c_int64, _vp, ctypes.c_int, _vp, ctypes.c_int64,
        ctypes.c_int64]),
    'sofima_xcorr_rowcache': (ctypes.c_int, [
        _vp, ctypes.POINTER(XcorrParams), _vp, _vp, ctypes.POINTER(ctypes.c_int32),
        ctypes.c_int32, ctypes.POINTER(ctypes.c_int32), ctypes.c_int32]),
    'sofima_xcorr_images': (ctypes.c_int, [
        _vp, ctypes.POINTER(XcorrParams), _vp, _vp, _vp, _vp, _vp, _vp,
        ctypes.c_int64, _vp]),
    'sofima_clean_flow': (ctypes.c_int, [
        _vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int64), ctypes.c_float,
        ctypes.c_float, ctypes.c_float, ctypes.c_float, _vp]),
    'sofima_reconcile_flows': (ctypes.c_int, [
        _vp, _vp, ctypes.POINTER(_vp), ctypes.c_int, ctypes.c_int,
        ctypes.POINTER(ctypes.c_int64), ctypes.c_float, ctypes.c_float, ctypes.c_int,
        ctypes.c_float]),
    'sofima_mask_irregular': (ctypes.c_int, [
        _vp, _vp, ctypes.c_int64, ctypes.c_int64, ctypes.POINTER(ctypes.c_double),
        ctypes.c_double, ctypes.c_double, ctypes.c_int, _vp]),
    'sofima_batched_peaks': (ctypes.c_int, [
        _vp, ctypes.c_int, _vp, ctypes.POINTER(ctypes.c_int64), ctypes.c_int64,
        ctypes.POINTER(ctypes.c_int32), ctypes.c_int, ctypes.c_float,
        ctypes.POINTER(ctypes.c_int32), _vp]),
}

EXPORTED_SYMBOLS = tuple(_PROTOS)

_lib = None
_lib_lock = threading.Lock()


def lib() -> ctypes.CDLL:
  """Loads the C-ABI library; raises if it has not been built."""
  global _lib
  with _lib_lock:
    if _lib is None:
      if not os.path.exists(LIB_PATH):
        raise NativeError(
            f'{LIB_PATH} not found: build it with '
            '`python -c "import __graft_entry__ as g; g.build()"`. '
            'sofima_b200 has no CPU fallback.')
      handle = ctypes.CDLL(LIB_PATH)
      for name, (res, args) in _PROTOS.items():
        fn = getattr(handle, name)
        fn.restype = res
        fn.argtypes = args
      if handle.sofima_abi_version() != 1:
        raise NativeError('ABI version mismatch; rebuild the library')
      _lib = handle
  return _lib


def check(ctx_handle, rc: int):
  if rc == OK:
    return
  msg = lib().sofima_last_error(ctx_handle)
  msg = msg.decode() if msg else f'error {rc}'
  if rc == EINVAL:
    raise ValueError(msg)
  if rc == EUNSUPPORTED:
    raise NotImplementedError(msg)
  if rc == ENOMEM:
    raise MemoryError(msg)
  raise NativeError(msg)


class Context:
  """A sofima_ctx bound to one CUDA device and the torch stream current at use.

  include/sofima_b200.h: "one context per host thread" -- a sofima_ctx owns its stream
  binding and its scratch buffers and is not re-entrant (ctypes releases the GIL during
  the native calls).  `get()` therefore hands every Python thread its own context per
  device; threads that share a GPU hold separate scratch memory.
  """

  _per_device: dict[tuple[int, int], 'Context'] = {}
  _lock = threading.Lock()

  def __init__(self, device: int):
    import torch  # plumbing only: device memory + streams
    if not torch.cuda.is_available():
      raise NativeError(
          'No CUDA device: sofima_b200 runs on B200 only (no CPU fallback).')
    self.device = int(device)
    self._torch = torch
    h = _vp()
    with torch.cuda.device(self.device):
      torch.cuda.current_stream()  # make sure the primary context exists
      check(None, lib().sofima_ctx_create(self.device, None, ctypes.byref(h)))
    self.handle = h

  @classmethod
  def get(cls, device: int | None = None) -> 'Context':
    import torch
    if device is None:
      if not torch.cuda.is_available():
        raise NativeError(
            'No CUDA device: sofima_b200 runs on B200 only (no CPU fallback).')
      device = torch.cuda.current_device()
    key = (int(device), threading.get_ident())
    with cls._lock:
      ctx = cls._per_device.get(key)
      if ctx is None:
        ctx = cls._per_device[key] = Context(device)
    return ctx

  def trim(self, keep_bytes: int = 0):
    """Frees the context's grow-only scratch buffers larger than `keep_bytes` (the flow
    path's row-spectra cache and spectra scratch can hold several GB after one large
    call).  Synchronises the stream."""
    check(self.handle, lib().sofima_ctx_trim(self.handle, int(keep_bytes)))

  def bind_stream(self):
    stream = self._torch.cuda.current_stream(self.device).cuda_stream
    check(self.handle, lib().sofima_ctx_set_stream(self.handle, _vp(stream)))

  def set_timing(self, on: bool):
    check(self.handle, lib().sofima_ctx_set_timing(self.handle, int(on)))

  def timing_report(self) -> dict:
    import json
    buf = ctypes.create_string_buffer(1 << 16)
    check(self.handle, lib().sofima_ctx_timing_report(self.handle, buf, len(buf)))
    return json.loads(buf.value.decode())

  @property
  def launch_count(self) -> int:
    return int(lib().sofima_ctx_launch_count(self.handle))

  def __del__(self):
    try:
      if self.handle:
        lib().sofima_ctx_destroy(self.handle)
    except Exception:  # pylint: disable=broad-except
      pass
