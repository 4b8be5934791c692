"""NumPy stand-ins for the handful of JAX / connectomics / dataclasses_json names
that /root/reference/{mesh,flow_field}.py import, so that the reference's OWN
source files can be executed in this image (JAX is not installable offline).

Used ONLY by tests/golden/make_golden.py to generate the committed fixtures.
It emulates JAX's default numerics: x64 disabled, i.e. float64 operands are
demoted to float32 whenever they meet an array, Python scalars are weakly typed,
loop-carried Python scalars become fp32 / int32 arrays.  The semantics of the
JAX-only primitives (dynamic_slice clamping, zero-padded SAME patches,
first-index argmax, NumPy-style .at[].set) are restated from the JAX docs.
"""

from __future__ import annotations

import dataclasses
import sys
import types

import numpy as np

_DEMOTE = {np.dtype(np.float64): np.float32, np.dtype(np.complex128): np.complex64,
           np.dtype(np.int64): np.int32}


def _demote(a):
  if isinstance(a, (np.ndarray, np.generic)):
    tgt = _DEMOTE.get(a.dtype)
    if tgt is not None and a.dtype.kind in 'fc':
      a = np.asarray(a).astype(tgt)
  return a


class _At:

  def __init__(self, arr):
    self._arr = arr

  def __getitem__(self, idx):
    arr = self._arr

    class _Setter:

      def set(self, value):
        out = np.array(arr, copy=True)
        out[idx] = value
        return out.view(JArray)

    return _Setter()


class JArray(np.ndarray):
  """ndarray that never promotes to float64 (JAX with x64 disabled)."""

  __array_priority__ = 1000

  def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
    conv = []
    for i in inputs:
      if isinstance(i, JArray):
        conv.append(i.view(np.ndarray))
      elif isinstance(i, (np.ndarray, np.generic)):
        conv.append(_demote(np.asarray(i)))
      else:
        conv.append(i)  # Python scalars stay weakly typed
    if out is not None:
      kwargs['out'] = tuple(
          o.view(np.ndarray) if isinstance(o, JArray) else o for o in out)
    res = getattr(ufunc, method)(*conv, **kwargs)
    if out is not None:
      return out[0] if len(out) == 1 else out
    return _wrap(res)

  @property
  def at(self):
    return _At(self)


def _wrap(res):
  if isinstance(res, tuple):
    return tuple(_wrap(r) for r in res)
  if isinstance(res, (np.ndarray, np.generic)):
    return np.asarray(_demote(res)).view(JArray)
  return res


def _unwrap(a):
  if isinstance(a, JArray):
    return a.view(np.ndarray)
  if isinstance(a, (list, tuple)):
    return type(a)(_unwrap(x) for x in a)
  return a


def _lift(fn):
  def wrapped(*args, **kwargs):
    args = [_unwrap(a) for a in args]
    kwargs = {k: _unwrap(v) for k, v in kwargs.items()}
    return _wrap(fn(*args, **kwargs))
  wrapped.__name__ = getattr(fn, '__name__', 'fn')
  return wrapped


def asjax(a):
  """Array as JAX would see it after device_put: fp32 / int32 / bool / uint8."""
  a = np.asarray(a)
  tgt = _DEMOTE.get(a.dtype)
  if tgt is not None:
    a = a.astype(tgt)
  return np.array(a, copy=True).view(JArray)


# --- jax.numpy -----------------------------------------------------------------


def _mean(a, axis=None, keepdims=False, **kw):
  # jnp.mean: fp32 sum / fp32 count (sum accumulated in fp64 here; exact for
  # the integer-valued images of the fixtures).
  a = np.asarray(_unwrap(a))
  if a.dtype.kind in 'fc' and a.dtype.itemsize >= 4 and a.dtype.kind == 'c':
    return _wrap(np.mean(a, axis=axis, keepdims=keepdims))
  tot = np.sum(a, axis=axis, keepdims=keepdims, dtype=np.float64)
  cnt = a.size / max(tot.size, 1)
  return _wrap(np.asarray(tot).astype(np.float32) / np.float32(cnt))


def _nanmean(a, axis=None, keepdims=False):
  a = np.asarray(_unwrap(a), dtype=np.float32)
  valid = ~np.isnan(a)
  tot = np.sum(np.where(valid, a, 0.0), axis=axis, keepdims=keepdims,
               dtype=np.float64)
  cnt = np.sum(valid, axis=axis, keepdims=keepdims)
  with np.errstate(all='ignore'):
    return _wrap(np.asarray(tot).astype(np.float32) / cnt.astype(np.float32))


def _array(obj, dtype=None, **kw):
  if isinstance(obj, (list, tuple)):
    obj = [_unwrap(o) for o in obj]
  out = np.array(_unwrap(obj), dtype=dtype)
  return asjax(out)


def _clip(a, a_min=None, a_max=None, *, min=None, max=None):  # pylint: disable=redefined-builtin
  lo = a_min if min is None else min
  hi = a_max if max is None else max
  return _wrap(np.clip(_unwrap(a), _unwrap(lo), _unwrap(hi)))


def _nan_to_num(x, copy=True, nan=0.0, posinf=None, neginf=None):
  return _wrap(np.nan_to_num(_unwrap(x), copy=True, nan=nan, posinf=posinf,
                             neginf=neginf))


def _build_jnp():
  jnp = types.ModuleType('jax.numpy')
  for name in ('sign', 'ones_like', 'zeros_like', 'pad', 'sqrt', 'vdot', 'where',
               'minimum', 'maximum', 'sum', 'max', 'min', 'argmax',
               'take_along_axis', 'isinf', 'isnan', 'round', 'fmax', 'square',
               'logical_not', 'ones', 'zeros', 'abs', 'cumsum', 'stack',
               'concatenate', 'floor', 'ceil', 'arange', 'meshgrid', 'full', 'transpose'):
    setattr(jnp, name, _lift(getattr(np, name)))
  jnp.array = _array
  jnp.asarray = _array
  jnp.full = lambda shape, fill, dtype=None: _wrap(np.full(shape, fill, dtype=dtype))

  class _R:

    def __getitem__(self, items):
      return _wrap(np.r_[tuple(np.asarray(_unwrap(i)) for i in items)])

  jnp.r_ = _R()
  jnp.mean = _mean
  jnp.nanmean = _nanmean
  jnp.clip = _clip
  jnp.nan_to_num = _nan_to_num
  jnp.unravel_index = lambda idx, shape: tuple(
      np.int32(i) for i in np.unravel_index(int(idx), shape))
  jnp.finfo = np.finfo
  jnp.float32, jnp.int32, jnp.uint32 = np.float32, np.int32, np.uint32
  jnp.inf, jnp.nan, jnp.newaxis = np.inf, np.nan, np.newaxis
  jnp.ndarray = np.ndarray
  linalg = types.ModuleType('jax.numpy.linalg')
  linalg.norm = _lift(np.linalg.norm)
  jnp.linalg = linalg
  fft = types.ModuleType('jax.numpy.fft')
  fft.rfftn = _lift(np.fft.rfftn)
  fft.irfftn = _lift(np.fft.irfftn)
  jnp.fft = fft
  return jnp


# --- jax / jax.lax ---------------------------------------------------------------


def _jit(fn=None, **kw):
  if fn is None:
    return lambda f: f
  return fn


def _carry(v):
  if isinstance(v, bool):
    return v
  if isinstance(v, int):
    return np.asarray(v, dtype=np.int32).view(JArray)
  if isinstance(v, float):
    return np.asarray(v, dtype=np.float32).view(JArray)
  if isinstance(v, np.generic):
    return asjax(v)
  return v


def _fori_loop(lo, hi, body, init):
  state = tuple(_carry(v) for v in init) if isinstance(init, tuple) else _carry(init)
  for i in range(lo, hi):
    state = body(i, state)
    if isinstance(state, tuple):
      state = tuple(_carry(v) for v in state)
  return state


def _dynamic_slice(operand, start_indices, slice_sizes):
  operand = _unwrap(operand)
  sel = []
  for st, sz, n in zip(np.asarray(_unwrap(start_indices)).tolist(), slice_sizes,
                       operand.shape):
    st = int(min(max(int(st), 0), n - int(sz)))
    sel.append(slice(st, st + int(sz)))
  return _wrap(operand[tuple(sel)])


def _dynamic_update_slice(operand, update, start_indices):
  operand = np.array(_unwrap(operand), copy=True)
  update = np.asarray(_unwrap(update))
  sel = []
  for st, sz, n in zip(start_indices, update.shape, operand.shape):
    st = int(min(max(int(st), 0), n - int(sz)))
    sel.append(slice(st, st + int(sz)))
  operand[tuple(sel)] = update
  return _wrap(operand)


def _dynamic_index_in_dim(operand, index, axis=0, keepdims=True):
  operand = _unwrap(operand)
  n = operand.shape[axis]
  index = int(index)
  if index < 0:
    index += n
  index = min(max(index, 0), n - 1)
  out = np.take(operand, [index] if keepdims else index, axis=axis)
  return _wrap(out)


def _cond(pred, true_fn, false_fn, *operands):
  return true_fn(*operands) if bool(pred) else false_fn(*operands)


def _scan(f, init, xs):
  carry, ys = init, []
  for x in xs:
    carry, y = f(carry, x)
    ys.append(y)
  return carry, ys


def _map_coordinates(input, coordinates, order, mode='constant', cval=0.0):  # pylint: disable=redefined-builtin
  """jax.scipy.ndimage.map_coordinates, order 1, modes 'constant' / 'nearest', as
  published in jax/_src/scipy/ndimage.py: explicit per-point loops in fp32."""
  assert order == 1 and mode in ('constant', 'nearest')
  inp = np.asarray(_unwrap(input), dtype=np.float32)
  coords = [np.asarray(_unwrap(c), dtype=np.float32) for c in coordinates]
  out = np.empty(coords[0].shape, np.float32)
  nd = inp.ndim
  f32 = np.float32
  with np.errstate(invalid='ignore'):
    for pos in np.ndindex(*coords[0].shape):
      axes = []
      for d in range(nd):
        c = coords[d][pos]
        lo = np.floor(c)
        uw = f32(c - lo)
        lw = f32(f32(1.0) - uw)
        i0 = int(lo) if np.isfinite(lo) and abs(lo) < 2**31 else -10**9
        axes.append(((i0, lw), (i0 + 1, uw)))
      acc = None
      corners = [[]]
      for d in range(nd):  # itertools.product order: last axis fastest
        corners = [c + [k] for c in corners for k in (0, 1)]
      for corner in corners:
        w, ok, idx = None, True, []
        for d, k in enumerate(corner):
          i, wd = axes[d][k]
          if mode == 'constant' and not 0 <= i < inp.shape[d]:
            ok = False
          idx.append(min(max(i, 0), inp.shape[d] - 1))
          w = wd if w is None else f32(w * wd)
        val = inp[tuple(idx)] if ok else f32(cval)
        term = f32(w * val)
        acc = term if acc is None else f32(acc + term)
      out[pos] = acc
  return _wrap(out)


def _conv_patches(lhs, filter_shape, window_strides, padding):
  """[b, 1, *sp] -> [b, prod(filter_shape), *sp]; SAME = zero padding."""
  lhs = _unwrap(lhs)
  assert lhs.shape[1] == 1 and str(padding).lower() == 'same'
  assert all(s == 1 for s in window_strides)
  sp = lhs.shape[2:]
  pads = [(0, 0), (0, 0)] + [((f - 1) // 2, f // 2) for f in filter_shape]
  p = np.pad(lhs, pads, mode='constant')
  outs = []
  for off in np.ndindex(*filter_shape):
    sel = (slice(None), 0) + tuple(slice(o, o + n) for o, n in zip(off, sp))
    outs.append(p[sel])
  return _wrap(np.stack(outs, axis=1))


def _vmap(fn):
  def mapped(*args):
    n = len(args[0])
    res = [fn(*[a[i] for a in args]) for i in range(n)]
    return _wrap(np.stack([np.asarray(_unwrap(r)) for r in res]))
  return mapped


def install():
  """Registers the stand-in modules in sys.modules (idempotent)."""
  if 'jax' in sys.modules and getattr(sys.modules['jax'], '_sofima_shim', False):
    return
  jnp = _build_jnp()
  jax = types.ModuleType('jax')
  jax._sofima_shim = True
  jax.numpy = jnp
  jax.jit = _jit
  jax.vmap = _vmap
  jax.Array = np.ndarray
  lax = types.ModuleType('jax.lax')
  lax.fori_loop = _fori_loop
  lax.dynamic_slice = _dynamic_slice
  lax.conv_general_dilated_patches = _conv_patches
  lax.dynamic_update_slice = _dynamic_update_slice
  lax.dynamic_index_in_dim = _dynamic_index_in_dim
  lax.cond = _cond
  lax.scan = _scan
  jax.lax = lax
  jscipy = types.ModuleType('jax.scipy')
  jndimage = types.ModuleType('jax.scipy.ndimage')
  jndimage.map_coordinates = _map_coordinates
  jscipy.ndimage = jndimage
  jax.scipy = jscipy
  sys.modules.update({'jax.scipy': jscipy, 'jax.scipy.ndimage': jndimage})
  tree_util = types.ModuleType('jax.tree_util')
  tree_util.register_dataclass = lambda *a, **k: None
  jax.tree_util = tree_util
  default_device = lambda *a, **k: None
  jax.default_device = default_device
  sys.modules.update({'jax': jax, 'jax.numpy': jnp, 'jax.lax': lax,
                      'jax.tree_util': tree_util})

  if 'dataclasses_json' not in sys.modules:
    dj = types.ModuleType('dataclasses_json')

    class DataClassJsonMixin:

      def to_dict(self):
        return dataclasses.asdict(self)

    dj.DataClassJsonMixin = DataClassJsonMixin
    dj.dataclass_json = lambda cls: cls
    sys.modules['dataclasses_json'] = dj

  # connectomics.common.{geom_utils, utils}: restated in oracle/flow_oracle.py.
  import os
  sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', '..'))
  from oracle import flow_oracle  # pylint: disable=g-import-not-at-top
  conn = types.ModuleType('connectomics')
  common = types.ModuleType('connectomics.common')
  geom = types.ModuleType('connectomics.common.geom_utils')
  geom.integral_image = flow_oracle.integral_image
  geom.query_integral_image = lambda s, d, st: flow_oracle.query_integral_image(
      _unwrap(s), d, st)
  utils = types.ModuleType('connectomics.common.utils')
  utils.batch = flow_oracle.batch
  common.geom_utils, common.utils = geom, utils
  # bounding_box: only the class is needed (type annotations, tests' start/size).
  from sofima_b200 import compat  # pylint: disable=g-import-not-at-top
  bbox = types.ModuleType('connectomics.common.bounding_box')
  bbox.BoundingBox = compat.BoundingBox
  common.bounding_box = bbox
  sys.modules['connectomics.common.bounding_box'] = bbox
  conn.common = common
  sys.modules.update({'connectomics': conn, 'connectomics.common': common,
                      'connectomics.common.geom_utils': geom,
                      'connectomics.common.utils': utils})


def load_reference(name: str, root: str = '/root/reference'):
  """Imports /root/reference/<name>.py as module `sofima.<name>` via the shim."""
  import importlib.util
  import os
  install()
  if 'sofima' not in sys.modules:
    pkg = types.ModuleType('sofima')
    pkg.__path__ = [root]
    sys.modules['sofima'] = pkg
  full = f'sofima.{name}'
  if full in sys.modules:
    return sys.modules[full]
  spec = importlib.util.spec_from_file_location(full, os.path.join(root, f'{name}.py'))
  mod = importlib.util.module_from_spec(spec)
  sys.modules[full] = mod
  spec.loader.exec_module(mod)
  return mod
