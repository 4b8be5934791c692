"""Drop-in for `sofima.mesh` (reference mesh.py) running on the B200 backend.

Same public names and signatures as the reference:

  IntegrationConfig        mesh.py:282-338
  inplane_force            mesh.py:42-169
  elastic_mesh_3d          mesh.py:192-279   (MESH_LINK_DIRECTIONS mesh.py:172-189)
  velocity_verlet          mesh.py:371-521
  relax_mesh               mesh.py:524-608

Positions are in the reference's relative format, component-major
`[2 or 3, ..., y, x]`.  Inputs may be NumPy arrays (copied to the GPU, results
come back as NumPy) or CUDA torch tensors (results stay on the device).  All
arithmetic runs in the CUDA kernels of csrc/mesh.cu through the C ABI; there is
no CPU path here.
"""

from __future__ import annotations

import collections.abc
import ctypes
import dataclasses
import json
import logging
from typing import Any, Callable, Sequence

import numpy as np

from . import _native

try:  # the reference derives from dataclasses_json.DataClassJsonMixin
  import dataclasses_json  # pytype: disable=import-error
  _JsonBase = dataclasses_json.DataClassJsonMixin
except ImportError:  # not installable offline: minimal stand-in, same method names

  class _JsonBase:

    def to_dict(self) -> dict[str, Any]:
      return dataclasses.asdict(self)

    def to_json(self, **kw) -> str:
      return json.dumps(self.to_dict(), **kw)

    @classmethod
    def from_dict(cls, kvs: dict[str, Any]):
      names = {f.name for f in dataclasses.fields(cls)}
      return cls(**{k: v for k, v in kvs.items() if k in names})

    @classmethod
    def from_json(cls, s: str):
      return cls.from_dict(json.loads(s))


@dataclasses.dataclass(frozen=True)
class IntegrationConfig(_JsonBase):
  """Parameters for numerical integration of the mesh state (mesh.py:282-338)."""

  dt: float  # time step size
  gamma: float  # damping constant
  k0: float  # spring constant for inter-section springs
  k: float  # spring constant for intra-section springs
  stride: tuple[float, float] | tuple[float, float, float]
  num_iters: int  # number of time steps to execute at once
  max_iters: int  # upper bound for simulation time
  stop_v_max: float  # stop when all node velocities are below this value

  fire: bool = True  # use the Fast Inertial Relaxation Engine
  f_alpha: float = 0.99
  f_inc: float = 1.1
  f_dec: float = 0.5
  alpha: float = 0.1
  n_min: int = 5
  dt_max: float = 10.0  # in units of 'dt'

  start_cap: float = 1e6
  final_cap: float = 1e6
  cap_scale: float = 1.1
  cap_upscale_every: int = 100

  prefer_orig_order: bool = False
  remove_drift: bool = False

  def __post_init__(self):
    object.__setattr__(self, 'stride', tuple(self.stride))


MESH_LINK_DIRECTIONS = (  # xyz, mesh.py:172-189
    (1, 0, 0), (0, 1, 0), (0, 0, 1),
    (1, 1, 0), (-1, 1, 0), (1, 0, 1), (-1, 0, 1), (0, 1, 1), (0, -1, 1),
    (1, 1, 1), (1, 1, -1), (1, -1, 1), (-1, 1, 1),
)

_INPLANE, _MESH3D = 0, 1


def _torch():
  import torch  # plumbing: device memory and streams
  return torch


def _is_tensor(a) -> bool:
  return type(a).__module__.startswith('torch')


def _to_device(a, ctx: _native.Context, copy: bool):
  """fp32 contiguous CUDA tensor holding `a` (a fresh buffer if `copy`)."""
  torch = _torch()
  dev = torch.device('cuda', ctx.device)
  if _is_tensor(a):
    if not a.is_cuda:
      t = a.to(dev, dtype=torch.float32)
      return t.contiguous()
    t = a.to(dtype=torch.float32)
    t = t.contiguous()
    return t.clone() if (copy and t.data_ptr() == a.data_ptr()) else t
  host = np.ascontiguousarray(np.asarray(a), dtype=np.float32)
  return torch.from_numpy(host).to(dev)


def _from_device(t, like):
  if _is_tensor(like):
    return t
  return _to_host(t)


def _to_host(t):
  """CUDA tensor -> NumPy array."""
  nbytes = t.numel() * t.element_size()
  if (1 << 20) <= nbytes <= (1 << 30):
    # Read back through torch's caching pinned-host allocator: a pageable destination costs a
    # bounce copy plus a page fault per 4 KB of fresh memory (14 ms for a 2048^2 mesh against
    # 0.7 ms of DMA).  The array keeps the pinned block alive; it returns to the cache when
    # the caller drops the array.
    torch = _torch()
    host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    host.copy_(t)
    return host.numpy()
  return t.cpu().numpy()


def _shape_pod(shape: Sequence[int], kind: int) -> _native.MeshShape:
  if kind == _INPLANE:
    if len(shape) != 4 or shape[0] != 2:
      raise ValueError(f'expected a [2, z, y, x] mesh, got shape {tuple(shape)}')
    return _native.MeshShape(2, shape[1], 1, shape[2], shape[3], 0)
  if len(shape) < 4 or shape[0] != 3:
    raise ValueError(
        f'expected a [3, [batch..], z, y, x] mesh, got shape {tuple(shape)}')
  nb = int(np.prod(shape[1:-3])) if len(shape) > 4 else 1
  return _native.MeshShape(3, nb, shape[-3], shape[-2], shape[-1], len(shape) - 4)


def _stride3(stride, kind: int):
  if kind == _INPLANE:
    if len(stride) != 2:
      raise ValueError('stride must be 2D.')
    vals = (float(stride[0]), float(stride[1]), 0.0)
  else:
    if not isinstance(stride, collections.abc.Sequence):
      stride = (stride,) * 3
    if len(stride) != 3:
      raise ValueError('stride must be 3D.')
    vals = tuple(float(s) for s in stride)
  return (ctypes.c_double * 3)(*vals)


def _force(kind: int, x, k, stride, prefer_orig_order, links=None):
  ctx = _native.Context.get(x.device.index if _is_tensor(x) and x.is_cuda else None)
  xd = _to_device(x, ctx, copy=False)
  shape = _shape_pod(tuple(xd.shape), kind)
  out = _torch().empty_like(xd)
  ctx.bind_stream()
  if links is None:
    links_p, nlinks = None, 0
  else:
    flat = [int(v) for d in links for v in d]
    if len(flat) != 3 * len(links):
      raise ValueError('links must be xyz triples')
    if any(abs(v) > 1 for v in flat):
      raise ValueError('Only |v| <= 1 values supported within links.')
    links_p, nlinks = (ctypes.c_int32 * len(flat))(*flat), len(links)
  rc = _native.lib().sofima_mesh_force_links(
      ctx.handle, kind, xd.data_ptr(), ctypes.byref(shape), float(k),
      _stride3(stride, kind), int(bool(prefer_orig_order)), links_p, nlinks,
      out.data_ptr())
  _native.check(ctx.handle, rc)
  return _from_device(out, x)


def inplane_force(x, k: float, stride: Sequence[float],
                  prefer_orig_order: bool = False):
  """In-plane forces on the nodes of a spring mesh (mesh.py:42-169).

  Args:
    x: [2, z, y, x] array of mesh node positions, in relative format
    k: spring constant
    stride: XY stride of the spring mesh grid
    prefer_orig_order: use the fold-preventing force formulation

  Returns:
    [2, z, y, x] array of forces
  """
  if len(stride) != 2:
    raise ValueError('stride must be 2D.')
  return _force(_INPLANE, x, k, stride, prefer_orig_order)


def elastic_mesh_3d(x, k: float, stride, prefer_orig_order: bool = False,
                    links=MESH_LINK_DIRECTIONS):
  """Internal forces on the nodes of a 3d spring mesh (mesh.py:192-279).

  Args:
    x: [3, [batch..], z, y, x] array of mesh node positions, in relative format
    k: spring constant for springs along the x direction
    stride: XYZ stride of the spring mesh grid (scalar or 3 values)
    prefer_orig_order: use the fold-preventing force formulation
    links: XYZ tuples of node links to consider, components in {-1, 0, 1}

  Returns:
    array of forces, same shape as x
  """
  assert x.shape[0] == 3
  use_links = None if tuple(map(tuple, links)) == MESH_LINK_DIRECTIONS else links
  return _force(_MESH3D, x, k, stride, prefer_orig_order, use_links)


def _force_kind(mesh_force: Callable[..., Any]) -> int:
  if mesh_force is inplane_force:
    return _INPLANE
  if mesh_force is elastic_mesh_3d:
    return _MESH3D
  raise NotImplementedError(
      'The CUDA backend runs the built-in force fields only '
      '(sofima_b200.mesh.inplane_force / elastic_mesh_3d); arbitrary Python '
      f'callables such as {mesh_force!r} cannot be traced into the kernel.')


def _config_pod(config: IntegrationConfig, kind: int) -> _native.IntegrationConfigPod:
  stride = tuple(config.stride)
  if kind == _INPLANE and len(stride) != 2:
    raise ValueError('stride must be 2D.')
  if kind == _MESH3D and len(stride) != 3:
    raise ValueError('stride must be 3D.')
  pod = _native.IntegrationConfigPod()
  for name in ('dt', 'gamma', 'k0', 'k', 'f_alpha', 'f_inc', 'f_dec', 'alpha',
               'dt_max', 'start_cap', 'final_cap', 'cap_scale'):
    setattr(pod, name, float(getattr(config, name)))
  for i in range(3):
    pod.stride[i] = float(stride[i]) if i < len(stride) else 0.0
  pod.num_iters = int(config.num_iters)
  pod.fire = int(bool(config.fire))
  pod.n_min = int(config.n_min)
  pod.cap_upscale_every = int(config.cap_upscale_every)
  pod.prefer_orig_order = int(bool(config.prefer_orig_order))
  pod.remove_drift = int(bool(config.remove_drift))
  return pod


class _Chunk:
  """Device-resident solver state for repeated velocity_verlet calls."""

  def __init__(self, x, v, prev, config: IntegrationConfig, kind: int, prev_fn=None):
    like = x
    dev = x.device.index if _is_tensor(x) and x.is_cuda else None
    self.ctx = _native.Context.get(dev)
    self.like = like
    self.kind = kind
    self.x = _to_device(x, self.ctx, copy=True)
    self.v = (_torch().zeros_like(self.x) if v is None
              else _to_device(v, self.ctx, copy=True))
    self.a = _torch().empty_like(self.x)
    self.prev = None if prev is None else _to_device(prev, self.ctx, copy=False)
    if self.v.shape != self.x.shape:
      raise ValueError('x and v must have the same shape')
    if self.prev is not None and self.prev.shape != self.x.shape:
      raise ValueError('x and prev must have the same shape')
    self.shape = _shape_pod(tuple(self.x.shape), kind)
    self.pod = _config_pod(config, kind)
    self.config = config
    self.target = None
    if prev_fn is not None:
      self.target = _stitch_target(prev_fn, kind, tuple(self.x.shape), self.ctx)

  def run(self, dt: float, alpha: float, cap: float):
    """One velocity_verlet call.  Returns (dt, alpha, n_pos, cap, e_kin, v_max)."""
    c_dt, c_alpha, c_cap = (ctypes.c_float(dt), ctypes.c_float(alpha),
                            ctypes.c_float(cap))
    n_pos, e_kin, v_max = ctypes.c_int32(0), ctypes.c_double(0), ctypes.c_float(0)
    self.ctx.bind_stream()
    if self.target is not None:
      rc = _native.lib().sofima_mesh_chunk_stitch(
          self.ctx.handle, self.kind, self.x.data_ptr(), self.v.data_ptr(), self.a.data_ptr(),
          ctypes.byref(self.target), ctypes.byref(self.shape), ctypes.byref(self.pod),
          ctypes.byref(c_dt), ctypes.byref(c_alpha), ctypes.byref(c_cap),
          ctypes.byref(n_pos), ctypes.byref(e_kin), ctypes.byref(v_max))
      _native.check(self.ctx.handle, rc)
      return (np.float32(c_dt.value), np.float32(c_alpha.value), int(n_pos.value),
              np.float32(c_cap.value), float(e_kin.value), np.float32(v_max.value))
    rc = _native.lib().sofima_mesh_chunk(
        self.ctx.handle, self.kind, self.x.data_ptr(), self.v.data_ptr(),
        self.a.data_ptr(), None if self.prev is None else self.prev.data_ptr(),
        ctypes.byref(self.shape), ctypes.byref(self.pod), ctypes.byref(c_dt),
        ctypes.byref(c_alpha), ctypes.byref(c_cap), ctypes.byref(n_pos),
        ctypes.byref(e_kin), ctypes.byref(v_max))
    _native.check(self.ctx.handle, rc)
    return (np.float32(c_dt.value), np.float32(c_alpha.value), int(n_pos.value),
            np.float32(c_cap.value), float(e_kin.value), np.float32(v_max.value))


def _stitch_target(prev_fn, kind: int, x_shape, ctx):
  """C-ABI descriptor of a device-side prev_fn (stitch_elastic.StitchTarget)."""
  describe = getattr(prev_fn, '_sofima_device_target', None)
  if describe is None:
    raise NotImplementedError(
        'The CUDA backend evaluates prev_fn inside the step kernel sequence, so it '
        'must be a device-side target description: build it with '
        'sofima_b200.stitch_elastic.target_mesh_fn(nbors, fx, fy, stride) (the '
        'stitching prev_fn of stitch_elastic.compute_target_mesh); arbitrary Python '
        f'callables such as {prev_fn!r} cannot be traced into it.')
  pod = describe(x_shape, ctx)
  if pod.ndim != (2 if kind == _INPLANE else 3):
    raise ValueError('prev_fn and mesh_force disagree on the mesh dimensionality')
  return pod


def velocity_verlet(x, v, prev, config: IntegrationConfig, force_cap: float,
                    fire_dt: float | None = None, fire_alpha: float | None = None,
                    mesh_force=inplane_force, prev_fn=None):
  """Executes `config.num_iters` (damped) velocity Verlet / FIRE steps.

  Same contract as the reference (mesh.py:371-521): returns
  `(x, v, a)` or, with FIRE, `(x, v, a, dt, alpha, n_pos, cap)`.  Inputs are not
  modified.
  """
  if prev is not None and prev_fn is not None:
    raise ValueError('Only one of: "prev" and "prev_fn" can be specified.')
  kind = _force_kind(mesh_force)
  chunk = _Chunk(x, v, prev, config, kind, prev_fn)
  dt = config.dt if fire_dt is None else fire_dt
  alpha = config.alpha if fire_alpha is None else fire_alpha
  dt, alpha, n_pos, cap, _, _ = chunk.run(float(dt), float(alpha), float(force_cap))
  out = tuple(_from_device(t, x) for t in (chunk.x, chunk.v, chunk.a))
  if config.fire:
    return out + (dt, alpha, n_pos, cap)
  return out


def relax_mesh(x, prev, config: IntegrationConfig, mesh_force=inplane_force,
               prev_fn=None):
  """Simulates mesh relaxation (mesh.py:524-608).

  Args:
    x: [2, z, y, x] (or [3, ..., z, y, x]) array of mesh node positions
    prev: optional array of the same shape against which to compute the force
      due to 0-length springs
    config: simulation parameters
    mesh_force: `inplane_force` or `elastic_mesh_3d` of this module
    prev_fn: optional device-side target description replacing `prev`, see
      `stitch_elastic.target_mesh_fn` (the reference takes a JAX callable here)

  Returns:
    tuple of: updated mesh positions, kinetic energy history, number of
    simulation steps executed
  """
  t = 0
  dt = config.dt
  alpha = config.alpha
  e_kin = []
  cap = config.start_cap

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

  kind = _force_kind(mesh_force)
  chunk = _Chunk(x, None, prev, config, kind, prev_fn)  # state stays on the device

  while t < config.max_iters:
    dt_n, alpha_n, n_pos, cap_n, ek, v_max = chunk.run(
        float(dt), float(alpha), float(cap))
    t += config.num_iters
    e_kin.append(ek)
    if config.fire:
      dt, alpha, cap = dt_n, alpha_n, cap_n
      logging.info(
          't=%r: dt=%f, alpha=%f, n_pos=%d, cap=%f, v_max=%f, e_kin=%f',
          t, dt, alpha, n_pos, cap, v_max, ek)

    if v_max < np.float32(config.stop_v_max):
      if np.float32(cap) >= np.float32(config.final_cap):
        break
      # Increase cap to ensure progress towards the termination condition.
      cap = min(np.float32(cap) * np.float32(config.cap_scale),
                np.float32(config.final_cap))

  return _from_device(chunk.x, x), e_kin, t
