"""One spring mesh relaxed on several GPUs of a node (BASELINE config 3).

The reference always solves a section on one device (processor/mesh.py:462); this
module shards the rows of a [2, z, y, x] mesh over the ranks of a torch.distributed
group (one process per GPU) and runs the same integrator as `mesh.relax_mesh`
(reference mesh.py:524-608).  Inside a chunk the ranks synchronise on the device:
the step kernel reads the neighbours' boundary rows through peer-mapped memory and
the FIRE partial sums travel with the step flags (csrc/mesh.cu, "Sharded mesh");
the host only joins the ranks between chunks to combine e_kin / v_max -- the same
place where the reference synchronises (mesh.py:585).

torch.distributed is plumbing here: it carries the 128-byte IPC blobs once and one
small all-reduce per chunk; there is no collective on the data path.
"""

from __future__ import annotations

import ctypes
import logging
from typing import Sequence

import numpy as np

from . import _native
from . import mesh as mesh_lib

ROW_ALIGN = 32  # tile height of the step kernel: slab boundaries fall on tile edges


def partition_rows(ny: int, nranks: int, align: int = ROW_ALIGN) -> list[tuple[int, int]]:
  """Splits `ny` rows into `nranks` contiguous slabs [y0, y1).

  Every slab except the last is a multiple of `align` rows; slabs are as even as the
  alignment allows.  Raises if there are fewer aligned blocks than ranks.
  """
  if nranks < 1:
    raise ValueError('nranks must be positive')
  if nranks == 1:
    return [(0, ny)]
  blocks = -(-ny // align)  # number of (possibly partial) aligned row blocks
  if blocks < nranks:
    raise ValueError(f'{ny} rows cannot be split into {nranks} slabs of >= {align} rows')
  base, extra = divmod(blocks, nranks)
  out, y = [], 0
  for r in range(nranks):
    nblk = base + (1 if r < extra else 0)
    y1 = min(ny, y + nblk * align)
    out.append((y, y1))
    y = y1
  assert out[-1][1] == ny
  return out


class ShardedMesh:
  """Rank-local slab of a row-sharded mesh, with the solver state on the device.

  Creating one is expensive (a peer-mapped allocation, an IPC handle exchange and two
  barriers, ~50 ms): `relax_mesh_sharded` keeps the object of every (slab shape, group) and
  only reloads the state for the next solve (`load`).
  """

  def __init__(self, x_local, prev_local, config: mesh_lib.IntegrationConfig, group=None):
    import torch
    import torch.distributed as dist
    self._torch, self._dist = torch, dist
    self.group = group
    self.rank = dist.get_rank(group) if dist.is_initialized() else 0
    self.nranks = dist.get_world_size(group) if dist.is_initialized() else 1
    self.ctx = _native.Context.get(torch.cuda.current_device())
    shape = tuple(x_local.shape)
    if len(shape) != 4 or shape[0] != 2:
      raise ValueError(f'expected a [2, z, y, x] slab, got {shape}')
    self.shape = shape
    shp = _native.MeshShape(2, shape[1], 1, shape[2], shape[3])
    lib = _native.lib()
    self.ctx.bind_stream()
    h = ctypes.c_void_p()
    _native.check(self.ctx.handle, lib.sofima_shard_create(
        self.ctx.handle, self.rank, self.nranks, ctypes.byref(shp), ctypes.byref(h)))
    self.handle = h
    blob = ctypes.create_string_buffer(128)
    _native.check(self.ctx.handle, lib.sofima_shard_export(h, blob))
    if self.nranks > 1:
      blobs = [None] * self.nranks
      dist.all_gather_object(blobs, blob.raw, group=group)
      allb = b''.join(blobs)
    else:
      allb = blob.raw
    _native.check(self.ctx.handle, lib.sofima_shard_connect(h, allb))
    nodes = torch.tensor([shape[1] * shape[2] * shape[3]], dtype=torch.int64,
                         device=torch.device('cuda', self.ctx.device))
    if self.nranks > 1:
      dist.all_reduce(nodes, group=group)
    self.global_nodes = int(nodes.item())
    self.load(x_local, prev_local, config)

  def load(self, x_local, prev_local, config: mesh_lib.IntegrationConfig):
    """(Re)loads positions (velocities = 0) and the fixed target of the slab."""
    if tuple(x_local.shape) != self.shape:
      raise ValueError(f'slab shape {tuple(x_local.shape)} != {self.shape}')
    self.config = config
    self.like = x_local
    self.pod = mesh_lib._config_pod(config, 0)
    x = mesh_lib._to_device(x_local, self.ctx, copy=False)
    prev = None if prev_local is None else mesh_lib._to_device(prev_local, self.ctx, copy=False)
    self.ctx.bind_stream()
    _native.check(self.ctx.handle, _native.lib().sofima_shard_set_state(
        self.handle, x.data_ptr(), None, None if prev is None else prev.data_ptr()))
    self._join()

  def _join(self):
    """All ranks' queued work is complete (chunk boundaries only)."""
    self._torch.cuda.synchronize()
    if self.nranks > 1 and self._dist.is_initialized():
      self._dist.barrier(group=self.group)

  def run(self, dt: float, alpha: float, cap: float):
    """One velocity_verlet chunk on the whole mesh.

    Returns (dt, alpha, n_pos, cap, e_kin, v_max) with e_kin / v_max reduced over
    the ranks.
    """
    torch, dist = self._torch, self._dist
    st = _native.MeshState()
    self.ctx.bind_stream()
    rc = _native.lib().sofima_shard_chunk(
        self.handle, ctypes.byref(self.pod), float(dt), float(alpha), float(cap),
        self.global_nodes, ctypes.byref(st))
    _native.check(self.ctx.handle, rc)
    red = torch.tensor([st.e_kin, float(st.v_max)], dtype=torch.float64,
                       device=torch.device('cuda', self.ctx.device))
    if self.nranks > 1:
      parts = [torch.empty_like(red) for _ in range(self.nranks)]
      dist.all_gather(parts, red, group=self.group)  # fixed rank order: deterministic sum
      e_kin = float(sum(float(p[0]) for p in parts))
      vals = [float(p[1]) for p in parts]
      v_max = float('nan') if any(np.isnan(vals)) else max(vals)
    else:
      e_kin, v_max = float(red[0]), float(red[1])
    self._join()
    if self.config.fire:
      return (np.float32(st.dt), np.float32(st.alpha), int(st.n_pos), np.float32(st.cap),
              e_kin, np.float32(v_max))
    return np.float32(dt), np.float32(alpha), -1, np.float32(cap), e_kin, np.float32(v_max)

  def state(self):
    """(x, v, a) of the local slab as CUDA tensors."""
    torch = self._torch
    dev = torch.device('cuda', self.ctx.device)
    out = [torch.empty(self.shape, dtype=torch.float32, device=dev) for _ in range(3)]
    _native.check(self.ctx.handle, _native.lib().sofima_shard_get_state(
        self.handle, *(t.data_ptr() for t in out)))
    return out

  def close(self):
    if getattr(self, 'handle', None):
      self._join()
      _native.lib().sofima_shard_destroy(self.handle)
      self.handle = None

  def __del__(self):
    try:
      if getattr(self, 'handle', None):
        _native.lib().sofima_shard_destroy(self.handle)
    except Exception:  # pylint: disable=broad-except
      pass


def relax_mesh_sharded(x_local, prev_local, config: mesh_lib.IntegrationConfig, group=None):
  """`mesh.relax_mesh` (reference mesh.py:524-608) on a row-sharded mesh.

  Every rank passes its own slab `[2, z, rows(rank), x]` (see `partition_rows`) and
  gets its slab of the solution back; `e_kin` and the step count are global.
  """
  if config.start_cap != config.final_cap:
    if not config.fire:
      raise NotImplementedError('Adaptive force capping is only supported with FIRE.')
    if config.cap_scale <= 1:
      raise ValueError('The scaling factor for the force cap has to be larger '
                       'than 1 when the initial and final cap are different.')
  shard = _cached_shard(x_local, prev_local, config, group)
  t, dt, alpha, cap, e_kin = 0, config.dt, config.alpha, config.start_cap, []
  try:
    while t < config.max_iters:
      dt_n, alpha_n, n_pos, cap_n, ek, v_max = shard.run(float(dt), float(alpha), float(cap))
      t += config.num_iters
      e_kin.append(ek)
      if config.fire:
        dt, alpha, cap = dt_n, alpha_n, cap_n
        logging.info('t=%r: dt=%f, alpha=%f, n_pos=%d, cap=%f, v_max=%f, e_kin=%f', t, dt,
                     alpha, n_pos, cap, v_max, ek)
      if v_max < np.float32(config.stop_v_max):
        if np.float32(cap) >= np.float32(config.final_cap):
          break
        cap = min(np.float32(cap) * np.float32(config.cap_scale),
                  np.float32(config.final_cap))
    x = shard.state()[0]
  except Exception:
    _drop_shard(shard)  # never reuse a shard whose ranks may be out of step
    raise
  return mesh_lib._from_device(x, x_local), e_kin, t


# Shards of recent solves, keyed by (device, slab shape, group): all ranks of a group call
# relax_mesh_sharded in lockstep, so they hit or miss the cache together.
_SHARDS: dict = {}
_MAX_SHARDS = 4


def _cached_shard(x_local, prev_local, config, group) -> ShardedMesh:
  import torch
  key = (torch.cuda.current_device(), tuple(x_local.shape), id(group))
  shard = _SHARDS.get(key)
  if shard is not None and getattr(shard, 'handle', None):
    shard.load(x_local, prev_local, config)
    return shard
  while len(_SHARDS) >= _MAX_SHARDS:
    _SHARDS.pop(next(iter(_SHARDS))).close()
  shard = _SHARDS[key] = ShardedMesh(x_local, prev_local, config, group)
  return shard


def _drop_shard(shard: ShardedMesh):
  for k, v in list(_SHARDS.items()):
    if v is shard:
      del _SHARDS[k]
  try:
    shard.close()
  except Exception:  # pylint: disable=broad-except
    pass


def clear_shard_cache():
  """Frees the peer-mapped buffers of all cached shards (collective: call on every rank,
  before the process group is destroyed)."""
  for shard in list(_SHARDS.values()):
    try:
      shard.close()
    except Exception:  # pylint: disable=broad-except
      pass
  _SHARDS.clear()


import atexit  # pylint: disable=wrong-import-position
atexit.register(clear_shard_cache)  # while the CUDA context and the library are still alive
