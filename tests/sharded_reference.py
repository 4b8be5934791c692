"""CPU emulation of the row-sharded mesh protocol with the oracle arithmetic.

Every rank holds its slab plus ONE halo row per side and, per step, receives the
neighbours' boundary rows of (v, a) and everybody's partial <a, v> -- exactly the
data the CUDA step kernel reads through peer memory (csrc/mesh.cu, "Sharded mesh").
Used by the gloo tests to show that this decomposition reproduces the single-domain
integrator (reference mesh.py:371-521) bit for bit.
"""

import numpy as np
import torch
import torch.distributed as dist

from oracle import mesh_oracle as mo
from sofima_b200.mesh_sharded import partition_rows

F32 = np.float32


def _exchange_rows(arr, own, rank, world):
  """Fills the halo rows of `arr` [2, z, rows, x] with the neighbours' boundary rows."""
  lo_own, hi_own = own
  reqs, bufs = [], {}
  def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))
  if rank > 0:
    reqs.append(dist.isend(t(arr[:, :, lo_own]), rank - 1))
    bufs['up'] = torch.empty(arr[:, :, 0].shape, dtype=torch.float32)
    reqs.append(dist.irecv(bufs['up'], rank - 1))
  if rank < world - 1:
    reqs.append(dist.isend(t(arr[:, :, hi_own - 1]), rank + 1))
    bufs['dn'] = torch.empty(arr[:, :, 0].shape, dtype=torch.float32)
    reqs.append(dist.irecv(bufs['dn'], rank + 1))
  for r in reqs:
    r.wait()
  if 'up' in bufs:
    arr[:, :, lo_own - 1] = bufs['up'].numpy()
  if 'dn' in bufs:
    arr[:, :, hi_own] = bufs['dn'].numpy()


def sharded_relax(x_full, prev_full, cfg):
  """relax_mesh on the calling rank's slab; returns (x_slab, e_kin, t, (y0, y1))."""
  rank, world = dist.get_rank(), dist.get_world_size()
  ny = x_full.shape[2]
  y0, y1 = partition_rows(ny, world)[rank]
  lo, hi = max(y0 - 1, 0), min(y1 + 1, ny)
  own = (y0 - lo, y0 - lo + (y1 - y0))  # own rows inside the extended slab
  x = np.array(x_full[:, :, lo:hi], dtype=F32)
  prev = None if prev_full is None else np.array(prev_full[:, :, lo:hi], dtype=F32)
  v = np.zeros_like(x)
  sl = (slice(None), slice(None), slice(*own))

  t, dt, alpha, cap, e_kin = 0, F32(cfg.dt), F32(cfg.alpha), F32(cfg.start_cap), []
  dt_ceiling = F32(cfg.dt_max * cfg.dt)
  with np.errstate(all='ignore'):
    while t < cfg.max_iters:
      n_pos = 0
      a = mo.total_force(x, prev, cap, cfg, mo.inplane_force)
      _exchange_rows(a, own, rank, world)
      for _ in range(cfg.num_iters):
        hdt2 = F32(0.5) * (dt * dt)
        x = x + (dt * v + hdt2 * a)
        a_prev = a
        a = mo.total_force(x, prev, cap, cfg, mo.inplane_force)
        hdtg = (F32(0.5) * dt) * F32(cfg.gamma)
        fact0, fact1 = F32(1.0) / (F32(1.0) + hdtg), F32(1.0) - hdtg
        v = fact0 * (v * fact1 + (F32(0.5) * dt) * (a_prev + a))
        a_norm = mo._norm0(a) + F32(1e-6)
        v_norm = mo._norm0(v)
        part = torch.tensor([mo.power_f64(a[sl], v[sl])], dtype=torch.float64)
        parts = [torch.empty_like(part) for _ in range(world)]
        dist.all_gather(parts, part)
        power = sum(float(p) for p in parts)  # rank order, identical everywhere
        v = v + alpha * (a / a_norm * v_norm - v)
        pos = power >= 0
        n_pos = n_pos + 1 if pos else 0
        if pos:
          if n_pos > cfg.n_min:
            dt = min(dt * F32(cfg.f_inc), dt_ceiling)
            alpha = alpha * F32(cfg.f_alpha)
          if n_pos > 0 and n_pos % cfg.cap_upscale_every == 0:
            cap = F32(cfg.cap_scale) * cap
        else:
          dt = dt * F32(cfg.f_dec)
          alpha = F32(cfg.alpha)
        cap = min(cap, F32(cfg.final_cap))
        v = v * F32(1.0 if pos else 0.0)
        # the halo rows of v and a were computed without their own neighbours:
        # replace them by the owners' values (x is reproduced exactly from x, v, a)
        _exchange_rows(v, own, rank, world)
        _exchange_rows(a, own, rank, world)
      t += cfg.num_iters
      ek, v_max = mo.chunk_stats(v[sl])
      red = torch.tensor([ek, float(v_max)], dtype=torch.float64)
      allr = [torch.empty_like(red) for _ in range(world)]
      dist.all_gather(allr, red)
      e_kin.append(sum(float(r[0]) for r in allr))
      v_max = max(float(r[1]) for r in allr)
      if v_max < F32(cfg.stop_v_max):
        if F32(cap) >= F32(cfg.final_cap):
          break
        cap = min(F32(cap) * F32(cfg.cap_scale), F32(cfg.final_cap))
  return x[sl], e_kin, t, (y0, y1)
