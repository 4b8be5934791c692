"""Multi-GPU check of the row-sharded mesh (launch with torchrun, one rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29511 tests/multi/run_sharded_mesh.py [ny nx iters sections]

Every rank relaxes its slab with relax_mesh_sharded; rank 0 also solves the whole
mesh on its own GPU with mesh.relax_mesh and compares (bit-exact expected).  Prints
one JSON line with the timing of both.
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import torch.distributed as dist

from sofima_b200 import mesh, mesh_sharded


def main():
  ny = int(sys.argv[1]) if len(sys.argv) > 1 else 512
  nx = int(sys.argv[2]) if len(sys.argv) > 2 else 384
  iters = int(sys.argv[3]) if len(sys.argv) > 3 else 200
  nz = int(sys.argv[4]) if len(sys.argv) > 4 else 2
  local = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local)
  dist.init_process_group('nccl', device_id=torch.device('cuda', local))
  rank, world = dist.get_rank(), dist.get_world_size()
  rng = np.random.default_rng(5)
  shape = (2, nz, ny, nx)
  prev = (rng.standard_normal(shape) * 4).astype(np.float32)
  prev[rng.random(shape) < 0.01] = np.nan
  x0 = np.zeros(shape, np.float32)
  cfg = mesh.IntegrationConfig(dt=0.001, gamma=0.0, k0=0.1, k=0.1, stride=(40.0, 40.0),
                               num_iters=iters, max_iters=2 * iters, stop_v_max=0.0,
                               fire=True, dt_max=1000.0, prefer_orig_order=True)
  y0, y1 = mesh_sharded.partition_rows(ny, world)[rank]
  xs = torch.from_numpy(np.ascontiguousarray(x0[:, :, y0:y1])).cuda()
  ps = torch.from_numpy(np.ascontiguousarray(prev[:, :, y0:y1])).cuda()
  torch.cuda.synchronize()
  dist.barrier()
  t0 = time.perf_counter()
  got, e_kin, t = mesh_sharded.relax_mesh_sharded(xs, ps, cfg)
  torch.cuda.synchronize()
  dist.barrier()
  t_sharded = time.perf_counter() - t0
  slabs = [torch.empty((2, nz, b - a, nx), device='cuda') for a, b in
           mesh_sharded.partition_rows(ny, world)]
  dist.all_gather(slabs, got.contiguous())
  ok, t_single, err = True, None, 0.0
  if rank == 0:
    full = torch.cat(slabs, dim=2).cpu().numpy()
    t0 = time.perf_counter()
    want, ek_w, t_w = mesh.relax_mesh(x0, prev, cfg)
    t_single = time.perf_counter() - t0
    same_nan = np.array_equal(np.isnan(full), np.isnan(want))
    err = float(np.nanmax(np.abs(full - want)))
    ok = same_nan and err == 0.0 and t == t_w and np.allclose(e_kin, ek_w, rtol=1e-9)
    print(json.dumps(dict(world=world, shape=shape, iters=iters, steps=t, ok=bool(ok),
                          max_abs_err=err, sharded_s=t_sharded, single_gpu_s=t_single,
                          us_per_step_sharded=t_sharded / t * 1e6)))
  mesh_sharded.clear_shard_cache()
  dist.destroy_process_group()
  if not ok:
    sys.exit(1)


if __name__ == '__main__':
  main()
