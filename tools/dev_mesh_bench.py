"""Developer timing of the mesh step kernel (not the bench contract; see bench.py)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sofima_b200 import mesh

def run(n=2048, z=1, iters=1000, poo=True, fire=True):
  cfg = mesh.IntegrationConfig(dt=0.001, gamma=0.0, k0=0.1, k=0.1, stride=(40., 40.),
                               num_iters=iters, max_iters=iters, stop_v_max=0.0, fire=fire,
                               dt_max=1000.0, prefer_orig_order=poo)
  g = torch.Generator(device='cuda').manual_seed(0)
  prev = torch.randn((2, z, n, n), device='cuda', generator=g) * 4
  x = torch.zeros_like(prev)
  ch = mesh._Chunk(x, None, prev, cfg, 0)
  ch.run(cfg.dt, cfg.alpha, cfg.start_cap)  # warm-up chunk
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  out = ch.run(cfg.dt, cfg.alpha, cfg.start_cap)
  e1.record(); torch.cuda.synchronize()
  ms = e0.elapsed_time(e1)
  nups = z * n * n * iters / (ms * 1e-3)
  print(json.dumps(dict(n=n, z=z, iters=iters, poo=poo, fire=fire, ms=ms, us_per_step=ms*1e3/iters,
                        gnups=nups/1e9, hbm_frac=nups*56/6534.5e9, state=[float(o) for o in out])))

def run_sharded_single(n=2048, iters=1000):
  from sofima_b200 import mesh_sharded
  cfg = mesh.IntegrationConfig(dt=0.001, gamma=0.0, k0=0.1, k=0.1, stride=(40., 40.),
                               num_iters=iters, max_iters=iters, stop_v_max=0.0, fire=True,
                               dt_max=1000.0, prefer_orig_order=True)
  g = torch.Generator(device='cuda').manual_seed(0)
  prev = torch.randn((2, 1, n, n), device='cuda', generator=g) * 4
  sh = mesh_sharded.ShardedMesh(torch.zeros_like(prev), prev, cfg)
  sh.run(cfg.dt, cfg.alpha, cfg.start_cap)
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  sh.run(cfg.dt, cfg.alpha, cfg.start_cap)
  torch.cuda.synchronize()
  ms = (time.perf_counter() - t0) * 1e3
  print(json.dumps(dict(sharded_single_rank=True, n=n, iters=iters, us_per_step=ms * 1e3 / iters)))
  sh.close()


if __name__ == '__main__':
  if len(sys.argv) > 1 and sys.argv[1] == 'sharded':
    run(2048, 1, 1000, True, True)
    run_sharded_single()
    sys.exit(0)
  run(2048, 1, 1000, True, True)
  run(2048, 1, 1000, False, True)
  run(2048, 1, 1000, True, False)
  run(1024, 1, 1000, True, True)
  run(4096, 1, 300, True, True)
  run(102, 16, 1000, True, True)
  run(205, 1, 1000, True, True)
