"""Small, fixed workloads for ncu captures (never a bench number)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from sofima_b200 import _native, flow_field, mesh

which = sys.argv[1] if len(sys.argv) > 1 else 'both'
dev = torch.device('cuda', 0)
ctx = _native.Context.get(0)
if which in ('mesh', 'both'):
  iters = int(os.environ.get('PROF_MESH_ITERS', '30'))
  cfg = bench.mesh_config(mesh, iters)
  prev = bench.synth_mesh(bench.MESH_N, 7, dev)
  ch = mesh._Chunk(torch.zeros_like(prev), None, prev, cfg, 0)
  for _ in range(2):
    ch.run(cfg.dt, cfg.alpha, cfg.start_cap)
  torch.cuda.synchronize()
if which in ('flow', 'both'):
  tiles = bench.synth_tile_pairs(1, bench.FLOW_TILE, 100, dev)
  g = (bench.FLOW_TILE - 120) // 40
  oyx = np.array(np.where(np.ones((g, g), bool))).T
  job = flow_field._FlowJob(ctx, oyx, (bench.FLOW_TILE,) * 2, (bench.FLOW_TILE,) * 2, (160, 160), (160, 160), (40, 40), 1024)
  for _ in range(2):
    job.run(tiles[0][0], tiles[0][1])
  torch.cuda.synchronize()
print('done')
