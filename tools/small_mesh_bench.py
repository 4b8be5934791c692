"""Per-step time of small meshes: persistent solver vs one launch per step (timing only)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sofima_b200 import mesh, _native
import bench
from sofima_b200 import stitch_elastic

def run(shape, iters=1000, stitch=False):
  rng = np.random.default_rng(0)
  cfg = mesh.IntegrationConfig(dt=0.001, gamma=0., k0=0.01, k=0.1, stride=(40., 40.), num_iters=iters,
                               max_iters=iters, stop_v_max=0.0, dt_max=100, prefer_orig_order=True,
                               start_cap=0.1, final_cap=10., remove_drift=stitch)
  if stitch:
    fx, fy, x0, nbors, stride = bench.synth_stitch(4, 4, shape[2:])
    prev_fn = stitch_elastic.target_mesh_fn(nbors, fx, fy, stride)
    xd = torch.from_numpy(x0).cuda(); prev = None
  else:
    prev_fn = None
    prev = torch.from_numpy((rng.standard_normal(shape) * 4).astype(np.float32)).cuda()
    xd = torch.zeros_like(prev)
  out = {}
  for mode in ('1', '0'):
    os.environ['SOFIMA_MESH_PERSISTENT'] = mode
    mesh.relax_mesh(xd, prev, cfg, prev_fn=prev_fn)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
      mesh.relax_mesh(xd, prev, cfg, prev_fn=prev_fn)
    torch.cuda.synchronize()
    out['persistent' if mode == '1' else 'per_launch'] = (time.perf_counter() - t0) / 3 / iters * 1e6
  return out

res = {}
for name, shape, st in (('205^2', (2, 1, 205, 205), False), ('64^2', (2, 1, 64, 64), False),
                        ('16x102^2', (2, 16, 102, 102), False), ('16x102^2 stitching', (2, 16, 102, 102), True)):
  res[name] = run(shape, stitch=st)
print(json.dumps(res))
