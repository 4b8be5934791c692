"""BASELINE config 2, mesh part: 4 x 4 tiles of 4096^2 px, stride 40 -> [2, 16, 102, 102]
tile meshes relaxed with the stitching prev_fn (never a bench.py number; prints timing)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import importlib.util
spec = importlib.util.spec_from_file_location('tsg', os.path.join(os.path.dirname(__file__), '..', 'tests', 'test_stitch_gpu.py'))
tsg = importlib.util.module_from_spec(spec); spec.loader.exec_module(tsg)
from sofima_b200 import mesh, stitch_elastic, _native

fx, fy, x, nbors, stride = tsg._big_case(seed=1, nt_x=4, nt_y=4, mesh_shape=(102, 102))
iters = int(os.environ.get('ITERS', '1000'))
cfg = mesh.IntegrationConfig(dt=0.001, gamma=0., k0=0.01, k=0.1, stride=stride, num_iters=iters,
                             max_iters=iters, stop_v_max=0.0, dt_max=100, prefer_orig_order=True,
                             start_cap=0.1, final_cap=10., remove_drift=True)
prev_fn = stitch_elastic.target_mesh_fn(nbors, fx, fy, stride)
xd = torch.from_numpy(x).cuda()
mesh.relax_mesh(xd, None, cfg, prev_fn=prev_fn)
torch.cuda.synchronize()
ctx = _native.Context.get(0)
t0 = time.perf_counter()
for _ in range(3):
  out, ek, t = mesh.relax_mesh(xd, None, cfg, prev_fn=prev_fn)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 3
nodes = x.shape[1] * x.shape[2] * x.shape[3]
ctx.set_timing(True)
mesh.relax_mesh(xd, None, cfg, prev_fn=prev_fn)
torch.cuda.synchronize()
rep = ctx.timing_report()
ctx.set_timing(False)
print(json.dumps({'workload': f'stitching relax [2,16,102,102], {iters} FIRE steps, prev_fn on device',
                  'us_per_step': dt / iters * 1e6, 'node_updates_per_s': nodes * iters / dt,
                  'kernels': rep}))
