"""Kernel A/B measurement of the flow path: the bench workload (one 4096^2 uint8 tile pair,
patch 160, step 40, batch 1024) on the library named by SOFIMA_B200_LIB, per-kernel device
times from the context's CUDA-event hooks, and a checksum of the outputs so that two builds
can be compared for identical integer flow vectors.  Not a bench number."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from sofima_b200 import _native, flow_field

dev = torch.device('cuda', 0)
ctx = _native.Context.get(0)
tiles = bench.synth_tile_pairs(6, bench.FLOW_TILE, 100, dev)
g = (bench.FLOW_TILE - 120) // 40
oyx = np.array(np.where(np.ones((g, g), bool))).T
job = flow_field._FlowJob(ctx, oyx, (bench.FLOW_TILE,) * 2, (bench.FLOW_TILE,) * 2,
                          (160, 160), (160, 160), (40, 40), 1024)
out_d = torch.empty((len(job.batches), 1024, 4), dtype=torch.float32, device=dev)
step = lambda i: job.run(tiles[i % 6][0], tiles[i % 6][1], out=out_d)
for i in range(4):
  step(i)
torch.cuda.synchronize()
K = 12
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(K):
  step(i)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
ctx.set_timing(True)
for i in range(6):
  step(i)
torch.cuda.synchronize()
rep = ctx.timing_report()
ctx.set_timing(False)
step(0)
pk = out_d.cpu().numpy().reshape(-1, 4)[:g * g]
np.save(os.environ.get('AB_OUT', '/tmp/ab_out.npy'), pk)
print(json.dumps({'lib': os.environ.get('SOFIMA_B200_LIB', 'default'), 'ms_per_step': ms,
                  'pairs_per_s': g * g / (ms * 1e-3),
                  'kernel_ms_per_step': {k: v['ms'] / 6 for k, v in rep.items()},
                  'flow_xy_sum': float(np.nansum(pk[:, :2])),
                  'stats_sum': float(np.nansum(pk[:, 2:]))}))
