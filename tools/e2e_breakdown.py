"""Where the host-array (e2e) flow_field call spends its time: image upload, index tables,
kernels, read-back + scatter.  Diagnostic, not a bench number."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from sofima_b200 import _native, flow_field

dev = torch.device('cuda', 0)
ctx = _native.Context.get(0)
pre, post, _ = bench.synth_tile_pairs(1, bench.FLOW_TILE, 100, dev)[0]
hp = torch.empty(pre.shape, dtype=torch.uint8, pin_memory=True); hp.copy_(pre)
hq = torch.empty(post.shape, dtype=torch.uint8, pin_memory=True); hq.copy_(post)
hp, hq = hp.numpy(), hq.numpy()
calc = flow_field.JAXMaskedXCorrWithStatsCalculator()
kw = dict(patch_size=160, step=40, batch_size=1024)
for _ in range(3):
  calc.flow_field(hp, hq, **kw)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
  calc.flow_field(hp, hq, **kw)
whole = (time.perf_counter() - t0) / 10 * 1e3
t0 = time.perf_counter()
for _ in range(10):
  a = flow_field._device_image(hp, ctx); b = flow_field._device_image(hq, ctx)
  torch.cuda.synchronize()
h2d = (time.perf_counter() - t0) / 10 * 1e3
g = (bench.FLOW_TILE - 120) // 40
oyx = np.array(np.where(np.ones((g, g), bool))).T
t0 = time.perf_counter()
for _ in range(10):
  job = flow_field._FlowJob(ctx, oyx, hp.shape, hq.shape, (160, 160), (160, 160), (40, 40), 1024)
  torch.cuda.synchronize()
tables = (time.perf_counter() - t0) / 10 * 1e3
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
  pk = job.run(a, b)
  torch.cuda.synchronize()
run = (time.perf_counter() - t0) / 10 * 1e3
out = np.full((4, g, g), np.nan, np.float32)
t0 = time.perf_counter()
for _ in range(10):
  job.scatter(pk.cpu().numpy(), out)
back = (time.perf_counter() - t0) / 10 * 1e3
print(json.dumps({'flow_field_ms': whole, 'h2d_ms': h2d, 'index_tables_ms': tables,
                  'kernels_ms': run, 'd2h_scatter_ms': back}))
