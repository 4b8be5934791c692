"""Full-size runs of BASELINE configs 4 and 5 (SURVEY 8 d) -- measurements for BASELINE.md
section 5, not the bench line.

  python tools/config_runs.py config4 [--sections 64] [--size 8192]
      (or under torch.distributed.run: the section pairs are split into contiguous blocks,
       one per rank; no data-path collective)
  python tools/config_runs.py config5 [--tiles 3] [--size 2048] [--depth 64]

config 4: 3-D EM section alignment, a [64, 8192, 8192] uint8 stack: flow between consecutive
  sections (patch 160, step 40, batch 1024: 40 401 patch pairs per section pair).  The stack
  lives in PINNED HOST memory; every section is uploaded once per rank on a copy stream while
  the previous pair is being correlated (double buffering), so the timed region includes the
  H2D traffic; the H2D-only time of the same sections is reported beside it.
config 5: one LICONN in-plane 3-D stitching problem (BASELINE: 32 of them, independent):
  a 3 x 3 grid of [64, 2048, 2048] uint8 tiles with ~10 % overlap -> compute_flow_map3d
  (patch 80^3, stride 40) -> filters -> elastic_mesh_3d relaxation with the stitching prev_fn
  -> ndimage_warp of every tile.  Per-stage seconds of ONE problem; the 32 problems shard over
  the GPUs without communication.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

PATCH, STEP, BATCH = 160, 40, 1024


def smooth_texture(shape, sigma, seed, dev):
  import torch.nn.functional as F
  g = torch.Generator(device=dev).manual_seed(seed)
  r = int(3 * sigma)
  k = torch.arange(-r, r + 1, device=dev, dtype=torch.float32)
  k = torch.exp(-0.5 * (k / sigma) ** 2)
  k = k / k.sum()
  t = torch.randn((1, 1) + tuple(s + 2 * r for s in shape), device=dev, generator=g)
  if len(shape) == 2:
    t = F.conv2d(F.conv2d(t, k.view(1, 1, 1, -1)), k.view(1, 1, -1, 1))
  else:
    t = F.conv3d(F.conv3d(F.conv3d(t, k.view(1, 1, 1, 1, -1)), k.view(1, 1, 1, -1, 1)),
                 k.view(1, 1, -1, 1, 1))
  t = t[0, 0]
  return (t - t.min()) / (t.max() - t.min()) * 255


def config4(args):
  import torch.distributed as dist
  from sofima_b200 import _native, flow_field
  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local)
  dev = torch.device('cuda', local)
  if world > 1:
    dist.init_process_group('nccl', device_id=dev)
  nsec, n = args.sections, args.size
  npairs = nsec - 1
  per = -(-npairs // world)
  p0, p1 = min(rank * per, npairs), min((rank + 1) * per, npairs)
  # synthetic stack: one texture drifting by a known integer offset per section + noise
  rng = np.random.default_rng(3)
  steps = rng.integers(-2, 3, size=(nsec, 2))
  steps[0] = 0
  drift = np.cumsum(steps, axis=0)                      # (dy, dx) of section z
  m = int(np.abs(drift).max()) + 8
  base = smooth_texture((n + 2 * m, n + 2 * m), 2.0, 3, dev)
  secs = list(range(p0, p1 + 1)) if p1 > p0 else []
  host = torch.empty((max(len(secs), 1), n, n), dtype=torch.uint8, pin_memory=True)
  g = torch.Generator(device=dev).manual_seed(100 + rank)
  for i, z in enumerate(secs):
    dy, dx = int(drift[z, 0]), int(drift[z, 1])
    s = base[m + dy:m + dy + n, m + dx:m + dx + n] + torch.randn((n, n), device=dev, generator=g) * 5
    host[i].copy_(s.clamp(0, 255).to(torch.uint8))
  del base
  torch.cuda.synchronize()

  ctx = _native.Context.get(local)
  gsz = (n - (PATCH - STEP)) // STEP
  oyx = np.array(np.where(np.ones((gsz, gsz), bool))).T
  job = flow_field._FlowJob(ctx, oyx, (n, n), (n, n), (PATCH,) * 2, (PATCH,) * 2, (STEP,) * 2, BATCH)
  nb = len(job.batches)
  out_d = torch.empty((max(p1 - p0, 1), nb, BATCH, 4), dtype=torch.float32, device=dev)
  ring = [torch.empty((n, n), dtype=torch.uint8, device=dev) for _ in range(3)]
  copy_stream = torch.cuda.Stream()
  main = torch.cuda.current_stream()

  def run(compute=True):
    ready = [torch.cuda.Event() for _ in secs]
    freed = [torch.cuda.Event() for _ in secs]
    for i in range(len(secs)):
      with torch.cuda.stream(copy_stream):
        if i >= 3:
          copy_stream.wait_event(freed[i - 3])         # the ring slot is free again
        ring[i % 3].copy_(host[i], non_blocking=True)
        ready[i].record(copy_stream)
      if i >= 1:
        main.wait_event(ready[i - 1])
        main.wait_event(ready[i])
        if compute:
          job.run(ring[(i - 1) % 3], ring[i % 3], out=out_d[i - 1])
        freed[i - 1].record(main)
    if secs:
      freed[-1].record(main)

  def barrier():
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
      torch.cuda.synchronize()

  run()           # warm-up (allocations, row-spectra scratch, twiddle tables)
  barrier()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  run()
  e1.record()
  barrier()
  ms = e0.elapsed_time(e1)
  e0.record()
  run(compute=False)
  e1.record()
  barrier()
  h2d_ms = e0.elapsed_time(e1)
  if world > 1:
    t = torch.tensor([ms, h2d_ms], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, h2d_ms = float(t[0]), float(t[1])
  # parity: the known drift between consecutive sections is recovered
  ok_frac = 1.0
  if p1 > p0:
    pk = out_d.cpu().numpy()
    fr = []
    for i, z in enumerate(range(p0, p1)):
      d = drift[z + 1] - drift[z]          # flow = position in pre - position in post
      flat = np.concatenate([pk[i, b, :len(pos)] for b, pos in enumerate(job.batches)])
      fr.append(np.mean((flat[:, 0] == d[1]) & (flat[:, 1] == d[0])))
    ok_frac = float(np.min(fr))
  if world > 1:
    t = torch.tensor([ok_frac], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    ok_frac = float(t[0])
  pairs = npairs * gsz * gsz
  if rank == 0:
    print(json.dumps({
        'config': f'4: [{nsec}, {n}, {n}] uint8 stack, {npairs} section pairs x {gsz * gsz} patch '
                  f'pairs, patch {PATCH}, step {STEP}, batch {BATCH}; contiguous blocks of section '
                  f'pairs per rank, pinned host stack, double-buffered H2D inside the timed region',
        'n_gpus': world, 'patch_pairs': pairs, 'seconds': ms * 1e-3,
        'patch_pairs_per_s': pairs / (ms * 1e-3),
        'h2d_only_seconds': h2d_ms * 1e-3,
        'h2d_bytes_per_rank': len(secs) * n * n,
        'h2d_gb_per_s_per_rank': len(secs) * n * n / (h2d_ms * 1e-3) / 1e9 if h2d_ms else None,
        'parity': {'check': 'integer flow vector == known section-to-section drift',
                   'min_fraction_exact_over_section_pairs': ok_frac}}))
  assert ok_frac > 0.99, ok_frac
  if world > 1:
    dist.destroy_process_group()


def config5(args):
  """One problem per call of `config5_problem`; with --problems P under torch.distributed.run
  the P independent problems are dealt round-robin to the ranks (no communication) and the
  job time is the maximum over the ranks."""
  import torch.distributed as dist
  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local)
  dev = torch.device('cuda', local)
  if world > 1:
    dist.init_process_group('nccl', device_id=dev)
  mine = list(range(rank, args.problems, world))
  torch.cuda.synchronize()
  if world > 1:
    dist.barrier()
  t0 = time.perf_counter()
  results = [config5_problem(args, dev, seed=5 + 17 * k) for k in mine]
  torch.cuda.synchronize()
  mine_s = time.perf_counter() - t0
  if world > 1:
    t = torch.tensor([mine_s], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    job_s = float(t.item())
  else:
    job_s = mine_s
  if rank == 0:
    res = dict(results[0])
    if args.problems > 1:
      res.update(problems=args.problems, n_gpus=world, job_seconds=job_s,
                 problems_per_rank=len(mine),
                 per_problem_seconds_rank0=[r['problem_seconds'] for r in results],
                 note='job_seconds = max over ranks of the wall time for its problems '
                      '(synthetic tiles are generated outside the per-stage timers but inside '
                      'job_seconds)')
    print(json.dumps(res))
  if world > 1:
    dist.destroy_process_group()


def config5_problem(args, dev, seed=5):
  from sofima_b200 import flow_utils, mesh, stitch_elastic, warp
  nt, n, nz = args.tiles, args.size, args.depth
  ov = int(n * 0.1)
  stepxy = n - ov
  rng = np.random.default_rng(seed)
  big = smooth_texture((nz + 16, (nt - 1) * stepxy + n + 64, (nt - 1) * stepxy + n + 64), 1.5, seed, dev)
  tiles, pos = {}, {}
  for ty in range(nt):
    for tx in range(nt):
      jz, jy, jx = (int(v) for v in rng.integers(-3, 4, 3))
      z0, y0, x0 = 8 + jz, 32 + ty * stepxy + jy, 32 + tx * stepxy + jx
      pos[tx, ty] = (z0, y0, x0)
      tiles[tx, ty] = big[z0:z0 + nz, y0:y0 + n, x0:x0 + n].to(torch.uint8).cpu().numpy()[None]
  del big
  torch.cuda.empty_cache()
  cx = np.full((3, 1, nt, nt), np.nan)
  cy = np.full((3, 1, nt, nt), np.nan)
  for (tx, ty), (z0, y0, x0) in pos.items():
    if (tx + 1, ty) in pos:
      z1, y1, x1 = pos[tx + 1, ty]
      cx[:, 0, ty, tx] = (x1 - x0 - n, y1 - y0, z1 - z0)
    if (tx, ty + 1) in pos:
      z1, y1, x1 = pos[tx, ty + 1]
      cy[:, 0, ty, tx] = (x1 - x0, y1 - y0 - n, z1 - z0)
  stride = (40, 40, 40)
  tile_size_xyz = (n, n, nz)
  res = {'config': f'5 (one of 32 problems): {nt} x {nt} tiles of [{nz}, {n}, {n}] uint8, '
                   'compute_flow_map3d patch 80^3 stride 40 -> clean/reconcile -> aggregate -> '
                   'elastic_mesh_3d relaxation with the stitching prev_fn -> ndimage_warp'}
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  flow_x, offsets_x = stitch_elastic.compute_flow_map3d(tiles, tile_size_xyz, cx, axis=0,
                                                        stride=stride, patch_size=(80, 80, 80))
  flow_y, offsets_y = stitch_elastic.compute_flow_map3d(tiles, tile_size_xyz, cy, axis=1,
                                                        stride=stride, patch_size=(80, 80, 80))
  torch.cuda.synchronize()
  res['flow_seconds'] = time.perf_counter() - t0
  res['patch_pairs'] = int(sum(np.prod(v.shape[1:]) for v in list(flow_x.values()) +
                               list(flow_y.values())))
  res['patch_pairs_per_s'] = res['patch_pairs'] / res['flow_seconds']
  t0 = time.perf_counter()
  kw = dict(min_peak_ratio=1.4, min_peak_sharpness=1.4, max_deviation=5, max_magnitude=0, dim=3)
  fine_x = {k: flow_utils.clean_flow(v, **kw) for k, v in flow_x.items()}
  fine_y = {k: flow_utils.clean_flow(v, **kw) for k, v in flow_y.items()}
  kw = dict(min_patch_size=10, max_gradient=-1, max_deviation=-1)
  fine_x = {k: flow_utils.reconcile_flows([v], **kw) for k, v in fine_x.items()}
  fine_y = {k: flow_utils.reconcile_flows([v], **kw) for k, v in fine_y.items()}
  res['filter_seconds'] = time.perf_counter() - t0
  valid = np.mean([np.isfinite(v[0]).mean() for v in list(fine_x.values()) + list(fine_y.values())])
  res['valid_flow_fraction'] = float(valid)
  coarse = np.zeros((3, nt, nt))
  fx, fy, x0m, nbors, key_to_idx = stitch_elastic.aggregate_arrays(
      (cx[:, 0], fine_x, offsets_x), (cy[:, 0], fine_y, offsets_y), list(tiles.keys()), coarse,
      stride=stride, tile_shape=tile_size_xyz[::-1])
  cfg = mesh.IntegrationConfig(dt=0.001, gamma=0., k0=0.01, k=0.1, stride=stride,
                               num_iters=1000, max_iters=args.mesh_max_iters, stop_v_max=0.001,
                               dt_max=100, prefer_orig_order=False, start_cap=0.1, final_cap=10.,
                               remove_drift=True)
  prev_fn = stitch_elastic.target_mesh_fn(nbors, fx, fy, stride)
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  x, e_kin, steps = mesh.relax_mesh(np.asarray(x0m, np.float32), None, cfg,
                                    mesh_force=mesh.elastic_mesh_3d, prev_fn=prev_fn)
  torch.cuda.synchronize()
  res['mesh_seconds'] = time.perf_counter() - t0
  nodes = int(np.prod(x.shape[1:]))
  res.update(mesh_nodes=nodes, mesh_steps=int(steps),
             mesh_node_updates_per_s=nodes * steps / res['mesh_seconds'],
             mesh_us_per_step=res['mesh_seconds'] / max(steps, 1) * 1e6,
             mesh_converged=bool(steps < cfg.max_iters))
  t0 = time.perf_counter()
  nvox = 0
  for key, idx in key_to_idx.items():
    w = warp.ndimage_warp(torch.from_numpy(tiles[key][0]).to(dev), x[:, idx], stride,
                          (256, 256, 64), (0, 0, 0))
    nvox += int(np.prod(tiles[key].shape))
  torch.cuda.synchronize()
  res['warp_seconds'] = time.perf_counter() - t0
  res['warp_voxels_per_s'] = nvox / res['warp_seconds']
  res['problem_seconds'] = (res['flow_seconds'] + res['filter_seconds'] + res['mesh_seconds'] +
                            res['warp_seconds'])
  if args.render:
    # the notebook's last step: StitchAndRender3dTiles over 512^3 boxes of the stitched volume
    from sofima_b200 import compat
    from sofima_b200.processor import warp as pwarp
    tile_ids = {key: 100 + idx for key, idx in key_to_idx.items()}
    by_id = {tid: tiles[key][0] for key, tid in tile_ids.items()}

    class Renderer(pwarp.StitchAndRender3dTiles):
      def _open_tile_volume(self, tile_id):
        return by_id[tile_id]

    Renderer.reset_cache()
    tile_map = [[tile_ids[tx, ty] for tx in range(nt)] for ty in range(nt)]
    r = Renderer(tile_map=tile_map, tile_pattern_path='{tile_id}', stride=stride,
                 tile_mesh_path={'key_to_idx': key_to_idx, 'x': np.asarray(x)},
                 offset=(0, 0, 0))
    full = compat.BoundingBox(start=(0, 0, 0), size=(nt * stepxy, nt * stepxy, nz))
    stitched = np.zeros(full.size[::-1], np.uint8)
    t0 = time.perf_counter()
    for y0 in range(0, int(full.size[1]), 512):
      for x0 in range(0, int(full.size[0]), 512):
        b = compat.BoundingBox(start=(x0, y0, 0), end=(min(x0 + 512, full.size[0]),
                                                       min(y0 + 512, full.size[1]), nz))
        sv = r.process(compat.Subvolume(np.zeros((1,) + tuple(b.size[::-1]), np.uint8), b))
        stitched[sv.bbox.to_slice3d()] = sv.data[0]
    torch.cuda.synchronize()
    res['render_seconds'] = time.perf_counter() - t0
    res['render_voxels_per_s'] = stitched.size / res['render_seconds']
    res['render_filled_fraction'] = float((stitched > 0).mean())
    res['problem_seconds'] += res['render_seconds']
    Renderer.reset_cache()
  res['extrapolated_32_problems_on_8_gpus_seconds'] = res['problem_seconds'] * 32 / 8
  return res


if __name__ == '__main__':
  ap = argparse.ArgumentParser()
  ap.add_argument('which', choices=['config4', 'config5'])
  ap.add_argument('--sections', type=int, default=64)
  ap.add_argument('--size', type=int, default=None)
  ap.add_argument('--tiles', type=int, default=3)
  ap.add_argument('--depth', type=int, default=64)
  ap.add_argument('--mesh-max-iters', type=int, default=20000)
  ap.add_argument('--render', action='store_true')
  ap.add_argument('--problems', type=int, default=1)
  a = ap.parse_args()
  if a.which == 'config4':
    a.size = a.size or 8192
    config4(a)
  else:
    a.size = a.size or 2048
    config5(a)
