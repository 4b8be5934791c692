#!/usr/bin/env python
"""Benchmark of the two SOFIMA hot paths on B200 (see BASELINE.json / DESIGN.md).

  python bench.py --gpus N --steps K --warmup W           # this repo (CUDA)
  python bench.py --impl reference --gpus N ...            # CPU arm (oracle port)

One JSON line on rank 0.  Primary metric: patch-pairs/s of the flow estimator; the
`mesh` object carries the second BASELINE metric (node-updates/s of the mesh
solver) with its own roofline / e2e / cpu_baseline.

A *step* is one pass of the hot path over one batch of synthetic input:
  flow  one 4096x4096 uint8 tile pair, patch 160, step 40 -> 9801 patch pairs
        (BASELINE configs[1] tile size; EM-2D batch_size 1024, em_2d.py:32-41)
  mesh  one velocity_verlet chunk of `--mesh-iters` (default 1000) FIRE steps on a
        [2, 1, 2048, 2048] mesh with a fixed `prev` (BASELINE configs[2]);
        node-updates = nodes x steps.
`value` is timed with the inputs resident in HBM; `e2e` goes through the public
Python API (sofima_b200.flow_field / sofima_b200.mesh, i.e. the C-ABI library)
with pinned HOST arrays, including H2D of the inputs and D2H of the result.
Between timed steps the inputs rotate over tile pairs totalling more than the
126 MB L2 (flow) / the mesh state itself exceeds L2 (134 MB).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

FLOW_TILE = 4096
PATCH, STEP, BATCH = 160, 40, 1024
MESH_N = 2048
DENSE_FLOP_PER_PAIR = 2.0 * PATCH**4          # p^4 MAC per patch pair
MESH_BYTES_PER_UPDATE = 56.0                  # 8 floats in, 6 out


def _step_traffic(which):
  """DRAM bytes of every kernel of one bench step (tools/step_traffic.py: an `ncu --metrics
  dram__bytes_*` pass over the same workload, committed under profiles/).  DRAM counters
  cannot be read from inside a timed run, so this is the one number of the line that is
  not measured in it; `source` says where it comes from."""
  path = os.path.join(ROOT, 'profiles', f'ncu_r2_{which}_step.json')
  try:
    d = json.load(open(path))
    return d, os.path.relpath(path, ROOT)
  except (OSError, ValueError):
    return None, None


def _peaks():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  try:
    d = json.load(open(path))
    return dict(hbm_gbs=float(d['hbm_gbs']), tflops=float(d['bf16_tflops_sustained']),
                source='measured (MEASURED_PEAKS.json)')
  except (OSError, KeyError, ValueError, TypeError):
    return dict(hbm_gbs=6650.0, tflops=1590.0,
                source='fallback (B200_PROFILING.md)')


def flow_kernel_rooflines(kern_ms, pairs, tile, patch, step, hbm_gbs):
  """HBM view of the flow kernels of one step: algorithmic bytes
  of each stage (what it has to read and write once, per patch pair or per tile pair)
  over its measured device time.  kern_ms: {timer name: ms per step}."""
  l = 2 * patch - 1                      # correlation image edge (319)
  nkx = (l + 1) // 2 + 1                 # half-spectrum bins of the 320-point rows (161)
  spec = l * nkx * 8                     # product spectra of one pair, complex64
  image = l * l * 4                      # correlation image of one pair, fp32
  rows = 2 * patch * nkx * 8             # cached row spectra one pair reads (both patches)
  nx = (tile - (patch - step)) // step   # distinct patch x starts (99)
  alg = {
      'flow_cols': pairs * (rows + spec),
      'flow_rows_inv': pairs * (spec + image),
      'flow_rowspec': 2 * (tile * tile + nx * tile * nkx * 8),
  }
  out = {}
  for name, nbytes in alg.items():
    ms = kern_ms.get(name)
    if ms:
      gbs = nbytes / (ms * 1e-3) / 1e9
      out[name] = {'bound': 'hbm', 'algorithmic_bytes_per_step': int(nbytes),
                   'achieved': gbs, 'peak': hbm_gbs, 'unit': 'GB/s', 'frac': gbs / hbm_gbs}
  return out


# ----------------------------------------------------------------------------------
# synthetic data
# ----------------------------------------------------------------------------------
def synth_tile_grid(nt_x, nt_y, size, overlap=0.1, seed=1):
  """nt_x x nt_y uint8 tiles cut from one texture with ~`overlap` nominal overlap and
  +-20 px jitter, plus the coarse offset maps compute_flow_map expects."""
  import scipy.ndimage as ndi
  rng = np.random.default_rng(seed)
  step = int(size * (1 - overlap))
  big = (nt_y - 1) * step + size + 64, (nt_x - 1) * step + size + 64
  tex = ndi.gaussian_filter(rng.standard_normal(big).astype(np.float32), 2.0)
  tex = ((tex - tex.min()) / (tex.max() - tex.min()) * 255).astype(np.uint8)
  pos = {(tx, ty): (ty * step + 32 + int(rng.integers(-20, 21)),
                    tx * step + 32 + int(rng.integers(-20, 21)))
         for tx in range(nt_x) for ty in range(nt_y)}
  tiles = {k: np.ascontiguousarray(tex[y0:y0 + size, x0:x0 + size]) for k, (y0, x0) in pos.items()}
  cxm = np.full((2, nt_y, nt_x), np.nan)
  cym = np.full((2, nt_y, nt_x), np.nan)
  for (tx, ty), (y0, x0) in pos.items():
    if (tx + 1, ty) in pos:
      y1, x1 = pos[tx + 1, ty]
      cxm[:, ty, tx] = (x1 - x0 - size, y1 - y0)
    if (tx, ty + 1) in pos:
      y1, x1 = pos[tx, ty + 1]
      cym[:, ty, tx] = (x1 - x0, y1 - y0 - size)
  return tiles, cxm, cym


def synth_stitch(nt_x, nt_y, mesh_shape, stride=(40.0, 40.0), seed=1):
  """Flow fields, coarse offsets and the neighbour table of an nt_x x nt_y tile grid
  with ~10 % overlap (3-5 flow columns / rows per seam), through aggregate_arrays."""
  import scipy.ndimage as ndi
  from sofima_b200 import stitch_elastic
  rng = np.random.default_rng(seed)
  my, mx = mesh_shape
  coords = [(tx, ty) for ty in range(nt_y) for tx in range(nt_x)]
  cx = np.full((2, nt_y, nt_x), np.nan)
  cy = np.full((2, nt_y, nt_x), np.nan)
  fine_x, fine_y, off_x, off_y = {}, {}, {}, {}

  def smooth(shape, amp):
    return (ndi.gaussian_filter(rng.standard_normal(shape), (0, 2, 2)) * amp).astype(np.float32)

  for tx, ty in coords:
    if tx + 1 < nt_x:
      oy, ox = my - int(rng.integers(0, 3)), int(rng.integers(3, 6))
      f = np.full((4, oy, ox), 1.0, np.float32)
      f[:2] = smooth((2, oy, ox), 5.0)
      fine_x[tx, ty] = f
      cx[:, ty, tx] = (mx * stride[1] - ox * stride[1] + rng.integers(-9, 9),
                       rng.integers(-50, 50))
      off_x[tx, ty] = (int(rng.integers(-4, 4)), int(rng.integers(-4, 4)))
    if ty + 1 < nt_y:
      oy, ox = int(rng.integers(3, 6)), mx - int(rng.integers(0, 3))
      f = np.full((4, oy, ox), 1.0, np.float32)
      f[:2] = smooth((2, oy, ox), 5.0)
      fine_y[tx, ty] = f
      cy[:, ty, tx] = (rng.integers(-50, 50),
                       my * stride[0] - oy * stride[0] + rng.integers(-9, 9))
      off_y[tx, ty] = (int(rng.integers(-4, 4)), int(rng.integers(-4, 4)))
  coarse = rng.standard_normal((2, nt_y, nt_x)) * 3
  fx, fy, x, nbors, _ = stitch_elastic.aggregate_arrays(
      (cx, fine_x, off_x), (cy, fine_y, off_y), coords, coarse, stride,
      (my * int(stride[0]), mx * int(stride[1])))
  return fx.astype(np.float32), fy.astype(np.float32), x.astype(np.float32), nbors, stride




def synth_tile_pairs(num, size, seed, device):
  """`num` uint8 tile pairs cut from smooth random textures with a known shift."""
  import torch
  import torch.nn.functional as F
  g = torch.Generator(device=device).manual_seed(seed)
  m = 32
  k = torch.arange(-6, 7, device=device, dtype=torch.float32)
  k = torch.exp(-0.5 * (k / 2.0) ** 2)
  k = (k / k.sum()).view(1, 1, 1, -1)
  pairs = []
  for i in range(num):
    base = torch.randn((1, 1, size + 2 * m + 12, size + 2 * m + 12), device=device,
                       generator=g)
    base = F.conv2d(F.conv2d(base, k), k.transpose(2, 3))[0, 0]
    base = (base - base.min()) / (base.max() - base.min()) * 255
    dy, dx = 3 + i % 3, -4 + i % 5
    pre = base[m:m + size, m:m + size].to(torch.uint8).contiguous()
    noise = torch.randn((size, size), device=device, generator=g) * 5
    post = (base[m + dy:m + dy + size, m + dx:m + dx + size] + noise).clamp(0, 255)
    pairs.append((pre, post.to(torch.uint8).contiguous(), (dy, dx)))
  return pairs


def synth_mesh(n, seed, device):
  """config 3: x0 = 0, prev = smooth displacement field (max 8 px), 1 % NaN."""
  import torch
  import torch.nn.functional as F
  g = torch.Generator(device=device).manual_seed(seed)
  k = torch.arange(-32, 33, device=device, dtype=torch.float32)
  k = torch.exp(-0.5 * (k / 16.0) ** 2)
  k = (k / k.sum()).view(1, 1, 1, -1)
  f = torch.randn((2, 1, n + 64, n + 64), device=device, generator=g)
  f = F.conv2d(F.conv2d(f, k), k.transpose(2, 3))
  f = f / f.abs().max() * 8.0
  prev = f.view(2, 1, n, n).contiguous()
  nan = torch.rand((1, 1, n, n), device=device, generator=g) < 0.01
  prev = torch.where(nan.expand_as(prev), torch.full_like(prev, float('nan')), prev)
  return prev.contiguous()


def mesh_config(mesh, iters):
  return mesh.IntegrationConfig(
      dt=0.001, gamma=0.0, k0=0.1, k=0.1, stride=(40.0, 40.0), num_iters=iters,
      max_iters=iters, stop_v_max=0.0, fire=True, dt_max=1000.0,
      prefer_orig_order=True)


# ----------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------


class ClockSampler:
  """Samples nvidia-smi clocks / throttle reasons during the timed region."""

  Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
       'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
       'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

  def __init__(self, gpu_index):
    self.gpu = gpu_index
    self.rows = []
    self.proc = None

  def __enter__(self):
    try:
      self.proc = subprocess.Popen(
          ['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
           '-i', str(self.gpu), '-lms', '100'], stdout=subprocess.PIPE,
          stderr=subprocess.DEVNULL, text=True)
      self.thread = threading.Thread(target=self._read, daemon=True)
      self.thread.start()
    except OSError:
      self.proc = None
    return self

  def _read(self):
    for line in self.proc.stdout:
      self.rows.append([c.strip() for c in line.split(',')])

  def __exit__(self, *exc):
    if self.proc:
      time.sleep(0.15)
      self.proc.terminate()
      try:
        self.proc.wait(timeout=2)
      except subprocess.TimeoutExpired:
        self.proc.kill()

  def summary(self):
    sm, mx, reasons = [], [], set()
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    for r in self.rows:
      try:
        sm.append(float(r[1]))
        mx.append(float(r[2]))
      except (ValueError, IndexError):
        continue
      for name, val in zip(names, r[5:9]):
        if val.lower().startswith('active'):
          reasons.add(name)
    if not sm:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
    return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)),
            'reasons': sorted(reasons), 'samples': len(sm)}


# ----------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores
# ----------------------------------------------------------------------------------


def cpu_flow_sample(size, seed=0):
  import scipy.ndimage as ndi
  rng = np.random.default_rng(seed)
  base = ndi.gaussian_filter(rng.standard_normal((size + 64, size + 64)), 2.0)
  base = ((base - base.min()) / (base.max() - base.min()) * 255).astype(np.uint8)
  pre = np.ascontiguousarray(base[32:32 + size, 32:32 + size])
  post = np.ascontiguousarray(base[35:35 + size, 28:28 + size])
  return pre, post


class CpuFlowArm:
  """The reference's unit of work on the CPU: one `batched_xcorr_peaks` call
  (flow_field.py:385-441, the jit boundary) on a whole reference batch of BATCH = 1024
  patch pairs of the SAME 4096 x 4096 tile pair, patch 160, step 40, that the GPU arm
  runs -- oracle/flow_oracle.py (NumPy + pocketfft fp32, all host threads)."""

  def __init__(self, size=FLOW_TILE):
    from oracle import flow_oracle
    self.fo = flow_oracle
    self.pre, self.post = cpu_flow_sample(size)
    g = (size - (PATCH - STEP)) // STEP
    oyx = np.array(np.where(np.ones((g, g), bool))).T
    self.batches = []
    for i in range(0, len(oyx), BATCH):       # flow_field.py:610-623
      pos = oyx[i:i + BATCH]
      real = pos.shape[0]
      if real < BATCH:
        pos = np.pad(pos, ((0, BATCH - real), (0, 0)), mode='edge')
      self.batches.append((pos * STEP, real))

  def step(self, i):
    """One reference batch; returns (patch pairs computed, seconds)."""
    starts, real = self.batches[i % len(self.batches)]
    t0 = time.perf_counter()
    peaks = self.fo.batched_xcorr_peaks(self.pre, self.post, None, None, (PATCH, PATCH),
                                        starts, None, post_starts=starts)
    sec = time.perf_counter() - t0
    assert peaks.shape == (BATCH, 4)
    ok = peaks[:real]
    assert np.all(ok[:, 0] == -4) and np.all(ok[:, 1] == 3), 'cpu arm parity'
    return BATCH, sec


def time_cpu_mesh(n, iters, reps=1):
  """C restatement (OpenMP, all cores) of one velocity_verlet chunk on n x n nodes."""
  from oracle import mesh_oracle_c
  from sofima_b200 import mesh
  rng = np.random.default_rng(2)
  prev = (rng.standard_normal((2, 1, n, n)) * 4).astype(np.float32)
  x = np.zeros_like(prev)
  cfg = mesh_config(mesh, iters)
  best = float('inf')
  for _ in range(reps):
    t0 = time.perf_counter()
    mesh_oracle_c.velocity_verlet(x, np.zeros_like(x), prev, cfg, cfg.start_cap)
    best = min(best, time.perf_counter() - t0)
  return n * n * iters / best, best, mesh_oracle_c.num_threads()


def run_reference(args):
  """`--impl reference`: the reference algorithm on the host cores.

  The reference is pure Python/JAX and JAX is not installable in this image, so the
  CPU arm is the oracle port (oracle/flow_oracle.py on pocketfft with all cores,
  oracle/mesh_oracle.c with OpenMP) -- `kind: "port"`.
  """
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  cores = len(os.sched_getaffinity(0))
  arm = CpuFlowArm()
  for i in range(args.warmup):
    arm.step(i)
  pairs, flow_s = 0, 0.0
  for i in range(args.steps):
    n, sec = arm.step(args.warmup + i)
    pairs += n
    flow_s += sec
  flow_v = pairs / flow_s
  mesh_iters = 10
  for _ in range(min(args.warmup, 1)):
    time_cpu_mesh(MESH_N, mesh_iters)
  mesh_s = 0.0
  for _ in range(args.steps):
    _, sec, threads = time_cpu_mesh(MESH_N, mesh_iters)
    mesh_s += sec
  mesh_v = MESH_N * MESH_N * mesh_iters * args.steps / mesh_s
  line = {
      'impl': 'reference', 'metric': 'patch-pairs/s', 'value': flow_v,
      'unit': 'patch-pairs/s', 'n_gpus': args.gpus, 'steps': args.steps,
      'warmup': args.warmup, 'ms_per_step': flow_s / args.steps * 1e3,
      'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
      'dtype': 'f32', 'data': 'synthetic',
      'config': {'workload': f'flow_field on one {FLOW_TILE}x{FLOW_TILE} uint8 tile pair, patch '
                             f'{PATCH}, step {STEP}, batch {BATCH}; a CPU step is ONE reference '
                             f'batch of {BATCH} patch pairs of that tile pair (bounded sample '
                             'of the same workload, same tile size and batch size)'},
      'cpu_baseline': {'value': flow_v, 'unit': 'patch-pairs/s', 'cores': cores,
                       'kind': 'port',
                       'sample': f'{args.steps} reference batches of {BATCH} patch pairs of the '
                                 f'{FLOW_TILE}^2 tile pair, oracle/flow_oracle.py '
                                 '(batched_xcorr_peaks), pocketfft fp32, '
                                 f'OMP_NUM_THREADS={os.environ.get("OMP_NUM_THREADS")}'},
      'e2e': {'value': flow_v, 'unit': 'patch-pairs/s', 'h2d_bytes_per_step': 0,
              'd2h_bytes_per_step': 0},
      'mesh': {'metric': 'node-updates/s', 'value': mesh_v, 'unit': 'node-updates/s',
               'ms_per_step': mesh_s / args.steps * 1e3,
               'cpu_baseline': {'value': mesh_v, 'unit': 'node-updates/s',
                                'cores': threads, 'kind': 'port',
                                'sample': f'{args.steps} x {mesh_iters} FIRE steps on '
                                          f'{MESH_N}^2 nodes, oracle/mesh_oracle.c'},
               'e2e': {'value': mesh_v, 'unit': 'node-updates/s',
                       'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}},
  }
  emit(line)


# ----------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------


def run_ours(args):
  import torch
  import torch.distributed as dist
  from sofima_b200 import _native, flow_field, mesh

  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  if not torch.cuda.is_available():
    raise SystemExit('bench.py needs a B200: sofima_b200 has no CPU fallback '
                     '(use --impl reference for the CPU arm)')
  torch.cuda.set_device(local)
  dev = torch.device('cuda', local)
  if world > 1:
    dist.init_process_group('nccl', device_id=dev)
  ctx = _native.Context.get(local)
  peaks = _peaks()
  K, W = args.steps, args.warmup

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def max_over_ranks(ms):
    if world > 1:
      t = torch.tensor([ms], device=dev, dtype=torch.float64)
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
      return float(t.item())
    return ms

  def timed(fn, steps):
    """Times `steps` calls of fn(i) on the device; barrier + sync on both sides."""
    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    l0 = ctx.launch_count
    e0.record()
    for i in range(steps):
      fn(i)
    e1.record()
    barrier()
    return max_over_ranks(e0.elapsed_time(e1)), ctx.launch_count - l0

  result = {}
  clocks = {}

  # ---------------- flow ----------------
  if args.path in ('both', 'flow'):
    npairs_tiles = 6  # 6 x (2 x 16.8 MB) = 201 MB of distinct inputs > 126 MB L2
    tiles = synth_tile_pairs(npairs_tiles, FLOW_TILE, 100 + rank, dev)
    calc = flow_field.JAXMaskedXCorrWithStatsCalculator()
    g = (FLOW_TILE - (PATCH - STEP)) // STEP
    oyx = np.array(np.where(np.ones((g, g), bool))).T
    job = flow_field._FlowJob(ctx, oyx, (FLOW_TILE,) * 2, (FLOW_TILE,) * 2,
                              (PATCH,) * 2, (PATCH,) * 2, (STEP,) * 2, BATCH)
    out_d = torch.empty((len(job.batches), BATCH, 4), dtype=torch.float32, device=dev)

    def flow_step(i):
      pre, post, _ = tiles[i % npairs_tiles]
      job.run(pre, post, out=out_d)

    with ClockSampler(local) as cs:  # sampled under load: warm-up + timed region
      for i in range(W):
        flow_step(i)
      ms, launches = timed(flow_step, K)
    clocks['flow'] = cs.summary()
    pairs_per_step = g * g
    flow_value = world * pairs_per_step * K / (ms * 1e-3)

    # parity spot check of the timed configuration: the known shift is recovered.
    pk = out_d.cpu().numpy()
    dy, dx = tiles[(K - 1) % npairs_tiles][2]
    flat = pk.reshape(-1, 4)[:pairs_per_step]
    assert np.all(flat[:, 0] == dx) and np.all(flat[:, 1] == dy), 'flow parity'

    # per-kernel durations (separate pass, CUDA events on the launching stream)
    ctx.set_timing(True)
    flow_step(0)
    rep = ctx.timing_report()
    ctx.set_timing(False)
    kern_ms = {k: v['ms'] for k, v in rep.items() if k.startswith('flow_')}
    tot_ms = sum(kern_ms.values())
    dom = max(kern_ms, key=kern_ms.get)
    achieved_tf = pairs_per_step * DENSE_FLOP_PER_PAIR / (tot_ms * 1e-3) / 1e12
    hbm_view = flow_kernel_rooflines(kern_ms, pairs_per_step, FLOW_TILE, PATCH, STEP,
                                     peaks['hbm_gbs'])
    traffic, traffic_src = _step_traffic('flow')
    ncu_k = {}
    if traffic:
      for name, a in traffic['kernels'].items():
        ncu_k[name] = {'dram_bytes_per_step': a['dram_read_bytes'] + a['dram_write_bytes'],
                       'issue_active_pct': round(a['issue_active_pct_avg'], 1),
                       'tensor_pipe_pct': round(a['tensor_pipe_pct_max'], 1)}

    # sustained: the same step repeated for at least two seconds (clocks settle under load)
    n_sust = max(K, int(np.ceil(2100.0 / (ms / K))))  # 5 % margin: never under 2 s
    with ClockSampler(local) as cs2:
      ms_sust, _ = timed(flow_step, n_sust)
    sustained = {'value': world * pairs_per_step * n_sust / (ms_sust * 1e-3),
                 'unit': 'patch-pairs/s', 'steps': n_sust, 'ms_per_step': ms_sust / n_sust,
                 'seconds': ms_sust * 1e-3, 'clocks': cs2.summary()}

    # e2e: public API with pinned host arrays.
    host = []
    for pre, post, _ in tiles[:3]:
      hp = torch.empty(pre.shape, dtype=torch.uint8, pin_memory=True)
      hq = torch.empty(post.shape, dtype=torch.uint8, pin_memory=True)
      hp.copy_(pre)
      hq.copy_(post)
      host.append((hp.numpy(), hq.numpy()))
    torch.cuda.synchronize()

    def flow_e2e(i):
      a, b = host[i % len(host)]
      out = calc.flow_field(a, b, PATCH, STEP, batch_size=BATCH)
      assert out.shape == (4, g, g)

    flow_e2e(0)
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
      flow_e2e(i)
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    result.update({
        'metric': 'patch-pairs/s', 'value': flow_value, 'unit': 'patch-pairs/s',
        'ms_per_step': ms / K, 'gpu_launches': launches,
        'sustained': sustained,
        'roofline': {
            # the dominant kernel against the resource the contract lets it be compared with;
            # what actually binds it is instruction issue (see binding_resource)
            'bound': 'hbm', 'kernel': dom,
            'achieved': hbm_view.get(dom, {}).get('achieved'),
            'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
            'frac': hbm_view.get(dom, {}).get('frac'),
            'algorithmic_bytes_per_launch': (hbm_view.get(dom, {}).get(
                'algorithmic_bytes_per_step', 0) // max(1, len(job.batches))),
            'traffic': traffic['total']['dram_bytes'] if traffic else None,
            'traffic_note': 'sum of dram__bytes_read + dram__bytes_write of ALL flow kernels of '
                            'one step (ncu metrics pass), beside algorithmic_io_bytes_per_step = '
                            'the two uint8 tiles in and the peak table out',
            'traffic_source': traffic_src,
            'algorithmic_io_bytes_per_step': 2 * FLOW_TILE * FLOW_TILE + pairs_per_step * 16,
            'binding_resource': 'instruction issue and latency: the fp32 FFT kernels run at '
                                '54-66 % issue-slot utilisation with the HBM and tensor pipes '
                                'far from their limits (ncu_per_kernel); the tensor cores carry '
                                'the row-spectra GEMM only (rowspec_tc_kernel)',
            'ncu_per_kernel': ncu_k,
            'peak_source': peaks['source'],
            'dominant_kernel': dom, 'kernel_ms_per_step': kern_ms,
            'kernels_hbm_view': hbm_view,
            'dense_equivalent': {
                'achieved': achieved_tf, 'peak': peaks['tflops'], 'unit': 'TFLOP/s',
                'frac': achieved_tf / peaks['tflops'],
                'note': 'side figure, NOT a hardware fraction: 2*160^4 FLOP per patch pair '
                        '(SURVEY 8d) over the summed device time of the flow kernels; the '
                        'executed work is ~13 MFLOP/pair of fp32 FFT on the CUDA cores plus '
                        'the exact int8 row-spectra GEMM on the tensor cores'}},
        'e2e': {'value': world * pairs_per_step * K / (e2e_ms * 1e-3),
                'unit': 'patch-pairs/s',
                'h2d_bytes_per_step': 2 * FLOW_TILE * FLOW_TILE + int(job.starts_d.numel()) * 4,
                'd2h_bytes_per_step': int(out_d.numel()) * 4},
    })
    del tiles, host

  # ---------------- mesh ----------------
  if args.path in ('both', 'mesh'):
    from sofima_b200 import mesh_sharded
    iters = args.mesh_iters
    cfg = mesh_config(mesh, iters)
    nodes = MESH_N * MESH_N
    # N > 1: ONE 2048^2 mesh, rows sharded over the ranks (strong scaling); every
    # rank synthesises the same field and keeps its slab.
    prev_full = synth_mesh(MESH_N, 7, dev)
    parts = mesh_sharded.partition_rows(MESH_N, world)
    y0, y1 = parts[rank]
    prev = prev_full[:, :, y0:y1].contiguous()
    if not (world > 1 and rank == 0):
      del prev_full  # rank 0 keeps it for the sharded-vs-single-GPU parity check below
    x0 = torch.zeros_like(prev)
    state = {'dt': cfg.dt, 'alpha': cfg.alpha, 'cap': cfg.start_cap}
    if world == 1:
      chunk = mesh._Chunk(x0, None, prev, cfg, 0)
    else:
      chunk = mesh_sharded.ShardedMesh(x0, prev, cfg)

    def mesh_step(i):
      dt, alpha, _, cap, _, _ = chunk.run(state['dt'], state['alpha'], state['cap'])
      state.update(dt=float(dt), alpha=float(alpha), cap=float(cap))

    with ClockSampler(local) as cs:
      for i in range(W):
        mesh_step(i)
      ms, launches = timed(mesh_step, K)
    clocks['mesh'] = cs.summary()
    mesh_value = nodes * iters * K / (ms * 1e-3)  # one global mesh at every N

    ctx.set_timing(True)
    mesh_step(0)
    rep = ctx.timing_report()
    ctx.set_timing(False)
    step_ms = max_over_ranks(rep['mesh_step']['ms'] / rep['mesh_step']['n'])
    step_ms_events = step_ms
    # The per-launch events serialise what the solver overlaps: on one GPU the one-block kernel
    # that adds the FIRE partial sums runs behind the step with programmatic dependent launch
    # and the next step's loads start under it (events between the launches switch that off);
    # a sharded step kernel also waits for its neighbours' flags, and events around a single
    # launch then include the skew between the ranks.  A launch can never take longer than
    # the step of the timed region, which is what bounds it here.
    step_ms = min(step_ms, ms / K / iters)
    local_nodes = (y1 - y0) * MESH_N
    achieved = local_nodes * MESH_BYTES_PER_UPDATE / (step_ms * 1e-3) / 1e9
    mtraffic, mtraffic_src = _step_traffic('mesh')
    mesh_traffic = None
    if mtraffic and world == 1:
      for name, a in mtraffic['kernels'].items():
        if 'mesh2d_kernel<1' in name:
          mesh_traffic = (a['dram_read_bytes'] + a['dram_write_bytes']) / a['launches']

    # N > 1: the sharded solve must equal the single-GPU solve bit for bit -- checked here,
    # inside the run whose numbers are reported, on a fresh 200-step relaxation.
    parity = None
    if world > 1:
      pcfg = mesh_config(mesh, 200)
      chunk.close()
      got, ek_s, t_s = mesh_sharded.relax_mesh_sharded(x0, prev, pcfg)
      slabs = [torch.empty((2, 1, b - a, MESH_N), device=dev) for a, b in parts]
      dist.all_gather(slabs, got.contiguous())
      if rank == 0:
        full = torch.cat(slabs, dim=2)
        want, ek_w, t_w = mesh.relax_mesh(torch.zeros_like(prev_full), prev_full, pcfg)
        same_nan = bool(torch.equal(torch.isnan(full), torch.isnan(want)))
        err = float(torch.nan_to_num(full - want).abs().max())
        parity = {'check': 'relax_mesh_sharded over %d ranks vs mesh.relax_mesh on one GPU, '
                           '200 FIRE steps on the same %d^2 mesh' % (world, MESH_N),
                  'max_abs_err': err, 'same_nan_pattern': same_nan, 'steps': [t_s, t_w],
                  'e_kin_rel_diff': abs(ek_s[-1] - ek_w[-1]) / max(abs(ek_w[-1]), 1e-30),
                  'bit_identical': same_nan and err == 0.0 and t_s == t_w}
        assert parity['bit_identical'], parity
        del prev_full, full, want

    hx = torch.zeros(x0.shape, dtype=torch.float32, pin_memory=True).numpy()
    hp = torch.empty(prev.shape, dtype=torch.float32, pin_memory=True)
    hp.copy_(prev)
    hp = hp.numpy()
    torch.cuda.synchronize()

    def mesh_e2e(i):
      if world == 1:
        out, e_kin, t = mesh.relax_mesh(hx, hp, cfg)
      else:
        out, e_kin, t = mesh_sharded.relax_mesh_sharded(hx, hp, cfg)
      assert t == iters and out.shape == hx.shape

    mesh_e2e(0)
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
      mesh_e2e(i)
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    result['mesh'] = {
        'metric': 'node-updates/s', 'value': mesh_value, 'unit': 'node-updates/s',
        'ms_per_step': ms / K, 'gpu_launches': launches,
        'scaling': 'strong (one 2048^2 mesh, rows sharded over the ranks; device-side '
                   'halo reads and step flags over NVLink peer memory, no collective '
                   'inside a chunk)' if world > 1 else 'n/a',
        'config': {'workload': f'mesh.relax_mesh chunk: {iters} FIRE steps, '
                               f'[2,1,{MESH_N},{MESH_N}] fp32, k0=0.1, k=0.1, stride 40, '
                               'prefer_orig_order, prev with 1% NaN; state 134 MB > L2'},
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peaks['hbm_gbs'],
                     'unit': 'GB/s', 'frac': achieved / peaks['hbm_gbs'],
                     'traffic': mesh_traffic, 'traffic_source': mtraffic_src,
                     'peak_source': peaks['source'],
                     'kernel': 'mesh2d_kernel<1,true,%s>' % ('true' if world > 1 else 'false'),
                     'kernel_us_per_launch': step_ms * 1e3,
                     'algorithmic_bytes_per_launch': local_nodes * MESH_BYTES_PER_UPDATE,
                     'kernel_us_per_launch_serialised': step_ms_events * 1e3,
                     'note': 'per-GPU: bytes of the rank-local slab / its launch time; the launch '
                             'time is the smaller of (a) CUDA events around each launch in a '
                             'separate chunk, which serialise the step and its one-block FIRE '
                             'reduce kernel (kernel_us_per_launch_serialised), and (b) the '
                             'device time of the timed chunk / its steps, an upper bound of the '
                             'step kernel inside the pipelined solve'},
        'e2e': {'value': nodes * iters * K / (e2e_ms * 1e-3),
                'unit': 'node-updates/s',
                'h2d_bytes_per_step': 2 * 2 * local_nodes * 4,
                'd2h_bytes_per_step': 2 * local_nodes * 4},
    }
    if parity is not None:
      result['mesh']['parity'] = parity

  # ---------------- fine flow of BASELINE config 2 (N = 1 only) ----------------
  # 4 x 4 grid of 4096^2 tiles with ~10 % overlap: stitch_elastic.compute_flow_map on the
  # 12 horizontal + 12 vertical overlap strips (one flow_field call per tile pair, host
  # tiles in, host flow fields out -- the reference's own call pattern).
  if args.path in ('both', 'flow') and world == 1 and rank == 0 and 'metric' in result:
    from sofima_b200 import stitch_elastic
    grid_tiles, cxm, cym = synth_tile_grid(4, 4, FLOW_TILE)

    def strips(i):
      n = 0
      for axis, cm in ((0, cxm), (1, cym)):
        fl, _ = stitch_elastic.compute_flow_map(grid_tiles, cm, axis, (PATCH, PATCH),
                                                (STEP, STEP), BATCH)
        n += sum(int(np.isfinite(f[0]).sum()) for f in fl.values())
      return n

    strips(0)
    t0 = time.perf_counter()
    n_pairs = sum(strips(i) for i in range(2))
    sec = time.perf_counter() - t0
    result['config2_flow'] = {
        'workload': 'stitch_elastic.compute_flow_map on the 24 overlap strips of a 4x4 grid '
                    f'of {FLOW_TILE}^2 uint8 tiles, patch {PATCH}, step {STEP}; host tiles '
                    'in, host flow fields out (e2e)',
        'value': n_pairs / sec, 'unit': 'patch-pairs/s', 'patch_pairs_per_grid': n_pairs // 2,
        'ms_per_grid': sec / 2 * 1e3}
    del grid_tiles

  # ---------------- stitching mesh of BASELINE config 2 (N = 1 only) ----------------
  # 4 x 4 tiles of 4096^2 px at stride 40 -> [2, 16, 102, 102] tile meshes relaxed
  # with the stitching prev_fn re-evaluated on the device in every step
  # (stitch_elastic.compute_target_mesh, notebooks/em_stitching.ipynb:545-603).
  if args.path in ('both', 'mesh') and world == 1 and rank == 0 and 'mesh' in result:
    from sofima_b200 import stitch_elastic
    fx, fy, sx0, nbors, sstride = synth_stitch(4, 4, (102, 102))
    s_iters = args.mesh_iters
    scfg = mesh.IntegrationConfig(
        dt=0.001, gamma=0., k0=0.01, k=0.1, stride=sstride, num_iters=s_iters,
        max_iters=s_iters, stop_v_max=0.0, dt_max=100, prefer_orig_order=True,
        start_cap=0.1, final_cap=10., remove_drift=True)
    prev_fn = stitch_elastic.target_mesh_fn(nbors, fx, fy, sstride)
    xd = torch.from_numpy(sx0).to(dev)

    def stitch_step(i):
      mesh.relax_mesh(xd, None, scfg, prev_fn=prev_fn)

    for i in range(W):
      stitch_step(i)
    s_ms, s_launches = timed(stitch_step, K)
    s_nodes = sx0.shape[1] * sx0.shape[2] * sx0.shape[3]
    t0 = time.perf_counter()
    for i in range(K):
      out, _, _ = mesh.relax_mesh(sx0, None, scfg, prev_fn=prev_fn)
    s_e2e = (time.perf_counter() - t0) * 1e3
    result['mesh']['stitching'] = {
        'workload': f'config 2 mesh: [2,16,102,102] tile meshes, {s_iters} FIRE steps, '
                    'stitching prev_fn evaluated on the device in every step',
        'value': s_nodes * s_iters * K / (s_ms * 1e-3), 'unit': 'node-updates/s',
        'us_per_integration_step': s_ms / K / s_iters * 1e3, 'gpu_launches': s_launches,
        'e2e': {'value': s_nodes * s_iters * K / (s_e2e * 1e-3), 'unit': 'node-updates/s',
                'h2d_bytes_per_step': int(sx0.nbytes + fx.nbytes + fy.nbytes),
                'd2h_bytes_per_step': int(sx0.nbytes)},
        'note': 'launch-latency bound (166 k nodes): two launches per integration step'}

  # ---------------- widened rows, N = 1 only ----------------
  if world == 1 and rank == 0 and args.path == 'both':
    from sofima_b200 import map_utils
    from oracle import stitch_oracle, flow_oracle
    widened = {}
    # f-1 compose_maps_fast: 8 sections of 1024^2 nodes, stride 40 (HBM: read map1 and
    # map2 once, write the result = 24 B per node).
    rngw = np.random.default_rng(5)
    m1 = torch.from_numpy((rngw.standard_normal((2, 8, 1024, 1024)) * 20).astype(np.float32)).to(dev)
    m2 = torch.from_numpy((rngw.standard_normal((2, 8, 1024, 1024)) * 5).astype(np.float32)).to(dev)
    cm = lambda i: map_utils.compose_maps_fast(m1, (0, 0, 0), 40.0, m2, (0, 3, 2), 40.0)
    for i in range(W):
      cm(i)
    c_ms, _ = timed(cm, 20)
    c_nodes = 8 * 1024 * 1024
    c_gbs = c_nodes * 24 / (c_ms / 20 * 1e-3) / 1e9
    sm1, sm2 = m1[:, :1, :512, :512].cpu().numpy(), m2[:, :1, :512, :512].cpu().numpy()
    t0 = time.perf_counter()
    stitch_oracle.compose_maps_fast(sm1, (0, 0, 0), 40.0, sm2, (0, 3, 2), 40.0)
    c_cpu = 512 * 512 / (time.perf_counter() - t0)
    widened['compose_maps_fast'] = {
        'workload': 'map_utils.compose_maps_fast, [2,8,1024,1024] maps, stride 40, HBM-resident',
        'value': c_nodes / (c_ms / 20 * 1e-3), 'unit': 'nodes/s', 'us_per_call': c_ms / 20 * 1e3,
        'roofline': {'bound': 'hbm', 'achieved': c_gbs, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                     'frac': c_gbs / peaks['hbm_gbs'], 'traffic': None,
                     'peak_source': peaks['source'],
                     'algorithmic_bytes_per_node': 24},
        'cpu_baseline': {'value': c_cpu, 'unit': 'nodes/s', 'cores': 1, 'kind': 'port',
                         'sample': 'one 512^2 section, oracle/stitch_oracle.py (NumPy)'}}
    del m1, m2
    # f-2 whole-strip masked correlation (stitch_rigid._estimate_offset): one 4096 x 300
    # strip pair with masks, host arrays in, offset out (e2e).
    strip = synth_tile_pairs(1, FLOW_TILE, 41, torch.device('cpu'))[0]
    sa = np.ascontiguousarray(strip[0].numpy()[:, -300:])
    sb = np.ascontiguousarray(strip[1].numpy()[:, -300:])  # same region, shifted by (3, -4)
    ma = rngw.random(sa.shape) < 0.05
    mb = rngw.random(sb.shape) < 0.05
    kw = dict(pre_mask=ma, post_mask=mb, patch_size=sa.shape, step=(1, 1), batch_size=1)
    calc.flow_field(sa, sb, **kw)
    t0 = time.perf_counter()
    for i in range(5):
      calc.flow_field(sa, sb, **kw)
    s_ms = (time.perf_counter() - t0) / 5 * 1e3
    t0 = time.perf_counter()
    flow_oracle.MaskedXCorrWithStatsCalculator().flow_field(sa, sb, **kw)
    s_cpu_ms = (time.perf_counter() - t0) * 1e3
    widened['whole_strip_masked_xcorr'] = {
        'workload': 'one masked correlation of a 4096 x 300 overlap strip pair '
                    '(8192 x 600-point transforms, Padfield normalisation, peak), e2e',
        'value': 1e3 / s_ms, 'unit': 'strip-pairs/s', 'ms_per_strip_pair': s_ms,
        'cpu_baseline': {'value': 1e3 / s_cpu_ms, 'unit': 'strip-pairs/s',
                         'cores': len(os.sched_getaffinity(0)), 'kind': 'port',
                         'sample': 'the same strip pair, oracle/flow_oracle.py (pocketfft)'}}
    # f-3 warp.ndimage_warp: one 8192^2 uint8 image through a 206^2-node map (stride 40),
    # order 1; device image in, device image out.  HBM: read + write the image = 2 B / px.
    from sofima_b200 import warp as warp_mod
    from oracle import warp_oracle
    import scipy.ndimage as ndi
    wimg = torch.randint(0, 255, (8192, 8192), dtype=torch.uint8, device=dev)
    wmap = np.stack([ndi.gaussian_filter(rngw.standard_normal((206, 206)), 6) * 400,
                     ndi.gaussian_filter(rngw.standard_normal((206, 206)), 6) * 400])
    wf = lambda i: warp_mod.ndimage_warp(wimg, wmap, (40, 40), (1024, 1024), (0, 0))
    for i in range(W):
      wf(i)
    ctx.set_timing(True)
    w_ms, _ = timed(wf, 10)
    w_rep = ctx.timing_report()
    ctx.set_timing(False)
    w_kernel_ms = w_rep['warp_image']['ms'] / w_rep['warp_image']['n']
    w_gbs = 2 * 8192 * 8192 / (w_kernel_ms * 1e-3) / 1e9
    crop = wimg[:1024, :1024].cpu().numpy()
    t0 = time.perf_counter()
    warp_oracle.ndimage_warp(crop, wmap[:, :27, :27], (40, 40))
    w_cpu = 1024 * 1024 / (time.perf_counter() - t0)
    widened['ndimage_warp'] = {
        'workload': 'warp.ndimage_warp, 8192^2 uint8 image, 206^2-node map (stride 40), '
                    'order 1, device image in / out',
        'value': 8192 * 8192 / (w_ms / 10 * 1e-3), 'unit': 'pixels/s',
        'ms_per_image': w_ms / 10, 'kernel_ms': w_kernel_ms,
        'roofline': {'bound': 'hbm', 'achieved': w_gbs, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                     'frac': w_gbs / peaks['hbm_gbs'], 'traffic': None,
                     'peak_source': peaks['source'], 'algorithmic_bytes_per_pixel': 2,
                     'note': 'float64 arithmetic per pixel (bit-exact vs SciPy) bounds the '
                             'kernel, not HBM'},
        'cpu_baseline': {'value': w_cpu, 'unit': 'pixels/s', 'cores': 1, 'kind': 'port',
                         'sample': 'one 1024^2 crop, oracle/warp_oracle.py (NumPy float64, '
                                   '== scipy.ndimage.map_coordinates)'}}
    del wimg
    # f-3 warp.warp_subvolume: 4 sections of 4096^2 uint8 through a smooth map (stride 32),
    # Lanczos-4 (the reference's default); device sections in / out.
    from sofima_b200 import compat as compat_mod
    from oracle import warp_cv_oracle
    sv_img = rngw.integers(0, 256, (1, 4, 4096, 4096), dtype=np.uint8)
    sv_map = np.stack([ndi.gaussian_filter(rngw.standard_normal((4, 129, 129)), (0, 3, 3)) * 60
                       for _ in range(2)])
    sv_box = compat_mod.BoundingBox(start=(0, 0, 0), size=(4096, 4096, 4))
    sv_mbox = compat_mod.BoundingBox(start=(0, 0, 0), size=(129, 129, 4))
    sv_dev = torch.from_numpy(sv_img).to(dev)
    sf = lambda i: warp_mod.warp_subvolume(sv_dev, sv_box, sv_map, sv_mbox, 32, sv_box)
    for i in range(W):
      sv_out = sf(i)
    ctx.set_timing(True)
    sv_ms, _ = timed(sf, 10)
    sv_rep = ctx.timing_report()
    ctx.set_timing(False)
    sv_kernel_ms = sv_rep['warp_subvolume']['ms'] / sv_rep['warp_subvolume']['n']
    sv_gbs = 2 * sv_img.size / (sv_kernel_ms * 1e-3) / 1e9
    entry = {
        'workload': 'warp.warp_subvolume, 4 sections of 4096^2 uint8, 129^2-node map (stride '
                    '32), Lanczos-4 (scipy grid interpolation + OpenCV convertMaps / remap '
                    'semantics), device sections in / out',
        'value': sv_img.size / (sv_ms / 10 * 1e-3), 'unit': 'pixels/s',
        'ms_per_call': sv_ms / 10, 'kernel_ms': sv_kernel_ms,
        'roofline': {'bound': 'hbm', 'achieved': sv_gbs, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                     'frac': sv_gbs / peaks['hbm_gbs'], 'traffic': None,
                     'peak_source': peaks['source'], 'algorithmic_bytes_per_pixel': 2,
                     'note': '64 taps per pixel served by L1 / L2: bound by load issue and '
                             'latency, not HBM (profiles/ncu_r2_warp_lanczos.txt)'}}
    try:
      t0 = time.perf_counter()
      sv_ref = warp_cv_oracle.reference_pipeline(sv_img[:, :1], sv_map[:, :1], 32, 'lanczos')
      sv_cpu = sv_ref.size / (time.perf_counter() - t0)
      entry['cpu_baseline'] = {
          'value': sv_cpu, 'unit': 'pixels/s', 'cores': 1, 'kind': 'reference',
          'sample': 'one of the sections: the scipy + cv2 calls of warp.py:144-165 with the '
                    'real libraries',
          'identical_to_gpu': bool(np.array_equal(sv_ref, sv_out[:, :1].cpu().numpy()))}
    except ImportError as e:
      entry['cpu_baseline'] = {'unavailable': str(e)}
    widened['warp_subvolume'] = entry
    del sv_dev, sv_out
    result['widened'] = widened

  # ---------------- CPU baseline (rank 0, N = 1 only) ----------------
  if rank == 0 and world == 1 and not args.no_cpu_baseline:
    cores = len(os.sched_getaffinity(0))
    if 'metric' in result:
      arm = CpuFlowArm()
      arm.step(0)
      n_cpu, s_cpu = 0, 0.0
      for i in range(6):
        n, sec = arm.step(1 + i)
        n_cpu += n
        s_cpu += sec
      result['cpu_baseline'] = {
          'value': n_cpu / s_cpu, 'unit': 'patch-pairs/s', 'cores': cores, 'kind': 'port',
          'sample': f'6 reference batches of {BATCH} patch pairs of the same {FLOW_TILE}^2 tile '
                    f'pair ({s_cpu:.1f} s), oracle/flow_oracle.py on pocketfft fp32'}
    if 'mesh' in result:
      v, s, threads = time_cpu_mesh(MESH_N, 40)
      result['mesh']['cpu_baseline'] = {
          'value': v, 'unit': 'node-updates/s', 'cores': threads, 'kind': 'port',
          'sample': f'40 FIRE steps on {MESH_N}^2 nodes ({s:.1f} s), '
                    'oracle/mesh_oracle.c (OpenMP)'}

  if rank == 0:
    if 'metric' not in result:  # --path mesh: promote the mesh numbers
      m = result.pop('mesh')
      result.update(m)
    line = {
        'n_gpus': world, 'steps': K, 'warmup': W, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'flow_field on one {FLOW_TILE}x{FLOW_TILE} uint8 tile '
                               f'pair per step, patch {PATCH}, step {STEP}, batch '
                               f'{BATCH} (9801 patch pairs); inputs rotate over 6 tile '
                               'pairs = 201 MB > L2',
                   'per_rank': 'each rank processes its own tile pairs (no collective)'},
        'clocks': clocks.get('flow', clocks.get('mesh')),
        'clocks_mesh': clocks.get('mesh'),
    }
    if args.path == 'mesh':
      line['config'] = result.pop('config')
    line.update(result)
    emit(line)
  if world > 1:
    from sofima_b200 import mesh_sharded as _ms
    _ms.clear_shard_cache()  # peer mappings go before the process group does
    dist.destroy_process_group()


_REAL_STDOUT = None


def _protect_stdout():
  """Libraries (NCCL banner, torchrun) must not write to stdout: the contract is ONE
  JSON line there.  fd 1 is pointed at stderr; emit() writes to the saved fd."""
  global _REAL_STDOUT
  if _REAL_STDOUT is None:
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)


def emit(line: dict):
  out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
  out.write(json.dumps(line) + '\n')
  out.flush()


def main():
  _protect_stdout()
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=5)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', choices=['ours', 'reference'], default='ours')
  ap.add_argument('--path', choices=['both', 'flow', 'mesh'], default='both')
  ap.add_argument('--mesh-iters', type=int, default=1000)
  ap.add_argument('--no-cpu-baseline', action='store_true')
  args = ap.parse_args()
  if args.warmup < 3 and args.impl == 'ours':
    args.warmup = max(args.warmup, 1)
  if args.impl == 'reference':
    # torch.distributed.run exports OMP_NUM_THREADS=1 to its workers; the CPU arm is meant to
    # use every host core, and OpenMP / OpenBLAS read the variable when they are loaded --
    # so it is set and the interpreter re-executed before anything else happens.
    cores = str(len(os.sched_getaffinity(0)))
    if int(os.environ.get('RANK', '0')) == 0 and os.environ.get('OMP_NUM_THREADS') != cores:
      env = dict(os.environ, OMP_NUM_THREADS=cores, SOFIMA_BENCH_REEXEC='1')
      if not os.environ.get('SOFIMA_BENCH_REEXEC'):
        os.dup2(_REAL_STDOUT.fileno(), 1)
        os.execve(sys.executable, [sys.executable] + sys.argv, env)
    run_reference(args)
  else:
    run_ours(args)


if __name__ == '__main__':
  main()
