"""Measurement of the warp path: warp.warp_subvolume on 4096 x 4096 uint8
sections through a smooth map (stride 32), Lanczos-4 (the reference's default).

  python tools/bench_warp.py [--sections 16] [--steps 10] [--cpu-sections 2] [--out f.json]

Prints one JSON line: device-resident throughput (CUDA events on the launching stream),
end-to-end throughput through the public call with NumPy in / NumPy out, the roofline of the
kernel (algorithmic bytes = one read + one write per pixel and channel) and the reference's
CPU pipeline (scipy RegularGridInterpolator + cv2.convertMaps + cv2.remap, the calls of
reference warp.py:144-165, restated with the real libraries) timed on the host cores.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.ndimage as ndi


def make_case(nz, size, stride, seed=0):
  rng = np.random.default_rng(seed)
  img = rng.integers(0, 256, (1, nz, size, size), dtype=np.uint8)
  m = size // stride + 1
  cmap = np.stack([ndi.gaussian_filter(rng.standard_normal((nz, m, m)), (0, 3, 3)) * 60
                   for _ in range(2)])
  return img, cmap


def cpu_pipeline(img, cmap, stride, inter, threads):
  """The reference's per-section work with the real libraries (warp.py:123-165)."""
  from oracle import warp_cv_oracle
  return warp_cv_oracle.reference_pipeline(img, cmap, stride, inter, threads)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--sections', type=int, default=16)
  ap.add_argument('--size', type=int, default=4096)
  ap.add_argument('--steps', type=int, default=10)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--cpu-sections', type=int, default=2)
  ap.add_argument('--interpolation', default='lanczos')
  ap.add_argument('--out', default=None)
  a = ap.parse_args()
  import torch
  from sofima_b200 import _native, compat, warp
  stride = 32
  img, cmap = make_case(a.sections, a.size, stride)
  box = compat.BoundingBox(start=(0, 0, 0), size=(a.size, a.size, a.sections))
  mbox = compat.BoundingBox(start=(0, 0, 0), size=(cmap.shape[3], cmap.shape[2], a.sections))
  ctx = _native.Context.get(0)
  img_d = torch.from_numpy(img).cuda()
  pixels = img.size

  def dev_step():
    return warp.warp_subvolume(img_d, box, cmap, mbox, stride, box, a.interpolation)

  for _ in range(a.warmup):
    got = dev_step()
  torch.cuda.synchronize()
  ctx.set_timing(True)
  t0 = time.perf_counter()
  for _ in range(a.steps):
    got = dev_step()
  torch.cuda.synchronize()
  dev_wall = (time.perf_counter() - t0) / a.steps
  rep = ctx.timing_report()
  ctx.set_timing(False)
  k = rep['warp_subvolume']
  kernel_s = k['ms'] / k['n'] * 1e-3
  # end to end: NumPy in, NumPy out (pageable host arrays, as a caller of the reference has)
  for _ in range(2):
    host = warp.warp_subvolume(img, box, cmap, mbox, stride, box, a.interpolation)
  t0 = time.perf_counter()
  for _ in range(a.steps):
    host = warp.warp_subvolume(img, box, cmap, mbox, stride, box, a.interpolation)
  e2e_s = (time.perf_counter() - t0) / a.steps
  assert np.array_equal(host, got.cpu().numpy())
  # CPU reference pipeline on a bounded sample, and parity on that sample
  threads = min(os.cpu_count() or 1, a.cpu_sections)
  cpu = None
  try:
    ns = a.cpu_sections
    t0 = time.perf_counter()
    ref = cpu_pipeline(img[:, :ns], cmap[:, :ns], stride, a.interpolation, threads)
    cpu_s = time.perf_counter() - t0
    cpu = dict(value=ref.size / cpu_s / 1e6, unit='Mpixel/s', cores=threads, kind='reference',
               sample=f'{ns} of the {a.sections} sections, scipy + cv2 calls of warp.py:144-165',
               identical_to_gpu=bool(np.array_equal(ref, host[:, :ns])))
  except ImportError as e:
    cpu = dict(unavailable=str(e))
  peaks = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))
  alg_bytes = 2 * pixels  # uint8: one source read + one destination write per pixel
  line = dict(
      metric='warped pixels/s', unit='Mpixel/s', value=pixels / kernel_s / 1e6,
      config=dict(workload=f'warp_subvolume, {a.sections} sections of {a.size}x{a.size} uint8, '
                  f'map stride {stride}, {a.interpolation}', image_mb=img.nbytes / 1e6),
      device_wall_ms=dev_wall * 1e3, kernel_ms=kernel_s * 1e3,
      e2e=dict(value=pixels / e2e_s / 1e6, unit='Mpixel/s', h2d_bytes_per_step=int(img.nbytes),
               d2h_bytes_per_step=int(img.nbytes)),
      roofline=dict(bound='hbm', achieved=alg_bytes / kernel_s / 1e9, peak=peaks['hbm_gbs'],
                    unit='GB/s', frac=alg_bytes / kernel_s / 1e9 / peaks['hbm_gbs'],
                    traffic=None, note='gather kernel: 64 taps per pixel come from L1/L2; the '
                    'binding resource is load issue, see profiles/'),
      cpu_baseline=cpu)
  print(json.dumps(line))
  if a.out:
    with open(a.out, 'w') as f:
      json.dump(line, f, indent=1)


if __name__ == '__main__':
  main()
