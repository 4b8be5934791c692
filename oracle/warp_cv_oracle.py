"""NumPy restatement of `warp.warp_subvolume` (TEST ORACLE).

Follows /root/reference/warp.py:58-186.  The per-pixel work of that function lives in two
third-party libraries that are not vendored in the reference:

  * scipy.interpolate.RegularGridInterpolator (method 'linear', bounds_error=False,
    fill_value=None) -- scipy is unpinned in the reference's setup.cfg; restated below
    from the published algorithm (scipy/interpolate/_rgi.py: `_find_indices`,
    `_evaluate_linear`, and the 2-d float64 fast path `evaluate_linear_2d`);
  * cv2.convertMaps(..., CV_16SC2) + cv2.remap(..., BORDER_CONSTANT 0) -- opencv-python
    unpinned; restated from the published algorithm (modules/imgproc/src/imgwarp.cpp:
    `initInterTab1D/2D`, `convertMaps`, `remapNearest`, `remapBilinear`, `remapBicubic`,
    `remapLanczos4`).

Both libraries ARE installed in this image (scipy 1.18, cv2 4.13), so every function here
is pinned bit-for-bit against the real routine in tests/test_oracle_warp_cv.py, and the
whole of `warp_subvolume` against outputs of the reference itself
(tests/golden/warp_cv_golden.npz, made by tests/golden/make_warp_cv_golden.py).

Restated rules:
  * map densification: per axis, interval i = clip(#{grid <= q} - 1, 0, n - 2),
    t = (q - g[i]) / (g[i + 1] - g[i]) in float64 (extrapolating outside the grid);
    float64 maps: sum of ((v * wy) * wx) over the corners (00, 01, 10, 11);
    other maps:   sum of (v * (wy * wx)); the result is cast to float32;
  * convertMaps: nearest: (rint(x), rint(y)) saturated to int16;
    otherwise ix = rint(x * 32) (int32, INT_MIN for NaN / overflow), integer part ix >> 5
    saturated to int16, fraction ix & 31;
  * tables: 1-d coefficient rows for 32 fractions (linear: 1 - f, f; cubic: A = -0.75;
    Lanczos-4: the sin/cos recurrence of imgwarp.cpp, normalised in float32);
    uint8 images use the 2-d integer table rint(wy * wx * 32768) whose sum is forced to
    32768 by correcting the largest / smallest of the central 2x2 entries, and
    (sum + 2**14) >> 15 saturated to uint8;
    other types use float32 products wy * wx, summed row by row when the whole footprint
    is inside the image (linear: the four terms left to right) and tap by tap otherwise;
    samples outside the image are 0;
  * integer outputs: rint, saturated.

Test infrastructure only: never imported from `sofima_b200/`.
"""

from __future__ import annotations

import numpy as np

INTER_BITS = 5
TAB = 1 << INTER_BITS
SCALE = 1 << 15
_INT_MIN = -(1 << 31)

KSIZE = {'linear': 2, 'cubic': 4, 'lanczos': 8}


# ---------------------------------------------------------------------------------------
# scipy.interpolate.RegularGridInterpolator, linear, extrapolating
# ---------------------------------------------------------------------------------------
def _intervals(grid, q):
  grid = np.asarray(grid, np.float64)
  q = np.asarray(q, np.float64)
  i = np.clip(np.searchsorted(grid, q, side='right') - 1, 0, len(grid) - 2)
  t = (q - grid[i]) / (grid[i + 1] - grid[i])
  return i, t


def rgi_linear_2d(grid_y, grid_x, values, qy, qx):
  """RegularGridInterpolator((grid_y, grid_x), values, bounds_error=False,
  fill_value=None)((qy, qx)) for broadcastable integer-or-float query arrays."""
  values = np.asarray(values)
  iy, ty = _intervals(grid_y, qy)
  ix, tx = _intervals(grid_x, qx)
  iy, ix, ty, tx = np.broadcast_arrays(iy, ix, ty, tx)
  fast = values.dtype == np.float64
  out = np.zeros(iy.shape, np.float64)
  for dy, wy in ((0, 1 - ty), (1, ty)):
    for dx, wx in ((0, 1 - tx), (1, tx)):
      v = values[iy + dy, ix + dx].astype(np.float64)
      out = out + ((v * wy) * wx if fast else v * (wy * wx))
  return out


# ---------------------------------------------------------------------------------------
# cv2.convertMaps to CV_16SC2
# ---------------------------------------------------------------------------------------
def _cv_round(v):
  """cvRound of float32 values: round-half-even, INT_MIN when not representable."""
  v = np.asarray(v, np.float32).astype(np.float64)
  ok = np.isfinite(v) & (np.abs(v) < 2.0**31)
  r = np.full(v.shape, _INT_MIN, np.int64)
  r[ok] = np.rint(v[ok]).astype(np.int64)
  return r


def convert_maps(dx, dy, nn):
  """(xy [h, w, 2] int16, frac [h, w] uint16 or None)."""
  dx = np.asarray(dx, np.float32)
  dy = np.asarray(dy, np.float32)
  if nn:
    ix, iy = _cv_round(dx), _cv_round(dy)
    xy = np.stack([np.clip(ix, -32768, 32767), np.clip(iy, -32768, 32767)], -1)
    return xy.astype(np.int16), None
  ix = _cv_round(dx * np.float32(TAB))
  iy = _cv_round(dy * np.float32(TAB))
  xy = np.stack([np.clip(ix >> INTER_BITS, -32768, 32767),
                 np.clip(iy >> INTER_BITS, -32768, 32767)], -1).astype(np.int16)
  frac = ((iy & (TAB - 1)) * TAB + (ix & (TAB - 1))).astype(np.uint16)
  return xy, frac


# ---------------------------------------------------------------------------------------
# interpolation tables
# ---------------------------------------------------------------------------------------
def _coeffs(method, x):
  x = np.float32(x)
  one = np.float32(1)
  if method == 'linear':
    return np.array([one - x, x], np.float32)
  if method == 'cubic':
    a = np.float32(-0.75)
    c = np.zeros(4, np.float32)
    c[0] = ((a * (x + one) - np.float32(5) * a) * (x + one) + np.float32(8) * a) * (x + one) \
        - np.float32(4) * a
    c[1] = ((a + np.float32(2)) * x - (a + np.float32(3))) * x * x + one
    c[2] = ((a + np.float32(2)) * (one - x) - (a + np.float32(3))) * (one - x) * (one - x) + one
    c[3] = one - c[0] - c[1] - c[2]
    return c
  assert method == 'lanczos'
  s45 = 0.70710678118654752440084436210485
  cs = [[1, 0], [-s45, -s45], [0, 1], [s45, -s45], [-1, 0], [s45, s45], [0, -1], [-s45, s45]]
  c = np.zeros(8, np.float32)
  if x < np.finfo(np.float32).eps:
    c[3] = 1
    return c
  total = np.float32(0)
  y0 = -(np.float64(x) + 3) * np.pi * 0.25
  s0, c0 = np.sin(y0), np.cos(y0)
  for i in range(8):
    y = -(np.float64(x) + 3 - i) * np.pi * 0.25
    c[i] = np.float32((cs[i][0] * s0 + cs[i][1] * c0) / (y * y))
    total = np.float32(total + c[i])
  return (c * (one / total)).astype(np.float32)


_TABS: dict = {}


def tables(method):
  """(tab1 [32, k] float32, itab [32, 32, k, k] int32) of an interpolation method."""
  if method in _TABS:
    return _TABS[method]
  k = KSIZE[method]
  tab1 = np.stack([_coeffs(method, np.float32(i) / np.float32(TAB)) for i in range(TAB)])
  itab = np.zeros((TAB, TAB, k, k), np.int32)
  c0 = k // 2
  for fy in range(TAB):
    for fx in range(TAB):
      v = np.outer(tab1[fy], tab1[fx]).astype(np.float32)
      t = np.rint(v * np.float32(SCALE)).astype(np.int32)
      if k == 2:  # products of k/32: exact, the sum is 32768 (a weight of 32768 behaves
        itab[fy, fx] = t  # like the saturated 32767 + 1 elsewhere after the >> 15)
        continue
      t = np.clip(t, -32768, 32767)
      diff = int(t.sum()) - SCALE
      if diff:
        lo = hi = (c0, c0)
        for a in range(c0, c0 + 2):
          for b in range(c0, c0 + 2):
            if t[a, b] < t[lo]:
              lo = (a, b)
            elif t[a, b] > t[hi]:
              hi = (a, b)
        if diff < 0:
          t[hi] -= diff
        else:
          t[lo] -= diff
      itab[fy, fx] = t
  _TABS[method] = (tab1, itab)
  return _TABS[method]


# ---------------------------------------------------------------------------------------
# cv2.remap on CV_16SC2 maps, BORDER_CONSTANT with value 0
# ---------------------------------------------------------------------------------------
def _saturate(v, dtype):
  dtype = np.dtype(dtype)
  if dtype.kind == 'f':
    return v.astype(dtype)
  info = np.iinfo(dtype)
  return np.clip(np.rint(v), info.min, info.max).astype(dtype)


def _fetch(img, yy, xx, dtype):
  ok = (xx >= 0) & (xx < img.shape[1]) & (yy >= 0) & (yy < img.shape[0])
  v = np.zeros(xx.shape, dtype)
  v[ok] = img[yy[ok], xx[ok]]
  return v


def remap(img, xy, frac, method):
  """cv2.remap(img, xy, frac, interpolation=method) for a 2-d image."""
  img = np.asarray(img)
  x = xy[..., 0].astype(np.int64)
  y = xy[..., 1].astype(np.int64)
  if method == 'nearest':
    return _fetch(img, y, x, img.dtype)
  k = KSIZE[method]
  tab1, itab = tables(method)
  fx = (frac & (TAB - 1)).astype(np.int64)
  fy = (frac >> INTER_BITS).astype(np.int64)
  x0, y0 = x - (k // 2 - 1), y - (k // 2 - 1)
  if img.dtype == np.uint8:
    acc = np.zeros(x.shape, np.int64)
    for i in range(k):
      for j in range(k):
        acc += itab[fy, fx, i, j] * _fetch(img, y0 + i, x0 + j, np.int64)
    return np.clip((acc + (1 << 14)) >> 15, 0, 255).astype(np.uint8)
  f32 = np.float32
  inner = (x0 >= 0) & (x0 + k - 1 < img.shape[1]) & (y0 >= 0) & (y0 + k - 1 < img.shape[0])
  rowwise = np.zeros(x.shape, f32)
  seq = np.zeros(x.shape, f32)
  for i in range(k):
    row = None
    for j in range(k):
      w = (tab1[fy, i] * tab1[fx, j]).astype(f32)
      term = (_fetch(img, y0 + i, x0 + j, f32) * w).astype(f32)
      row = term if row is None else (row + term).astype(f32)
      seq = (seq + term).astype(f32)
    rowwise = (rowwise + row).astype(f32)
  total = seq if k == 2 else np.where(inner, rowwise, seq)
  return _saturate(total, img.dtype)


# ---------------------------------------------------------------------------------------
# warp.warp_subvolume
# ---------------------------------------------------------------------------------------
def warp_subvolume(image, image_box, coord_map, map_box, stride, out_box, interpolation=None,
                   offset=0.0):
  """warp.py:58-186.  Boxes are anything with `.start` / `.size` (xyz)."""
  image = np.asarray(image)
  back = None
  if image.dtype == np.uint64:  # warp.py:93-99
    method = 'nearest'
    ids = np.unique(np.append(image.ravel(), np.uint64(0)))
    image = np.searchsorted(ids, image).astype(np.int32)
    back = ids
    orig_dtype = np.uint64
  else:
    method = interpolation or 'lanczos'
    orig_dtype = image.dtype
    if image.dtype == np.uint32:  # warp.py:109-115
      if image.max() >= 2**16:
        raise ValueError('Image warping supported up to uint16 only. For segmentation '
                         'data, use uint64.')
      image = image.astype(np.uint16)
  coord_map = np.asarray(coord_map)
  skipped = np.all(np.isnan(coord_map), axis=(0, 2, 3))  # warp.py:117-119
  mstart = np.asarray(map_box.start)
  istart = np.asarray(image_box.start)
  ostart = np.asarray(out_box.start)
  osize = [int(v) for v in out_box.size]
  # map_utils.to_absolute without a box (map_utils.py:150-185), then warp.py:124-126:
  # NumPy in-place adds (computed in float64, stored in the map's own dtype).
  abs_map = coord_map.copy()
  gy, gx = np.mgrid[:coord_map.shape[2], :coord_map.shape[3]]
  abs_map[0] += gx * float(stride)
  abs_map[1] += gy * float(stride)
  abs_map += (mstart[:2] * stride - istart[:2] + offset).reshape(2, 1, 1, 1)
  map_y = (np.arange(coord_map.shape[2]) + mstart[1]) * stride - ostart[1] + offset
  map_x = (np.arange(coord_map.shape[3]) + mstart[0]) * stride - ostart[0] + offset
  warped = np.zeros([image.shape[0], osize[2], osize[1], osize[0]], image.dtype)
  qy = np.arange(osize[1])[:, None]
  qx = np.arange(osize[0])[None, :]
  for z in range(image.shape[1]):
    if skipped[z]:
      continue
    dx = rgi_linear_2d(map_y, map_x, abs_map[0, z], qy, qx).astype(np.float32)
    dy = rgi_linear_2d(map_y, map_x, abs_map[1, z], qy, qx).astype(np.float32)
    xy, frac = convert_maps(dx, dy, method == 'nearest')
    for c in range(image.shape[0]):
      warped[c, z] = remap(image[c, z], xy, frac, method)
  if back is not None:
    return back[warped]
  return warped.astype(orig_dtype)


def reference_pipeline(image, coord_map, stride, interpolation='lanczos', threads=1):
  """The reference's per-section calls (warp.py:123-165) with the REAL scipy and cv2, for
  boxes that all start at the origin and cover the image: the CPU baseline of the warp path
  (bench.py, tools/bench_warp.py) and a full-size cross-check.  Raises ImportError without
  cv2."""
  import cv2 as cv  # pylint: disable=g-import-not-at-top
  from concurrent import futures  # pylint: disable=g-import-not-at-top
  from scipy import interpolate  # pylint: disable=g-import-not-at-top
  flag = {'nearest': cv.INTER_NEAREST, 'linear': cv.INTER_LINEAR, 'cubic': cv.INTER_CUBIC,
          'lanczos': cv.INTER_LANCZOS4}[interpolation]
  n, nz, h, w = image.shape
  gy, gx = np.mgrid[:coord_map.shape[2], :coord_map.shape[3]]
  abs_map = coord_map.copy()
  abs_map[0] += gx * stride
  abs_map[1] += gy * stride
  pts = (np.arange(coord_map.shape[2]) * float(stride),
         np.arange(coord_map.shape[3]) * float(stride))
  out = np.zeros_like(image)
  oy, ox = np.mgrid[:h, :w]

  def section(z):
    dx = interpolate.RegularGridInterpolator(pts, abs_map[0, z], bounds_error=False,
                                             fill_value=None)((oy, ox)).astype(np.float32)
    dy = interpolate.RegularGridInterpolator(pts, abs_map[1, z], bounds_error=False,
                                             fill_value=None)((oy, ox)).astype(np.float32)
    m1, m2 = cv.convertMaps(dx, dy, dstmap1type=cv.CV_16SC2,
                            nninterpolation=(flag == cv.INTER_NEAREST))
    for c in range(n):
      out[c, z] = cv.remap(image[c, z], m1, m2, interpolation=flag)

  with futures.ThreadPoolExecutor(max_workers=threads) as ex:
    list(ex.map(section, range(nz)))
  return out


class _Box:
  """Minimal stand-in for a bounding box (xyz start / size) in the known-answer tests."""

  def __init__(self, start, size):
    self.start, self.size = np.asarray(start), np.asarray(size)
    self.end = self.start + self.size


def check_reference_kats(warp_subvolume_fn, box=_Box):
  """The reference's own warp_subvolume tests (tests/warp_test.py:27-78) against any
  implementation: label translation with ids beyond int32, and a 45-degree rotation of a
  rhombus with the default (Lanczos) interpolation."""
  image = np.zeros((1, 2, 100, 100), dtype=np.uint64)
  image[0, 0, 40, 30] = 42
  image[0, 1, 50, 40] = 2**40
  cmap = np.zeros((2, 2, 15, 15))
  cmap[0, 0] = 10
  cmap[1, 1] = 17
  warped = warp_subvolume_fn(image, box(start=(0, 0, 0), size=(100, 100, 2)), cmap,
                             box(start=(0, 0, 0), size=(15, 15, 2)), 10,
                             box(start=(10, 20, 0), size=(90, 80, 2)))
  expected = np.zeros((1, 2, 80, 90))
  expected[0, 0, 20, 10] = 42
  expected[0, 1, 13, 30] = 2**40
  np.testing.assert_array_equal(warped, expected)
  hy, hx = np.mgrid[-50:50, -50:50]
  img = np.zeros((1, 1, 100, 100), dtype=np.uint8)
  img[0, 0][np.abs(hy) + np.abs(hx) < 25] = 255
  ang = np.pi / 4
  rmap = np.zeros((2, 1, 10, 10))
  rmap[0, 0] = (np.cos(ang) * hx[::10, ::10] - np.sin(ang) * hy[::10, ::10]) - hx[::10, ::10]
  rmap[1, 0] = (np.sin(ang) * hx[::10, ::10] + np.cos(ang) * hy[::10, ::10]) - hy[::10, ::10]
  full = box(start=(0, 0, 0), size=(100, 100, 1))
  warped = warp_subvolume_fn(img, full, rmap, box(start=(0, 0, 0), size=(10, 10, 1)), 10, full)
  mask = np.zeros((1, 1, 100, 100), dtype=bool)
  mask[0, 0, 33:68, 33:68] = True
  assert np.all(warped[mask] > 128) and np.all(warped[~mask] < 64)
