// Flow-field post-filters on the device: the small filters that sit between
// every pair of hot-path calls in the reference's section loops --
//   flow_utils.clean_flow        (reference flow_utils.py:37-78)
//   flow_utils.reconcile_flows   (reference flow_utils.py:81-135)
//   map_utils.mask_irregular     (reference map_utils.py:737-786)
// so that EstimateMissingFlow (processor/flow.py:809-815) and RelaxMesh.relax_mesh
// (processor/mesh.py:466-471) do not have to take the flow field / the mesh to the host
// between two kernels.  Every comparison is the reference's fp32 comparison; medians are
// exact order statistics; connected components are exact -- the results equal the NumPy /
// SciPy filters bit for bit (tests/test_flowfilt_gpu.py against tests/golden/
// flow_utils_golden.npz).
#include <cfloat>

#include "common.cuh"

namespace sofima {
namespace filt {

constexpr int kThreads = 256;

__device__ __forceinline__ float nan_to_num(float v) {  // np.nan_to_num defaults
  if (v != v) return 0.f;
  return fminf(fmaxf(v, -FLT_MAX), FLT_MAX);
}
__device__ __forceinline__ int reflect(int i, int n) {  // scipy.ndimage mode='reflect'
  while (i < 0 || i >= n) i = i < 0 ? -i - 1 : 2 * n - i - 1;
  return i;
}

// Median of the (2 rz + 1) x 3 x 3 window of component `c` (NaN -> 0) around (z, y, x).
__device__ float median_window(const float* __restrict__ f, int nz, int ny, int nx, int rz,
                               int z, int y, int x) {
  float v[27];
  int n = 0;
  for (int dz = -rz; dz <= rz; ++dz) {
    const int zz = reflect(z + dz, nz);
    for (int dy = -1; dy <= 1; ++dy) {
      const int yy = reflect(y + dy, ny);
      for (int dx = -1; dx <= 1; ++dx) {
        const int xx = reflect(x + dx, nx);
        v[n++] = nan_to_num(f[((size_t)zz * ny + yy) * nx + xx]);
      }
    }
  }
  // selection up to the middle element
  const int mid = n / 2;
  for (int i = 0; i <= mid; ++i) {
    int m = i;
    for (int j = i + 1; j < n; ++j) m = v[j] < v[m] ? j : m;
    const float t = v[i]; v[i] = v[m]; v[m] = t;
  }
  return v[mid];
}

// flow [nc][nz][ny][nx] with nc = dim or dim + 2 -> out [dim][nz][ny][nx]
__global__ void __launch_bounds__(kThreads)
clean_flow_kernel(const float* __restrict__ flow, int nc, int dim, int nz, int ny, int nx,
                  float min_peak_ratio, float min_peak_sharpness, float max_magnitude,
                  float max_deviation, float* __restrict__ out) {
  const size_t n = (size_t)nz * ny * nx;
  const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  const int x = (int)(i % nx), y = (int)((i / nx) % ny), z = (int)(i / ((size_t)nx * ny));
  bool reject = false;
  if (nc == dim + 2) {
    const float sharp = fabsf(flow[(size_t)dim * n + i]);
    const float ratio = fabsf(flow[(size_t)(dim + 1) * n + i]);
    reject = (sharp < min_peak_sharpness) || (ratio > 0.0f && ratio < min_peak_ratio);
  }
  float vec[3];
  for (int c = 0; c < dim; ++c) vec[c] = flow[(size_t)c * n + i];
  if (max_magnitude > 0.f) {
    float m = fabsf(vec[0]);  // np.max over components propagates NaN; NaN > t is False
    bool has_nan = vec[0] != vec[0];
    for (int c = 1; c < dim; ++c) { m = fmaxf(m, fabsf(vec[c])); has_nan |= vec[c] != vec[c]; }
    reject |= !has_nan && m > max_magnitude;
  }
  if (max_deviation > 0.f) {
    float m = 0.f;
    bool has_nan = false;
    for (int c = 0; c < dim; ++c) {
      const float med = median_window(flow + (size_t)c * n, nz, ny, nx, dim == 3 ? 1 : 0, z, y, x);
      const float d = fabsf(med - vec[c]);
      has_nan |= d != d;
      m = fmaxf(m, d);
    }
    reject |= !has_nan && m > max_deviation;
  }
  for (int c = 0; c < dim; ++c) out[(size_t)c * n + i] = reject ? NAN : vec[c];
}

// ---- reconcile_flows ---------------------------------------------------------------
// out[still invalid] = other, flow_utils.py:101-107
__global__ void __launch_bounds__(kThreads)
fill_kernel(float* __restrict__ out, const float* __restrict__ other, int nc, size_t n,
            float min_delta_z) {
  const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  const float o0 = out[i];
  bool fill = o0 != o0;
  if (nc == 3) fill = fill && fabsf(other[2 * n + i]) >= min_delta_z;
  if (!fill) return;
  for (int c = 0; c < nc; ++c) out[(size_t)c * n + i] = other[(size_t)c * n + i];
}

// steep (x component along x, y component along y, borders against 0), flow_utils.py:110-116
__global__ void __launch_bounds__(kThreads)
gradient_mask_kernel(const float* __restrict__ f, int nz, int ny, int nx, float max_gradient,
                     unsigned char* __restrict__ mask) {
  const size_t n = (size_t)nz * ny * nx;
  const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  const int x = (int)(i % nx), y = (int)((i / nx) % ny);
  const float* fx = f;
  const float* fy = f + n;
  const float cx = fx[i], cy = fy[i];
  const float lx = x > 0 ? fx[i - 1] : 0.f, rx = x + 1 < nx ? fx[i + 1] : 0.f;
  const float uy = y > 0 ? fy[i - nx] : 0.f, dy = y + 1 < ny ? fy[i + nx] : 0.f;
  const bool steep = fabsf(cx - lx) > max_gradient || fabsf(rx - cx) > max_gradient ||
                     fabsf(cy - uy) > max_gradient || fabsf(dy - cy) > max_gradient;
  mask[i] = steep ? 1 : 0;
}

// |3 x 3 median - value| of the first two components, flow_utils.py:117-119
__global__ void __launch_bounds__(kThreads)
deviation_mask_kernel(const float* __restrict__ f, int nz, int ny, int nx, float max_deviation,
                      unsigned char* __restrict__ mask) {
  const size_t n = (size_t)nz * ny * nx;
  const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  const int x = (int)(i % nx), y = (int)((i / nx) % ny), z = (int)(i / ((size_t)nx * ny));
  float m = 0.f;
  bool has_nan = false;
  for (int c = 0; c < 2; ++c) {
    const float med = median_window(f + (size_t)c * n, nz, ny, nx, 0, z, y, x);
    const float d = fabsf(med - f[(size_t)c * n + i]);
    has_nan |= d != d;
    m = fmaxf(m, d);
  }
  mask[i] = (!has_nan && m > max_deviation) ? 1 : 0;
}

__global__ void __launch_bounds__(kThreads)
apply_mask_kernel(float* __restrict__ f, int nc, size_t n, const unsigned char* __restrict__ mask) {
  const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n || !mask[i]) return;
  for (int c = 0; c < nc; ++c) f[(size_t)c * n + i] = NAN;
}

// Connected components of the valid pixels of every section (4-connectivity, as
// scipy.ndimage.label's default structure): union-find with atomic min hooking.
__device__ int find_root(const int* lab, int i) {
  while (true) {
    const int p = reinterpret_cast<const volatile int*>(lab)[i];  // others hook concurrently
    if (p == i) return i;
    i = p;
  }
}
__device__ void unite(int* lab, int a, int b) {
  while (true) {
    a = find_root(lab, a);
    b = find_root(lab, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }  // hook the larger root under the smaller
    const int old = atomicMin(&lab[a], b);
    if (old == a) return;
    a = old;
  }
}
__global__ void __launch_bounds__(kThreads)
ccl_init_kernel(const float* __restrict__ f, int nc, size_t n, int* lab) {
  const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  bool valid = true;
  for (int c = 0; c < nc; ++c) valid &= !(f[(size_t)c * n + i] != f[(size_t)c * n + i]);
  lab[i] = valid ? (int)i : -1;
}
__global__ void __launch_bounds__(kThreads)
ccl_merge_kernel(int* lab, int ny, int nx, size_t n) {
  const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n || lab[i] < 0) return;
  const int x = (int)(i % nx), y = (int)((i / nx) % ny);
  if (x > 0 && lab[i - 1] >= 0) unite(lab, (int)i, (int)i - 1);
  if (y > 0 && lab[i - nx] >= 0) unite(lab, (int)i, (int)i - nx);
}
__global__ void __launch_bounds__(kThreads)
ccl_count_kernel(int* lab, int* size, size_t n) {
  const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n || lab[i] < 0) return;
  const int r = find_root(lab, (int)i);
  atomicAdd(&size[r], 1);
}
__global__ void __launch_bounds__(kThreads)
ccl_small_kernel(const int* lab, const int* size, size_t n, int min_patch_size,
                 unsigned char* mask) {
  const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  bool tiny = false;
  if (lab[i] >= 0) tiny = size[find_root(lab, (int)i)] < min_patch_size;
  mask[i] = tiny ? 1 : 0;
}

// ---- mask_irregular -----------------------------------------------------------------
// The reference adds the stride (an int64 / float64 NumPy scalar) to the fp32 differences
// and compares with frac * stride, i.e. everything after np.diff happens in float64.
__global__ void __launch_bounds__(kThreads)
irregular_bad_kernel(const float* __restrict__ m, int ny, int nx, double sx, double sy,
                     double lo_x, double lo_y, double hi_x, double hi_y,
                     unsigned char* __restrict__ bad) {
  const int n = ny * nx;
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  const int x = i % nx, y = i / nx;
  const float* mx = m;
  const float* my = m + n;
  const double dx = (double)(x + 1 < nx ? mx[i + 1] - mx[i] : 0.f) + sx;
  const double dy = (double)(y + 1 < ny ? my[i + nx] - my[i] : 0.f) + sy;
  bad[i] = (dx < lo_x || dy < lo_y || dx > hi_x || dy > hi_y) ? 1 : 0;
}
__global__ void __launch_bounds__(kThreads)
irregular_apply_kernel(float* __restrict__ m, int ny, int nx, int radius,
                       const unsigned char* __restrict__ bad, unsigned char* __restrict__ out) {
  const int n = ny * nx;
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  const int x = i % nx, y = i / nx;
  bool b = false;
  for (int dy = -radius; dy <= radius && !b; ++dy) {
    const int yy = y + dy;
    if (yy < 0 || yy >= ny) continue;
    for (int dx = -radius; dx <= radius; ++dx) {
      const int xx = x + dx;
      if (xx < 0 || xx >= nx) continue;
      if (bad[yy * nx + xx]) { b = true; break; }
    }
  }
  out[i] = b ? 1 : 0;
  if (b) { m[i] = NAN; m[n + i] = NAN; }
}

static unsigned int blocks_for(size_t n) { return (unsigned int)((n + kThreads - 1) / kThreads); }

}  // namespace filt
}  // namespace sofima

extern "C" {

int sofima_clean_flow(sofima_ctx* ctx, const float* flow, int nc, int dim, const int64_t* zyx,
                      float min_peak_ratio, float min_peak_sharpness, float max_magnitude,
                      float max_deviation, float* out) {
  using namespace sofima;
  if (!ctx) return fail(nullptr, SOFIMA_EINVAL, "ctx is NULL");
  if (!flow || !out || !zyx) return fail(ctx, SOFIMA_EINVAL, "NULL argument");
  if ((dim != 2 && dim != 3) || nc < dim || nc > dim + 2 || nc == dim + 1)
    return fail(ctx, SOFIMA_EINVAL, "clean_flow: %d channels for dim %d", nc, dim);
  const size_t n = (size_t)zyx[0] * zyx[1] * zyx[2];
  if (n == 0) return SOFIMA_OK;
  if (zyx[0] > INT32_MAX || zyx[1] > INT32_MAX || zyx[2] > INT32_MAX)
    return fail(ctx, SOFIMA_EINVAL, "flow field too large");
  DeviceGuard guard(ctx->device);
  LaunchTimer timer(ctx, "clean_flow");
  filt::clean_flow_kernel<<<filt::blocks_for(n), filt::kThreads, 0, ctx->stream>>>(
      flow, nc, dim, (int)zyx[0], (int)zyx[1], (int)zyx[2], min_peak_ratio, min_peak_sharpness,
      max_magnitude, max_deviation, out);
  SOFIMA_CHECK_LAUNCH(ctx);
  return SOFIMA_OK;
}

int sofima_reconcile_flows(sofima_ctx* ctx, float* out, const float* const* others, int nothers,
                           int nc, const int64_t* zyx, float max_gradient, float max_deviation,
                           int min_patch_size, float min_delta_z) {
  using namespace sofima;
  if (!ctx) return fail(nullptr, SOFIMA_EINVAL, "ctx is NULL");
  if (!out || !zyx || (nothers > 0 && !others)) return fail(ctx, SOFIMA_EINVAL, "NULL argument");
  if (nc != 2 && nc != 3) return fail(ctx, SOFIMA_EINVAL, "reconcile_flows: 2 or 3 channels");
  const size_t n = (size_t)zyx[0] * zyx[1] * zyx[2];
  if (n == 0) return SOFIMA_OK;
  if (n > (size_t)INT32_MAX) return fail(ctx, SOFIMA_EINVAL, "flow field too large");
  DeviceGuard guard(ctx->device);
  const int nz = (int)zyx[0], ny = (int)zyx[1], nx = (int)zyx[2];
  const unsigned int nb = filt::blocks_for(n);
  void *mask = nullptr, *lab = nullptr;
  int rc;
  if ((rc = scratch(ctx, "filt.mask", n, &mask))) return rc;
  LaunchTimer timer(ctx, "reconcile_flows");
  for (int k = 0; k < nothers; ++k) {
    filt::fill_kernel<<<nb, filt::kThreads, 0, ctx->stream>>>(out, others[k], nc, n, min_delta_z);
    SOFIMA_CHECK_LAUNCH(ctx);
  }
  unsigned char* m = static_cast<unsigned char*>(mask);
  if (max_gradient > 0.f) {
    filt::gradient_mask_kernel<<<nb, filt::kThreads, 0, ctx->stream>>>(out, nz, ny, nx,
                                                                       max_gradient, m);
    SOFIMA_CHECK_LAUNCH(ctx);
    filt::apply_mask_kernel<<<nb, filt::kThreads, 0, ctx->stream>>>(out, nc, n, m);
    SOFIMA_CHECK_LAUNCH(ctx);
  }
  if (max_deviation > 0.f) {
    filt::deviation_mask_kernel<<<nb, filt::kThreads, 0, ctx->stream>>>(out, nz, ny, nx,
                                                                        max_deviation, m);
    SOFIMA_CHECK_LAUNCH(ctx);
    filt::apply_mask_kernel<<<nb, filt::kThreads, 0, ctx->stream>>>(out, nc, n, m);
    SOFIMA_CHECK_LAUNCH(ctx);
  }
  if (min_patch_size > 0) {
    if ((rc = scratch(ctx, "filt.labels", 2 * n * sizeof(int), &lab))) return rc;
    int* labels = static_cast<int*>(lab);
    int* sizes = labels + n;
    SOFIMA_CUDA(ctx, cudaMemsetAsync(sizes, 0, n * sizeof(int), ctx->stream));
    filt::ccl_init_kernel<<<nb, filt::kThreads, 0, ctx->stream>>>(out, nc, n, labels);
    SOFIMA_CHECK_LAUNCH(ctx);
    filt::ccl_merge_kernel<<<nb, filt::kThreads, 0, ctx->stream>>>(labels, ny, nx, n);
    SOFIMA_CHECK_LAUNCH(ctx);
    filt::ccl_count_kernel<<<nb, filt::kThreads, 0, ctx->stream>>>(labels, sizes, n);
    SOFIMA_CHECK_LAUNCH(ctx);
    filt::ccl_small_kernel<<<nb, filt::kThreads, 0, ctx->stream>>>(labels, sizes, n,
                                                                   min_patch_size, m);
    SOFIMA_CHECK_LAUNCH(ctx);
    filt::apply_mask_kernel<<<nb, filt::kThreads, 0, ctx->stream>>>(out, nc, n, m);
    SOFIMA_CHECK_LAUNCH(ctx);
  }
  return SOFIMA_OK;
}

int sofima_mask_irregular(sofima_ctx* ctx, float* coord_map, int64_t ny, int64_t nx,
                          const double* stride_xy, double frac, double max_frac,
                          int dilation_iters, uint8_t* out_mask) {
  using namespace sofima;
  if (!ctx) return fail(nullptr, SOFIMA_EINVAL, "ctx is NULL");
  if (!coord_map || !stride_xy || !out_mask) return fail(ctx, SOFIMA_EINVAL, "NULL argument");
  if (ny * nx == 0) return SOFIMA_OK;
  if (ny * nx > INT32_MAX) return fail(ctx, SOFIMA_EINVAL, "map too large");
  if (dilation_iters < 0) return fail(ctx, SOFIMA_EINVAL, "dilation_iters < 0");
  DeviceGuard guard(ctx->device);
  const size_t n = (size_t)(ny * nx);
  void* bad = nullptr;
  int rc;
  if ((rc = scratch(ctx, "filt.mask", n, &bad))) return rc;
  LaunchTimer timer(ctx, "mask_irregular");
  filt::irregular_bad_kernel<<<filt::blocks_for(n), filt::kThreads, 0, ctx->stream>>>(
      coord_map, (int)ny, (int)nx, stride_xy[0], stride_xy[1], frac * stride_xy[0],
      frac * stride_xy[1], max_frac * stride_xy[0], max_frac * stride_xy[1],
      static_cast<unsigned char*>(bad));
  SOFIMA_CHECK_LAUNCH(ctx);
  filt::irregular_apply_kernel<<<filt::blocks_for(n), filt::kThreads, 0, ctx->stream>>>(
      coord_map, (int)ny, (int)nx, dilation_iters, static_cast<const unsigned char*>(bad),
      out_mask);
  SOFIMA_CHECK_LAUNCH(ctx);
  return SOFIMA_OK;
}

}  // extern "C"
