// Image warping through a coordinate map on B200.
//
// Replaces the per-voxel work of warp.ndimage_warp (reference warp.py:189-335): for
// every output voxel the (absolute, source-unit) coordinate map is interpolated
// linearly at (index - offset) / stride (warp.py:300-306) and the image is sampled at
// the resulting position (warp.py:309).  Both samplings are
// scipy.ndimage.map_coordinates(order 0 / 1, mode='constant', cval=0) in the
// reference; this kernel follows that routine operation by operation in float64
// (compiled with -fmad=false): a sample with any coordinate outside [0, n - 1] is 0,
// per axis i0 = floor(c), t = c - i0, weights (1 - t, t), corners visited with the last
// axis fastest, term = ((value * w0) * w1) * w2, terms added in visiting order; unsigned
// outputs are (T) min(t + 0.5, max) for t > 0 and 0 otherwise.  The result equals SciPy's
// bit for bit (tests/test_warp_gpu.py).  The reference's box tiling (work_size, overlap,
// parallelism) only bounds host memory and does not change any output value.
//
// One thread per output voxel, x fastest: coalesced stores, the coordinate map (one
// node per `stride` voxels) stays in L1 / L2, the image gather is the HBM traffic.
#include "common.cuh"

namespace sofima {
namespace warp {

constexpr int kThreads = 256;

struct Params {
  const double* src_map;  // [DIM][m zyx] absolute source coordinates, xyz components
  const void* image;
  void* out;
  int m[3], img[3], outn[3];      // zyx extents (2-d: [0] = 1)
  double offset[3], stride[3];    // zyx
  int order, dtype;
};

template <int DIM>
__device__ __forceinline__ bool inside(const double (&c)[3], const int (&n)[3]) {
  bool ok = true;
#pragma unroll
  for (int a = 3 - DIM; a < 3; ++a) ok = ok && c[a] >= 0.0 && c[a] <= (double)(n[a] - 1);
  return ok;
}

// Linear interpolation weights / base indices of SciPy's order-1 spline.
template <int DIM>
__device__ __forceinline__ void nodes(const double (&c)[3], const int (&n)[3], int (&i0)[3],
                                      int (&i1)[3], double (&w0)[3], double (&w1)[3]) {
#pragma unroll
  for (int a = 3 - DIM; a < 3; ++a) {
    const double f = floor(c[a]);
    const double t = c[a] - f;
    i0[a] = (int)f;
    i1[a] = min(i0[a] + 1, n[a] - 1);
    w0[a] = 1.0 - t;
    w1[a] = t;
  }
}

template <int DIM, typename Load>
__device__ __forceinline__ double interp1(const double (&c)[3], const int (&n)[3], Load load) {
  int i0[3] = {0, 0, 0}, i1[3] = {0, 0, 0};
  double w0[3] = {1.0, 1.0, 1.0}, w1[3] = {0.0, 0.0, 0.0};
  nodes<DIM>(c, n, i0, i1, w0, w1);
  double acc = 0.0;
#pragma unroll
  for (int k = 0; k < (1 << DIM); ++k) {
    const int bz = DIM == 3 ? (k >> 2) & 1 : 0, by = (k >> 1) & 1, bx = k & 1;
    const int z = DIM == 3 ? (bz ? i1[0] : i0[0]) : 0;
    const int y = by ? i1[1] : i0[1], x = bx ? i1[2] : i0[2];
    double term = load(z, y, x);
    if (DIM == 3) term = term * (bz ? w1[0] : w0[0]);
    term = term * (by ? w1[1] : w0[1]);
    term = term * (bx ? w1[2] : w0[2]);
    acc = acc + term;
  }
  return acc;
}

template <typename T>
__device__ __forceinline__ T to_output(double t);
template <>
__device__ __forceinline__ float to_output<float>(double t) { return (float)t; }
template <>
__device__ __forceinline__ uint8_t to_output<uint8_t>(double t) {
  return t > 0.0 ? (uint8_t)fmin(t + 0.5, 255.0) : (uint8_t)0;
}
template <>
__device__ __forceinline__ uint16_t to_output<uint16_t>(double t) {
  return t > 0.0 ? (uint16_t)fmin(t + 0.5, 65535.0) : (uint16_t)0;
}
template <>
__device__ __forceinline__ uint32_t to_output<uint32_t>(double t) {
  return t > 0.0 ? (uint32_t)fmin(t + 0.5, 4294967295.0) : 0u;
}

template <int DIM, typename T>
__global__ void __launch_bounds__(kThreads) warp_kernel(const Params p) {
  const int x = blockIdx.x * kThreads + threadIdx.x;
  const int y = blockIdx.y, z = blockIdx.z;
  if (x >= p.outn[2]) return;
  const int idx[3] = {z, y, x};
  // position in the coordinate map (warp.py:300-301)
  double mc[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int a = 3 - DIM; a < 3; ++a) mc[a] = ((double)idx[a] - p.offset[a]) / p.stride[a];
  const long long mvol = (long long)p.m[0] * p.m[1] * p.m[2];
  double dense[3] = {0.0, 0.0, 0.0};
  if (inside<DIM>(mc, p.m)) {
#pragma unroll
    for (int a = 3 - DIM; a < 3; ++a) {
      const double* comp = p.src_map + (long long)(2 - a) * mvol;  // z <- [2], y <- [1], x <- [0]
      dense[a] = interp1<DIM>(mc, p.m, [&](int zz, int yy, int xx) {
        return __ldg(comp + ((long long)zz * p.m[1] + yy) * p.m[2] + xx);
      });
    }
  }  // outside the map: all coordinates 0 (cval), i.e. the first image voxel
  const T* img = static_cast<const T*>(p.image);
  auto pixel = [&](int zz, int yy, int xx) {
    return (double)__ldg(img + ((long long)zz * p.img[1] + yy) * p.img[2] + xx);
  };
  double t = 0.0;
  if (inside<DIM>(dense, p.img)) {
    if (p.order == 0) {
      int q[3] = {0, 0, 0};
#pragma unroll
      for (int a = 3 - DIM; a < 3; ++a)
        q[a] = min(max((int)floor(dense[a] + 0.5), 0), p.img[a] - 1);
      t = pixel(q[0], q[1], q[2]);
    } else {
      t = interp1<DIM>(dense, p.img, pixel);
    }
  }
  static_cast<T*>(p.out)[((long long)z * p.outn[1] + y) * p.outn[2] + x] = to_output<T>(t);
}

template <int DIM>
static int launch(sofima_ctx* ctx, const Params& p) {
  const dim3 grid((unsigned)ceil_div(p.outn[2], kThreads), (unsigned)p.outn[1],
                  (unsigned)p.outn[0]);
  switch (p.dtype) {
    case SOFIMA_U8: warp_kernel<DIM, uint8_t><<<grid, kThreads, 0, ctx->stream>>>(p); break;
    case SOFIMA_F32: warp_kernel<DIM, float><<<grid, kThreads, 0, ctx->stream>>>(p); break;
    case SOFIMA_U16: warp_kernel<DIM, uint16_t><<<grid, kThreads, 0, ctx->stream>>>(p); break;
    case SOFIMA_U32: warp_kernel<DIM, uint32_t><<<grid, kThreads, 0, ctx->stream>>>(p); break;
    default: return fail(ctx, SOFIMA_EINVAL, "unsupported image dtype %d", p.dtype);
  }
  return SOFIMA_OK;
}

}  // namespace warp
}  // namespace sofima

extern "C" int sofima_warp_image(sofima_ctx* ctx, int dim, const void* image, int img_dtype,
                                 const int64_t* image_shape, const double* src_map,
                                 const int64_t* map_shape, const double* offset,
                                 const double* stride, int order, void* out,
                                 const int64_t* out_shape) {
  using namespace sofima;
  if (!ctx) return fail(nullptr, SOFIMA_EINVAL, "ctx is NULL");
  if (dim != 2 && dim != 3) return fail(ctx, SOFIMA_EINVAL, "dim must be 2 or 3 (got %d)", dim);
  if (order != 0 && order != 1)
    return fail(ctx, SOFIMA_EUNSUPPORTED, "interpolation order %d: only 0 and 1 are built", order);
  if (!image_shape || !map_shape || !out_shape || !offset || !stride)
    return fail(ctx, SOFIMA_EINVAL, "NULL argument");
  warp::Params p;
  memset(&p, 0, sizeof(p));
  long long nout = 1;
  for (int a = 0; a < 3; ++a) {
    const int j = a - (3 - dim);
    const int64_t is = j >= 0 ? image_shape[j] : 1, ms = j >= 0 ? map_shape[j] : 1,
                  os = j >= 0 ? out_shape[j] : 1;
    if (is < 1 || ms < 1 || os < 0 || is > INT32_MAX || ms > INT32_MAX || os > INT32_MAX ||
        (a < 2 && os > 65535))
      return fail(ctx, SOFIMA_EINVAL, "extent out of range");
    p.img[a] = (int)is; p.m[a] = (int)ms; p.outn[a] = (int)os;
    p.offset[a] = j >= 0 ? offset[j] : 0.0;
    p.stride[a] = j >= 0 ? stride[j] : 1.0;
    if (j >= 0 && !(p.stride[a] != 0.0)) return fail(ctx, SOFIMA_EINVAL, "stride must be non-zero");
    nout *= os;
  }
  if (nout == 0) return SOFIMA_OK;
  if (!image || !src_map || !out) return fail(ctx, SOFIMA_EINVAL, "NULL array argument");
  p.image = image; p.src_map = src_map; p.out = out;
  p.order = order; p.dtype = img_dtype;
  DeviceGuard guard(ctx->device);
  LaunchTimer timer(ctx, "warp_image");
  const int rc = dim == 2 ? warp::launch<2>(ctx, p) : warp::launch<3>(ctx, p);
  if (rc) return rc;
  SOFIMA_CHECK_LAUNCH(ctx);
  return SOFIMA_OK;
}
