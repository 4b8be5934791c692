// Section-wise image warping through an xy coordinate map on B200.
//
// Replaces the per-pixel work of warp.warp_subvolume (reference warp.py:58-186).  For
// every output pixel the reference
//   1. densifies the coarse map with scipy's RegularGridInterpolator (linear,
//      extrapolating) in float64 and casts to float32      (warp.py:144-153),
//   2. quantises the coordinates with cv2.convertMaps to int16 + 5 fraction bits
//      (warp.py:155-160),
//   3. samples the section with cv2.remap, constant border 0 (warp.py:162-165).
// This kernel does the three steps per thread with the same arithmetic (compiled with
// -fmad=false), so the output equals the reference's bit for bit:
//   * densify: interval i = clip(#{g <= q} - 1, 0, n - 2), t = (q - g[i]) / (g[i+1] - g[i]);
//     float64 maps add ((v * wy) * wx) over the corners 00, 01, 10, 11, other maps
//     v * (wy * wx) (the two evaluation routes inside scipy);
//   * quantise: rint (half to even) of c * 32 (c for nearest neighbour), INT_MIN when NaN
//     or not representable, >> 5 saturated to int16, & 31 as the table row;
//   * sample: uint8 images use OpenCV's 15-bit integer tables (2-d products rounded, sum
//     forced to 32768), (sum + 2^14) >> 15; other types use float32 products of the 1-d
//     coefficient rows, accumulated row by row when the footprint is inside the image and
//     tap by tap otherwise; integer outputs are rounded half to even and saturated.
// The coefficient tables are built once on the host by the published formulas (linear,
// cubic with A = -0.75, Lanczos-4) and kept in device memory.
//
// One thread per output pixel, x fastest (coalesced stores); the map and the tables
// stay in L1 / L2, the image gather is the HBM traffic.  All channels of a pixel share
// the quantised coordinate.
#include <cfloat>
#include <cmath>
#include <mutex>

#include "common.cuh"

namespace sofima {
namespace warpcv {

constexpr int kTab = 32;  // 5 fraction bits

struct Tables {
  float tab1[3][kTab][8];  // [linear, cubic, lanczos][fraction][tap]
  int16_t itab4[kTab * kTab][16];
  int16_t itab8[kTab * kTab][64];
};

static void coeffs(int method, float x, float* c) {
  if (method == 0) {
    c[0] = 1.f - x;
    c[1] = x;
  } else if (method == 1) {
    const float a = -0.75f;
    const float x1 = x + 1.f, xm = 1.f - x;
    float t = a * x1;
    t = t - 5.f * a; t = t * x1; t = t + 8.f * a; t = t * x1; t = t - 4.f * a;
    c[0] = t;
    float u = (a + 2.f) * x; u = u - (a + 3.f); u = u * x; u = u * x; u = u + 1.f;
    c[1] = u;
    float v = (a + 2.f) * xm; v = v - (a + 3.f); v = v * xm; v = v * xm; v = v + 1.f;
    c[2] = v;
    float r = 1.f - c[0]; r = r - c[1]; r = r - c[2];
    c[3] = r;
  } else {
    const double s45 = 0.70710678118654752440084436210485;
    const double cs[8][2] = {{1, 0}, {-s45, -s45}, {0, 1}, {s45, -s45},
                             {-1, 0}, {s45, s45}, {0, -1}, {-s45, s45}};
    if (x < FLT_EPSILON) {
      for (int i = 0; i < 8; ++i) c[i] = 0.f;
      c[3] = 1.f;
      return;
    }
    const double pi = 3.1415926535897932384626433832795;
    const double y0 = -((double)x + 3.0) * pi * 0.25, s0 = std::sin(y0), c0 = std::cos(y0);
    float sum = 0.f;
    for (int i = 0; i < 8; ++i) {
      const double y = -((double)x + 3.0 - i) * pi * 0.25;
      c[i] = (float)((cs[i][0] * s0 + cs[i][1] * c0) / (y * y));
      sum = sum + c[i];
    }
    sum = 1.f / sum;
    for (int i = 0; i < 8; ++i) c[i] = c[i] * sum;
  }
}

template <int K>
static void int_table(const float (*tab1)[8], int16_t (*itab)[K * K]) {
  for (int fy = 0; fy < kTab; ++fy)
    for (int fx = 0; fx < kTab; ++fx) {
      int16_t* t = itab[fy * kTab + fx];
      int sum = 0;
      for (int i = 0; i < K; ++i)
        for (int j = 0; j < K; ++j) {
          const float v = tab1[fy][i] * tab1[fx][j];
          long r = std::lrint((double)(v * 32768.f));  // half to even
          r = r < -32768 ? -32768 : (r > 32767 ? 32767 : r);
          t[i * K + j] = (int16_t)r;
          sum += (int)r;
        }
      const int diff = sum - 32768;
      if (diff == 0) continue;
      const int c0 = K / 2;
      int lo = c0 * K + c0, hi = lo;
      for (int a = c0; a < c0 + 2; ++a)
        for (int b = c0; b < c0 + 2; ++b) {
          if (t[a * K + b] < t[lo]) lo = a * K + b;
          else if (t[a * K + b] > t[hi]) hi = a * K + b;
        }
      if (diff < 0) t[hi] = (int16_t)(t[hi] - diff);
      else t[lo] = (int16_t)(t[lo] - diff);
    }
}

static const Tables& host_tables() {
  static Tables* tabs = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    tabs = new Tables();
    memset(tabs, 0, sizeof(Tables));
    for (int m = 0; m < 3; ++m)
      for (int f = 0; f < kTab; ++f) coeffs(m, (float)f * (1.f / kTab), tabs->tab1[m][f]);
    int_table<4>(tabs->tab1[1], tabs->itab4);
    int_table<8>(tabs->tab1[2], tabs->itab8);
  });
  return *tabs;
}

struct Params {
  const void* image;     // [n][nz][ih][iw]
  void* out;             // [n][nz][oh][ow]
  const double* abs_map; // [2][nz][my][mx]: x then y source coordinate of every node
  const int* iy;         // [oh] interval of every output row in the node grid
  const int* ix;         // [ow]
  const double* ty;      // [oh] normalised distance within the interval
  const double* tx;      // [ow]
  const uint8_t* skip;   // [nz] or null
  const Tables* tabs;
  int n, nz, ih, iw, oh, ow, my, mx;
  int map_f64;
};

// Interval and normalised distance of every output row / column in the node grid: they
// depend on one axis only, so a small kernel tabulates them once per call and the per-pixel
// kernel is free of float64 divisions.
__global__ void axis_table_kernel(const double* __restrict__ g, int n, int count,
                                  int* __restrict__ idx, double* __restrict__ dist) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= count) return;
  const double g0 = g[0], g1 = g[1];
  const double guess = floor(((double)q - g0) / (g1 - g0));
  int i = guess < 0.0 ? 0 : (guess > (double)(n - 2) ? n - 2 : (int)guess);
  while (i > 0 && (double)q < g[i]) --i;
  while (i < n - 2 && (double)q >= g[i + 1]) ++i;
  idx[q] = i;
  dist[q] = ((double)q - g[i]) / (g[i + 1] - g[i]);
}

__device__ __forceinline__ double densify(const double* __restrict__ v, int mx, int iy, int ix,
                                          double ty, double tx, bool f64) {
  const double v00 = __ldg(v + (size_t)iy * mx + ix), v01 = __ldg(v + (size_t)iy * mx + ix + 1);
  const double v10 = __ldg(v + (size_t)(iy + 1) * mx + ix),
               v11 = __ldg(v + (size_t)(iy + 1) * mx + ix + 1);
  const double wy0 = 1.0 - ty, wx0 = 1.0 - tx;
  double acc = 0.0;
  if (f64) {
    acc = acc + (v00 * wy0) * wx0;
    acc = acc + (v01 * wy0) * tx;
    acc = acc + (v10 * ty) * wx0;
    acc = acc + (v11 * ty) * tx;
  } else {
    acc = acc + v00 * (wy0 * wx0);
    acc = acc + v01 * (wy0 * tx);
    acc = acc + v10 * (ty * wx0);
    acc = acc + v11 * (ty * tx);
  }
  return acc;
}

// cvRound on x86: nearest even; the "integer indefinite" value when not representable.
__device__ __forceinline__ int cv_round(float v) {
  if (!(fabsf(v) < 2147483648.f)) return INT_MIN;
  return __float2int_rn(v);
}

__device__ __forceinline__ int sat16(int v) { return max(-32768, min(32767, v)); }

template <typename T> __device__ __forceinline__ T saturate_out(float v);
template <> __device__ __forceinline__ float saturate_out<float>(float v) { return v; }
template <> __device__ __forceinline__ uint16_t saturate_out<uint16_t>(float v) {
  return (uint16_t)max(0, min(65535, cv_round(v)));
}
template <> __device__ __forceinline__ int16_t saturate_out<int16_t>(float v) {
  return (int16_t)sat16(cv_round(v));
}

// acc + (two signed 16-bit weights) . (low / high two unsigned bytes of `px`)
__device__ __forceinline__ int dp2a_lo(int w, uint32_t px, int acc) {
  int d;
  asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(px), "r"(acc));
  return d;
}
__device__ __forceinline__ int dp2a_hi(int w, uint32_t px, int acc) {
  int d;
  asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(px), "r"(acc));
  return d;
}

// Integer-table sampling of a uint8 section.
template <int K>
__device__ __forceinline__ uint8_t sample_u8(const uint8_t* __restrict__ img, int ih, int iw,
                                             int sx, int sy, int fx, int fy, const Tables* tabs) {
  const int x0 = sx - (K / 2 - 1), y0 = sy - (K / 2 - 1);
  int acc = 0;
  if (K == 2) {
    const int w[4] = {(32 - fx) * (32 - fy) * 32, fx * (32 - fy) * 32, (32 - fx) * fy * 32,
                      fx * fy * 32};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int yy = y0 + i;
      if ((unsigned)yy >= (unsigned)ih) continue;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int xx = x0 + j;
        if ((unsigned)xx < (unsigned)iw) acc += w[i * 2 + j] * (int)__ldg(img + (size_t)yy * iw + xx);
      }
    }
  } else {
    const int16_t* w = K == 4 ? tabs->itab4[fy * kTab + fx] : tabs->itab8[fy * kTab + fx];
    const bool inner = x0 >= 0 && x0 + K <= iw && y0 >= 0 && y0 + K <= ih;
    if (inner && K == 8) {
      // Footprint inside the section: every row is 8 consecutive bytes, fetched as the (2 or
      // 3) aligned words around them; the 16-bit weights are consumed pairwise by dp2a.
      const uint8_t* s = img + (size_t)y0 * iw + x0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(s + (size_t)i * iw);
        const uint32_t* wp = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
        const unsigned sh = (unsigned)(a & 3) * 8;
        const uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = sh ? __ldg(wp + 2) : 0u;
        const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
        const int4 q = __ldg(reinterpret_cast<const int4*>(w + i * 8));
        acc = dp2a_lo(q.x, lo, acc);
        acc = dp2a_hi(q.y, lo, acc);
        acc = dp2a_lo(q.z, hi, acc);
        acc = dp2a_hi(q.w, hi, acc);
      }
    } else if (inner) {
      const uint8_t* s = img + (size_t)y0 * iw + x0;
#pragma unroll
      for (int i = 0; i < K; ++i) {
        int wi[K];
#pragma unroll
        for (int j = 0; j < K; j += 2) {
          const int pr = __ldg(reinterpret_cast<const int*>(w + i * K + j));
          wi[j] = (int)(int16_t)(pr & 0xffff);
          wi[j + 1] = pr >> 16;
        }
#pragma unroll
        for (int j = 0; j < K; ++j) acc += wi[j] * (int)__ldg(s + (size_t)i * iw + j);
      }
    } else {
      for (int i = 0; i < K; ++i) {
        const int yy = y0 + i;
        if ((unsigned)yy >= (unsigned)ih) continue;
        for (int j = 0; j < K; ++j) {
          const int xx = x0 + j;
          if ((unsigned)xx < (unsigned)iw)
            acc += (int)w[i * K + j] * (int)__ldg(img + (size_t)yy * iw + xx);
        }
      }
    }
  }
  return (uint8_t)max(0, min(255, (acc + (1 << 14)) >> 15));
}

// Float-table sampling (uint16, int16, float32 sections).
template <typename T, int K>
__device__ __forceinline__ T sample_f(const T* __restrict__ img, int ih, int iw, int sx, int sy,
                                      int fx, int fy, const Tables* tabs) {
  const int x0 = sx - (K / 2 - 1), y0 = sy - (K / 2 - 1);
  constexpr int M = K == 2 ? 0 : (K == 4 ? 1 : 2);
  float wy[K], wx[K];
#pragma unroll
  for (int i = 0; i < K; ++i) {
    wy[i] = tabs->tab1[M][fy][i];
    wx[i] = tabs->tab1[M][fx][i];
  }
  const bool inner = x0 >= 0 && x0 + K <= iw && y0 >= 0 && y0 + K <= ih;
  float total = 0.f;
  if (inner && K > 2) {
    const T* s = img + (size_t)y0 * iw + x0;
#pragma unroll
    for (int i = 0; i < K; ++i) {
      float row = (float)__ldg(s + (size_t)i * iw) * (wy[i] * wx[0]);
#pragma unroll
      for (int j = 1; j < K; ++j) row = row + (float)__ldg(s + (size_t)i * iw + j) * (wy[i] * wx[j]);
      total = total + row;
    }
  } else {
#pragma unroll
    for (int i = 0; i < K; ++i) {
      const int yy = y0 + i;
      if ((unsigned)yy >= (unsigned)ih) continue;
#pragma unroll
      for (int j = 0; j < K; ++j) {
        const int xx = x0 + j;
        if ((unsigned)xx < (unsigned)iw)
          total = total + (float)__ldg(img + (size_t)yy * iw + xx) * (wy[i] * wx[j]);
      }
    }
  }
  return saturate_out<T>(total);
}

template <typename T, int K>
__device__ __forceinline__ T sample(const T* img, int ih, int iw, int sx, int sy, int fx, int fy,
                                    const Tables* tabs) {
  if constexpr (K == 1) {
    if ((unsigned)sx < (unsigned)iw && (unsigned)sy < (unsigned)ih)
      return __ldg(img + (size_t)sy * iw + sx);
    return (T)0;
  } else if constexpr (sizeof(T) == 1) {
    return sample_u8<K>(img, ih, iw, sx, sy, fx, fy, tabs);
  } else if constexpr (sizeof(T) == 4 && !std::is_floating_point<T>::value) {
    return (T)0;  // 4-byte integers are nearest-neighbour only (checked on the host)
  } else {
    return sample_f<T, K>(img, ih, iw, sx, sy, fx, fy, tabs);
  }
}

constexpr int kBX = 32, kBY = 8;

template <typename T, int K>
__global__ void __launch_bounds__(kBX* kBY) remap_kernel(const Params p) {
  const int x = blockIdx.x * kBX + threadIdx.x, y = blockIdx.y * kBY + threadIdx.y;
  const int z = blockIdx.z;
  if (x >= p.ow || y >= p.oh) return;
  if (p.skip && p.skip[z]) return;  // the output is zero-filled beforehand
  const int iy = __ldg(p.iy + y), ix = __ldg(p.ix + x);
  const double ty = __ldg(p.ty + y), tx = __ldg(p.tx + x);
  const size_t plane = (size_t)p.my * p.mx;
  const float cx = (float)densify(p.abs_map + (size_t)z * plane, p.mx, iy, ix, ty, tx, p.map_f64);
  const float cy = (float)densify(p.abs_map + ((size_t)p.nz + z) * plane, p.mx, iy, ix, ty, tx,
                                  p.map_f64);
  int sx, sy, fx = 0, fy = 0;
  if (K == 1) {
    sx = sat16(cv_round(cx));
    sy = sat16(cv_round(cy));
  } else {
    const int qx = cv_round(cx * 32.f), qy = cv_round(cy * 32.f);
    sx = sat16(qx >> 5);
    sy = sat16(qy >> 5);
    fx = qx & 31;
    fy = qy & 31;
  }
  const T* img = static_cast<const T*>(p.image);
  T* out = static_cast<T*>(p.out);
  for (int c = 0; c < p.n; ++c) {
    const T* sec = img + ((size_t)c * p.nz + z) * p.ih * p.iw;
    out[(((size_t)c * p.nz + z) * p.oh + y) * p.ow + x] =
        sample<T, K>(sec, p.ih, p.iw, sx, sy, fx, fy, p.tabs);
  }
}

template <typename T>
static void launch_k(sofima_ctx* ctx, const Params& p, int ksize) {
  const dim3 grid((unsigned)ceil_div(p.ow, kBX), (unsigned)ceil_div(p.oh, kBY), (unsigned)p.nz);
  const dim3 block(kBX, kBY);
  switch (ksize) {
    case 1: remap_kernel<T, 1><<<grid, block, 0, ctx->stream>>>(p); break;
    case 2: remap_kernel<T, 2><<<grid, block, 0, ctx->stream>>>(p); break;
    case 4: remap_kernel<T, 4><<<grid, block, 0, ctx->stream>>>(p); break;
    default: remap_kernel<T, 8><<<grid, block, 0, ctx->stream>>>(p); break;
  }
}

}  // namespace warpcv
}  // namespace sofima

extern "C" int sofima_warp_subvolume(sofima_ctx* ctx, const void* image, int img_dtype,
                                     const int64_t* image_shape, const double* abs_map,
                                     int map_is_f64, const double* grid_y, const double* grid_x,
                                     int64_t my, int64_t mx, const uint8_t* skip,
                                     int interpolation, void* out, int64_t oh, int64_t ow) {
  using namespace sofima;
  if (!ctx) return fail(nullptr, SOFIMA_EINVAL, "ctx is NULL");
  if (!image_shape) return fail(ctx, SOFIMA_EINVAL, "NULL argument");
  if (interpolation < 0 || interpolation > 3)
    return fail(ctx, SOFIMA_EINVAL, "interpolation must be 0 (nearest), 1 (linear), 2 (cubic) "
                "or 3 (lanczos), got %d", interpolation);
  const int64_t n = image_shape[0], nz = image_shape[1], ih = image_shape[2], iw = image_shape[3];
  if (n < 0 || nz < 0 || ih < 1 || iw < 1 || oh < 0 || ow < 0 || ih > INT32_MAX ||
      iw > INT32_MAX || oh > INT32_MAX || ow > INT32_MAX || nz > 65535 || n > INT32_MAX)
    return fail(ctx, SOFIMA_EINVAL, "extent out of range");
  if (my < 2 || mx < 2 || my > INT32_MAX || mx > INT32_MAX)
    return fail(ctx, SOFIMA_EINVAL, "the coordinate map needs at least 2 nodes per axis "
                "(got %lld x %lld)", (long long)my, (long long)mx);
  size_t esize = 0;
  switch (img_dtype) {
    case SOFIMA_U8: esize = 1; break;
    case SOFIMA_U16: case SOFIMA_I16: esize = 2; break;
    case SOFIMA_F32: case SOFIMA_U32: esize = 4; break;
    default: return fail(ctx, SOFIMA_EINVAL, "unsupported image dtype %d", img_dtype);
  }
  if (img_dtype == SOFIMA_U32 && interpolation != 0)
    return fail(ctx, SOFIMA_EUNSUPPORTED, "32-bit integer sections (label ids) are sampled "
                "with nearest-neighbour interpolation only");
  if (n * nz * oh * ow == 0) return SOFIMA_OK;
  if (!image || !abs_map || !grid_y || !grid_x || !out)
    return fail(ctx, SOFIMA_EINVAL, "NULL array argument");
  DeviceGuard guard(ctx->device);
  void* dtabs = nullptr;
  const bool fresh = ctx->scratch.find("warpcv.tables") == ctx->scratch.end();
  if (int rc = scratch(ctx, "warpcv.tables", sizeof(warpcv::Tables), &dtabs)) return rc;
  if (fresh) {
    const cudaError_t e = cudaMemcpyAsync(dtabs, &warpcv::host_tables(), sizeof(warpcv::Tables),
                                          cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) {
      ctx->scratch.erase("warpcv.tables");
      cudaFree(dtabs);
      return fail(ctx, SOFIMA_ECUDA, "table upload: %s", cudaGetErrorString(e));
    }
  }
  warpcv::Params p;
  p.image = image; p.out = out; p.abs_map = abs_map;
  void* axes = nullptr;
  const size_t n_ax = (size_t)oh + (size_t)ow;
  if (int rc = scratch(ctx, "warpcv.axes", n_ax * (sizeof(double) + sizeof(int)), &axes)) return rc;
  double* dist = static_cast<double*>(axes);
  int* idx = reinterpret_cast<int*>(dist + n_ax);
  p.ty = dist; p.tx = dist + oh; p.iy = idx; p.ix = idx + oh;
  p.skip = skip; p.tabs = static_cast<const warpcv::Tables*>(dtabs);
  p.n = (int)n; p.nz = (int)nz; p.ih = (int)ih; p.iw = (int)iw; p.oh = (int)oh; p.ow = (int)ow;
  p.my = (int)my; p.mx = (int)mx; p.map_f64 = map_is_f64;
  LaunchTimer timer(ctx, "warp_subvolume");
  if (skip) {
    const cudaError_t e = cudaMemsetAsync(out, 0, (size_t)(n * nz * oh * ow) * esize, ctx->stream);
    if (e != cudaSuccess) return fail(ctx, SOFIMA_ECUDA, "memset: %s", cudaGetErrorString(e));
  }
  warpcv::axis_table_kernel<<<(unsigned)ceil_div<int64_t>(oh, 256), 256, 0, ctx->stream>>>(
      grid_y, (int)my, (int)oh, idx, dist);
  warpcv::axis_table_kernel<<<(unsigned)ceil_div<int64_t>(ow, 256), 256, 0, ctx->stream>>>(
      grid_x, (int)mx, (int)ow, idx + oh, dist + oh);
  static const int ks[4] = {1, 2, 4, 8};
  const int k = ks[interpolation];
  switch (img_dtype) {
    case SOFIMA_U8: warpcv::launch_k<uint8_t>(ctx, p, k); break;
    case SOFIMA_U16: warpcv::launch_k<uint16_t>(ctx, p, k); break;
    case SOFIMA_I16: warpcv::launch_k<int16_t>(ctx, p, k); break;
    case SOFIMA_F32: warpcv::launch_k<float>(ctx, p, k); break;
    default: warpcv::launch_k<uint32_t>(ctx, p, 1); break;
  }
  SOFIMA_CHECK_LAUNCH(ctx);
  return SOFIMA_OK;
}
