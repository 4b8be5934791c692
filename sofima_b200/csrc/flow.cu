// Patch-flow estimator on B200: batched (masked) cross-correlation + peak statistics.
//
// Replaces the device side of flow_field.batched_xcorr_peaks (reference
// flow_field.py:385-441): _batched_xcorr (:278-371) -> masked_xcorr (:36-156) ->
// _batched_peaks (:205-275) / _peak_stats (:178-202).
//
// Algorithm: the reference's own -- zero-padded real FFT correlation in fp32 --
// written as shared-memory Stockham FFTs (any 2^a 3^b 5^c length, which is what
// scipy's next_fast_len returns), three fused stages per batch:
//   rows_fwd : gather patch rows from the images (uint8/float, clamped starts,
//              mean subtraction, mask zeroing, flip of the 'post' patch), two real
//              rows per complex FFT, half spectrum out               -> T
//   cols     : per spectral column: forward FFT of every input, spectral products,
//              inverse FFT                                           -> U
//   rows_inv : two Hermitian rows per complex inverse FFT, crop, scale -> image
// then the Padfield normalisation (masked path) and the peak search.
// The spectra never leave L2: the batch is processed in sub-batches whose scratch
// footprint is sized well below the 126 MB L2.
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "common.cuh"

namespace sofima {
namespace flow {

constexpr int kThreads = 256;
constexpr int kMaxStages = 12;
constexpr int kMaxFftLen = 8192;

struct FftPlan {
  int L;
  int nstages;
  int radix[kMaxStages];
  int s[kMaxStages];   // stride of the stage (product of earlier radices)
  int m[kMaxStages];   // n / r
  int bf[kMaxStages];  // butterflies per transform = L / r
  unsigned mg_bf[kMaxStages], mg_s[kMaxStages];  // magic multipliers (0: divisor 1)
  const float2* tw;    // exp(-2 pi i k / L), k < L (device)
};

__host__ inline unsigned magic(unsigned d) {
  return d == 1 ? 0u : (unsigned)((0x100000000ull + d - 1) / d);
}
__device__ __forceinline__ unsigned fdiv(unsigned n, unsigned mg) {
  return mg == 0 ? n : __umulhi(n, mg);
}

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// multiply by -i (forward) or +i (inverse)
template <bool INV>
__device__ __forceinline__ float2 rot(float2 a) {
  return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}
template <bool INV>
__device__ __forceinline__ float2 twid(const float2* tw, int idx) {
  float2 w = tw[idx];
  if (INV) w.y = -w.y;
  return w;
}

// One Stockham pass over `nfft` transforms stored as [nfft][pitch] in shared memory.
template <bool INV>
__device__ void fft_pass(const float2* __restrict__ src, float2* __restrict__ dst, int nfft,
                         const FftPlan& P, int st, const float2* tw, int pitch) {
  const int r = P.radix[st], s = P.s[st], m = P.m[st], bf = P.bf[st], L = pitch;
  const unsigned mgb = P.mg_bf[st], mgs = P.mg_s[st];
  const int total = nfft * bf;
  const int sm = s * m;
  for (int t = threadIdx.x; t < total; t += blockDim.x) {
    const int f = (int)fdiv((unsigned)t, mgb);
    const int u = t - f * bf;
    const int p = (int)fdiv((unsigned)u, mgs);
    const int q = u - p * s;
    const float2* in = src + f * L + q + s * p;
    float2* out = dst + f * L + q + s * r * p;
    const int tb = p * s;
    if (r == 4) {
      const float2 a0 = in[0], a1 = in[sm], a2 = in[2 * sm], a3 = in[3 * sm];
      const float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3);
      const float2 t3 = rot<INV>(csub(a1, a3));
      out[0] = cadd(t0, t2);
      out[s] = cmul(cadd(t1, t3), twid<INV>(tw, tb));
      out[2 * s] = cmul(csub(t0, t2), twid<INV>(tw, 2 * tb));
      out[3 * s] = cmul(csub(t1, t3), twid<INV>(tw, 3 * tb));
    } else if (r == 2) {
      const float2 a0 = in[0], a1 = in[sm];
      out[0] = cadd(a0, a1);
      out[s] = cmul(csub(a0, a1), twid<INV>(tw, tb));
    } else if (r == 5) {
      const float c1 = 0.30901699437494745f, c2 = -0.80901699437494745f;
      const float s1 = 0.95105651629515353f, s2 = 0.58778525229247314f;
      const float2 a0 = in[0], a1 = in[sm], a2 = in[2 * sm], a3 = in[3 * sm], a4 = in[4 * sm];
      const float2 t1 = cadd(a1, a4), t2 = cadd(a2, a3), t3 = csub(a1, a4), t4 = csub(a2, a3);
      const float2 m1 = make_float2(a0.x + c1 * t1.x + c2 * t2.x, a0.y + c1 * t1.y + c2 * t2.y);
      const float2 m2 = make_float2(a0.x + c2 * t1.x + c1 * t2.x, a0.y + c2 * t1.y + c1 * t2.y);
      const float2 n1 = rot<INV>(make_float2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y));
      const float2 n2 = rot<INV>(make_float2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y));
      out[0] = make_float2(a0.x + t1.x + t2.x, a0.y + t1.y + t2.y);
      out[s] = cmul(cadd(m1, n1), twid<INV>(tw, tb));
      out[2 * s] = cmul(cadd(m2, n2), twid<INV>(tw, 2 * tb));
      out[3 * s] = cmul(csub(m2, n2), twid<INV>(tw, 3 * tb));
      out[4 * s] = cmul(csub(m1, n1), twid<INV>(tw, 4 * tb));
    } else {  // r == 3
      const float h = 0.86602540378443865f;
      const float2 a0 = in[0], a1 = in[sm], a2 = in[2 * sm];
      const float2 t1 = cadd(a1, a2);
      const float2 t2 = make_float2(a0.x - 0.5f * t1.x, a0.y - 0.5f * t1.y);
      const float2 d = csub(a1, a2);
      const float2 t3 = rot<INV>(make_float2(h * d.x, h * d.y));
      out[0] = cadd(a0, t1);
      out[s] = cmul(cadd(t2, t3), twid<INV>(tw, tb));
      out[2 * s] = cmul(csub(t2, t3), twid<INV>(tw, 2 * tb));
    }
  }
}

// Full transform of `nfft` lines; returns the buffer that holds the result.
// `pitch`: float2 between consecutive lines (0 = P.L; an odd pitch keeps column-wise accesses
// to the lines conflict-free).
template <bool INV>
__device__ float2* block_fft(float2* a, float2* b, int nfft, const FftPlan& P, const float2* tw,
                             int pitch = 0) {
  if (pitch == 0) pitch = P.L;
  float2 *src = a, *dst = b;
  for (int st = 0; st < P.nstages; ++st) {
    fft_pass<INV>(src, dst, nfft, P, st, tw, pitch);
    __syncthreads();
    float2* t = src; src = dst; dst = t;
  }
  return src;
}

__device__ __forceinline__ void load_twiddles(float2* tw_s, const FftPlan& P) {
  for (int i = threadIdx.x; i < P.L; i += blockDim.x) tw_s[i] = __ldg(&P.tw[i]);
}

// ---------------------------------------------------------------------------------
// Problem description shared by the kernels (2-d).
// ---------------------------------------------------------------------------------
struct Image {
  const void* data;
  const uint8_t* mask;  // may be null
  int h, w;             // image extent
  int mh, mw;           // mask extent
  int ph, pw;           // patch extent
};

struct MeanPartial;

struct Slot {
  int src;    // 0 = pre, 1 = post (flipped)
  int xform;  // 0: (v - mean) * valid, 1: valid indicator, 2: ((v - mean) * valid)^2
};

struct Problem {
  Image img[2];
  int dtype;               // SOFIMA_U8 | SOFIMA_F32
  const int32_t* starts[2];  // [B][2] (y, x), device
  const struct MeanPartial* parts;  // [B][2][kMeanGroups] partial patch sums
  int has_mean;            // constant mean instead of the per-patch mean
  float mean;
  int nslots;
  Slot slot[6];
  int PY;                  // max patch height (row capacity of T)
  int sy, sx;              // correlation image extent (pre + post - 1)
  int nkx;                 // Lx / 2 + 1
  long long b0;            // first pair of this sub-batch
  int nb;                  // pairs in this sub-batch
};

__device__ __forceinline__ int clamp_start(int st, int size, int extent) {
  // jax.lax.dynamic_slice clamps the start so that the slice stays in bounds.
  const int hi = extent - size;
  return st < 0 ? 0 : (st > hi ? hi : st);
}

__device__ __forceinline__ float load_px(const void* data, int dtype, long long i) {
  return dtype == SOFIMA_U8 ? (float)static_cast<const uint8_t*>(data)[i]
                            : static_cast<const float*>(data)[i];
}

// Per-patch mean (flow_field.py:340-353): masked mean over the unmasked pixels.
// grid = (pair, image, row group): every block sums kMeanRows-th of the rows and
// writes one (sum, count) partial; the consumers add the kMeanGroups partials in a
// fixed order and divide (patch_mean()).  Sums are exact for uint8 images.
constexpr int kMeanGroups = 8;

struct MeanPartial {
  double sum;
  int count;
  int pad;
};

__global__ void __launch_bounds__(kThreads)
patch_mean_kernel(Problem P, MeanPartial* parts) {
  const int which = blockIdx.y;
  const long long b = P.b0 + blockIdx.x;
  const Image& I = P.img[which];
  const int y0 = clamp_start(P.starts[which][b * 2 + 0], I.ph, I.h);
  const int x0 = clamp_start(P.starts[which][b * 2 + 1], I.pw, I.w);
  int my0 = 0, mx0 = 0;
  if (I.mask) {
    my0 = clamp_start(P.starts[which][b * 2 + 0], I.ph, I.mh);
    mx0 = clamp_start(P.starts[which][b * 2 + 1], I.pw, I.mw);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rows_per = ceil_div(I.ph, kMeanGroups);
  const int ya = blockIdx.z * rows_per, yb = min(I.ph, ya + rows_per);
  double sum = 0.0;
  unsigned int isum = 0;
  int cnt = 0;
  for (int y = ya + warp; y < yb; y += kThreads / 32) {
    const long long row = (long long)(y0 + y) * I.w + x0;
    const uint8_t* mrow = I.mask ? I.mask + (long long)(my0 + y) * I.mw + mx0 : nullptr;
    if (P.dtype == SOFIMA_U8) {
      const uint8_t* src = static_cast<const uint8_t*>(I.data) + row;
#pragma unroll 4
      for (int x = lane; x < I.pw; x += 32) {
        const bool valid = mrow ? (mrow[x] == 0) : true;
        isum += valid ? (unsigned)src[x] : 0u;
        cnt += valid;
      }
    } else {
      const float* src = static_cast<const float*>(I.data) + row;
#pragma unroll 4
      for (int x = lane; x < I.pw; x += 32) {
        const bool valid = mrow ? (mrow[x] == 0) : true;
        sum += valid ? (double)src[x] : 0.0;
        cnt += valid;
      }
    }
  }
  if (P.dtype == SOFIMA_U8) sum = (double)isum;
  __shared__ double rs[kThreads / 32];
  __shared__ int rc[kThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if (lane == 0) { rs[warp] = sum; rc[warp] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0; int c = 0;
    for (int w = 0; w < kThreads / 32; ++w) { t += rs[w]; c += rc[w]; }
    MeanPartial mp;
    mp.sum = t; mp.count = c; mp.pad = 0;
    parts[(b * 2 + which) * kMeanGroups + blockIdx.z] = mp;
  }
}

// uint8 images without masks (every BASELINE flow configuration): the patch rows are
// read as aligned 32-bit words and summed four pixels at a time with dp4a; bytes
// outside [x0, x0 + pw) are masked off.  Same grid, same exact integer partial sums.
__global__ void __launch_bounds__(kThreads)
patch_sum_u8_kernel(Problem P, MeanPartial* parts) {
  // grid = (pair, image); warp g sums row group g (kMeanGroups == warps per block).
  static_assert(kMeanGroups == kThreads / 32, "one warp per row group");
  const int which = blockIdx.y;
  const long long b = P.b0 + blockIdx.x;
  const Image& I = P.img[which];
  const int y0 = clamp_start(P.starts[which][b * 2 + 0], I.ph, I.h);
  const int x0 = clamp_start(P.starts[which][b * 2 + 1], I.pw, I.w);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rows_per = ceil_div(I.ph, kMeanGroups);
  const int ya = warp * rows_per, yb = min(I.ph, ya + rows_per);
  const uint8_t* img = static_cast<const uint8_t*>(I.data);
  const uint8_t* img_end = img + (long long)I.h * I.w;
  unsigned int isum = 0;
#pragma unroll 4
  for (int y = ya; y < yb; ++y) {
    const uint8_t* src = img + (long long)(y0 + y) * I.w + x0;
    const int mis = (int)(reinterpret_cast<uintptr_t>(src) & 3);
    const uint8_t* wbase = src - mis;
    const int nwords = (mis + I.pw + 3) >> 2;
    for (int i = lane; i < nwords; i += 32) {
      const uint8_t* wp = wbase + 4 * i;
      const int first = max(mis - 4 * i, 0);            // first valid byte of the word
      const int last = min(mis + I.pw - 4 * i, 4);      // one past the last valid byte
      if (wp >= img && wp + 4 <= img_end) {
        unsigned int wv = __ldg(reinterpret_cast<const unsigned int*>(wp));
        unsigned int m = 0xffffffffu;
        if (first > 0) m &= 0xffffffffu << (8 * first);
        if (last < 4) m &= (1u << (8 * last)) - 1u;
        isum = __dp4a(wv & m, 0x01010101u, isum);
      } else {  // the word straddles the ends of the image buffer
        for (int j = first; j < last; ++j) isum += wp[j];
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) isum += __shfl_xor_sync(0xffffffffu, isum, o);
  if (lane == 0) {
    MeanPartial mp;
    mp.sum = (double)isum;
    mp.count = (yb > ya ? yb - ya : 0) * I.pw;
    mp.pad = 0;
    parts[(b * 2 + which) * kMeanGroups + warp] = mp;
  }
}

// fp32 sum / fp32 count, as jnp.mean / jnp.nanmean (0 / 0 -> NaN).
__device__ __forceinline__ float patch_mean(const Problem& P, long long b, int which) {
  if (P.has_mean) return P.mean;
  const MeanPartial* mp = P.parts + (b * 2 + which) * kMeanGroups;
  double t = 0.0;
  int c = 0;
#pragma unroll
  for (int j = 0; j < kMeanGroups; ++j) { t += mp[j].sum; c += mp[j].count; }
  return __fdiv_rn((float)t, (float)c);
}

// ---------------------------------------------------------------------------------
// Stage 1: forward row FFTs.  grid = (row-pair groups, slot, pair).
// T layout: [slot][pair][PY][nkx] float2.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
rows_fwd_kernel(Problem P, FftPlan F, int R, float2* __restrict__ T) {
  extern __shared__ float2 smem[];
  float2* buf0 = smem;
  float2* buf1 = smem + (size_t)R * F.L;
  float2* tw_s = buf1 + (size_t)R * F.L;
  const Slot sl = P.slot[blockIdx.y];
  const Image& I = P.img[sl.src];
  const long long b = P.b0 + blockIdx.z;
  const int rp0 = blockIdx.x * R;  // first row pair
  const int nrows = I.ph;
  if (2 * rp0 >= nrows) return;
  load_twiddles(tw_s, F);

  const int y0 = clamp_start(P.starts[sl.src][b * 2 + 0], I.ph, I.h);
  const int x0 = clamp_start(P.starts[sl.src][b * 2 + 1], I.pw, I.w);
  int my0 = 0, mx0 = 0;
  if (I.mask) {
    my0 = clamp_start(P.starts[sl.src][b * 2 + 0], I.ph, I.mh);
    mx0 = clamp_start(P.starts[sl.src][b * 2 + 1], I.pw, I.mw);
  }
  const float mean = patch_mean(P, b, sl.src);
  const bool flip = sl.src == 1;  // curr[::-1, ::-1], flow_field.py:78-79

  auto sample = [&](int y, int x) -> float {
    if (y >= nrows) return 0.f;
    const int yy = flip ? I.ph - 1 - y : y, xx = flip ? I.pw - 1 - x : x;
    bool valid = true;
    if (I.mask) valid = I.mask[(long long)(my0 + yy) * I.mw + mx0 + xx] == 0;
    if (sl.xform == 1) return valid ? 1.f : 0.f;
    float v = load_px(I.data, P.dtype, (long long)(y0 + yy) * I.w + x0 + xx) - mean;
    v = valid ? v : 0.f;  // where(mask, 0, patch - mean), flow_field.py:73-76
    return sl.xform == 2 ? v * v : v;
  };

  for (int i = threadIdx.x; i < R * F.L; i += kThreads) {
    const int f = i / F.L, x = i - f * F.L;
    float2 z = make_float2(0.f, 0.f);
    const int y = 2 * (rp0 + f);
    if (x < I.pw && y < nrows) z = make_float2(sample(y, x), sample(y + 1, x));
    buf0[i] = z;
  }
  __syncthreads();
  const float2* Z = block_fft<false>(buf0, buf1, R, F, tw_s);

  float2* Tb = T + ((size_t)blockIdx.y * P.nb + blockIdx.z) * P.PY * P.nkx;
  for (int i = threadIdx.x; i < R * P.nkx; i += kThreads) {
    const int f = i / P.nkx, k = i - f * P.nkx;
    const int y = 2 * (rp0 + f);
    if (y >= nrows) continue;
    const float2 a = Z[f * F.L + k];
    const float2 c = Z[f * F.L + (k == 0 ? 0 : F.L - k)];
    // X_even = (Z[k] + conj Z[L-k]) / 2 ; X_odd = (Z[k] - conj Z[L-k]) / (2i)
    Tb[(size_t)y * P.nkx + k] = make_float2(0.5f * (a.x + c.x), 0.5f * (a.y - c.y));
    if (y + 1 < nrows)
      Tb[(size_t)(y + 1) * P.nkx + k] = make_float2(0.5f * (a.y + c.y), 0.5f * (c.x - a.x));
  }
}

// ---------------------------------------------------------------------------------
// Stage 2: column FFTs, spectral products, inverse column FFTs.
// grid = (column groups, pair).  U layout: [out][pair][sy][nkx] float2.
// ---------------------------------------------------------------------------------
struct Products {
  int nout;
  int a[6], b[6];  // out[i] = spectrum[a[i]] * spectrum[b[i]]
};

__global__ void __launch_bounds__(kThreads)
cols_kernel(Problem P, FftPlan F, int C, Products pr, const float2* __restrict__ T,
            float2* __restrict__ U) {
  extern __shared__ float2 smem[];
  const int L = F.L;
  float2* w0 = smem;
  float2* w1 = w0 + (size_t)C * L;
  float2* spec = w1 + (size_t)C * L;               // [nslots][C][L]
  float2* tw_s = spec + (size_t)P.nslots * C * L;
  load_twiddles(tw_s, F);
  const int k0 = blockIdx.x * C;
  const int nc = min(C, P.nkx - k0);

  for (int sl = 0; sl < P.nslots; ++sl) {
    const int rows = P.img[P.slot[sl].src].ph;
    const float2* Tb = T + ((size_t)sl * P.nb + blockIdx.y) * P.PY * P.nkx;
    for (int i = threadIdx.x; i < C * L; i += kThreads) {
      const int y = i / C, c = i - y * C;  // c fastest: contiguous global reads
      float2 v = make_float2(0.f, 0.f);
      if (c < nc && y < rows) v = Tb[(size_t)y * P.nkx + k0 + c];
      w0[c * L + y] = v;
    }
    __syncthreads();
    const float2* res = block_fft<false>(w0, w1, C, F, tw_s);
    float2* dst = spec + (size_t)sl * C * L;
    for (int i = threadIdx.x; i < C * L; i += kThreads) dst[i] = res[i];
    __syncthreads();
  }
  for (int o = 0; o < pr.nout; ++o) {
    const float2* A = spec + (size_t)pr.a[o] * C * L;
    const float2* B = spec + (size_t)pr.b[o] * C * L;
    for (int i = threadIdx.x; i < C * L; i += kThreads) w0[i] = cmul(A[i], B[i]);
    __syncthreads();
    const float2* res = block_fft<true>(w0, w1, C, F, tw_s);
    float2* Ub = U + ((size_t)o * P.nb + blockIdx.y) * P.sy * P.nkx;
    for (int i = threadIdx.x; i < C * P.sy; i += kThreads) {
      const int y = i / C, c = i - y * C;
      if (c < nc) Ub[(size_t)y * P.nkx + k0 + c] = res[c * L + y];
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------
// Stage 2 for LONG columns (whole-strip correlations of the coarse tile offsets,
// stitch_rigid.py:39-67: a 4096 x 300 strip gives 8192-point columns).  One line no
// longer fits next to the spectra of all slots in shared memory, so the forward
// spectra go through global memory: S layout [slot][pair][nkx][L] (lines contiguous).
//   cols_fwd_long: grid (column, slot, pair);  cols_inv_long: grid (column, out, pair).
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
cols_fwd_long_kernel(Problem P, FftPlan F, const float2* __restrict__ T,
                     float2* __restrict__ S) {
  extern __shared__ float2 smem[];
  const int L = F.L;
  float2* w0 = smem;
  float2* w1 = w0 + L;
  float2* tw_s = w1 + L;
  load_twiddles(tw_s, F);
  const int col = blockIdx.x, sl = blockIdx.y;
  const int rows = P.img[P.slot[sl].src].ph;
  const float2* Tb = T + ((size_t)sl * P.nb + blockIdx.z) * P.PY * P.nkx + col;
  for (int y = threadIdx.x; y < L; y += kThreads)
    w0[y] = y < rows ? Tb[(size_t)y * P.nkx] : make_float2(0.f, 0.f);
  __syncthreads();
  const float2* res = block_fft<false>(w0, w1, 1, F, tw_s);
  float2* Sb = S + (((size_t)sl * P.nb + blockIdx.z) * P.nkx + col) * L;
  for (int i = threadIdx.x; i < L; i += kThreads) Sb[i] = res[i];
}

__global__ void __launch_bounds__(kThreads)
cols_inv_long_kernel(Problem P, FftPlan F, Products pr, const float2* __restrict__ S,
                     float2* __restrict__ U) {
  extern __shared__ float2 smem[];
  const int L = F.L;
  float2* w0 = smem;
  float2* w1 = w0 + L;
  float2* tw_s = w1 + L;
  load_twiddles(tw_s, F);
  const int col = blockIdx.x, o = blockIdx.y;
  const float2* A = S + (((size_t)pr.a[o] * P.nb + blockIdx.z) * P.nkx + col) * L;
  const float2* B = S + (((size_t)pr.b[o] * P.nb + blockIdx.z) * P.nkx + col) * L;
  for (int i = threadIdx.x; i < L; i += kThreads) w0[i] = cmul(A[i], B[i]);
  __syncthreads();
  const float2* res = block_fft<true>(w0, w1, 1, F, tw_s);
  float2* Ub = U + ((size_t)o * P.nb + blockIdx.z) * P.sy * P.nkx + col;
  for (int y = threadIdx.x; y < P.sy; y += kThreads) Ub[(size_t)y * P.nkx] = res[y];
}

// ---------------------------------------------------------------------------------
// Stage 3: inverse row FFTs (two Hermitian rows per complex transform), crop, scale.
// grid = (row-pair groups, out, pair).  dst[out] + pair * sy * sx.
// ---------------------------------------------------------------------------------
struct Outputs {
  float* dst[6];
};

__global__ void __launch_bounds__(kThreads)
rows_inv_kernel(Problem P, FftPlan F, int R, const float2* __restrict__ U, Outputs outs,
                float scale) {
  extern __shared__ float2 smem[];
  float2* buf0 = smem;
  float2* buf1 = smem + (size_t)R * F.L;
  float2* tw_s = buf1 + (size_t)R * F.L;
  const int L = F.L;
  const int rp0 = blockIdx.x * R;
  if (2 * rp0 >= P.sy) return;
  load_twiddles(tw_s, F);
  const float2* Ub = U + ((size_t)blockIdx.y * P.nb + blockIdx.z) * P.sy * P.nkx;
  const bool even = (L & 1) == 0;
  for (int i = threadIdx.x; i < R * L; i += kThreads) {
    const int f = i / L, k = i - f * L;
    const int y = 2 * (rp0 + f);
    float2 z = make_float2(0.f, 0.f);
    if (y < P.sy) {
      const int kk = (k < P.nkx) ? k : L - k;
      float2 u0 = Ub[(size_t)y * P.nkx + kk];
      float2 u1 = (y + 1 < P.sy) ? Ub[(size_t)(y + 1) * P.nkx + kk] : make_float2(0.f, 0.f);
      // c2r ignores the imaginary part of the DC and Nyquist bins.
      if (kk == 0 || (even && kk == L / 2)) { u0.y = 0.f; u1.y = 0.f; }
      if (k >= P.nkx) { u0.y = -u0.y; u1.y = -u1.y; }
      z = make_float2(u0.x - u1.y, u0.y + u1.x);  // u0 + i u1
    }
    buf0[i] = z;
  }
  __syncthreads();
  const float2* Z = block_fft<true>(buf0, buf1, R, F, tw_s);
  float* out = outs.dst[blockIdx.y] + (size_t)(P.b0 + blockIdx.z) * P.sy * P.sx;
  for (int i = threadIdx.x; i < R * P.sx; i += kThreads) {
    const int f = i / P.sx, x = i - f * P.sx;
    const int y = 2 * (rp0 + f);
    if (y >= P.sy) continue;
    const float2 z = Z[f * L + x];
    out[(size_t)y * P.sx + x] = z.x * scale;
    if (y + 1 < P.sy) out[(size_t)(y + 1) * P.sx + x] = z.y * scale;
  }
}

// ---------------------------------------------------------------------------------
// Padfield normalisation (masked path), flow_field.py:113-155.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ float atomic_max_nonneg(float* addr, float v) {
  // valid for v >= 0 (IEEE order == integer order); NaN is handled by the caller.
  return __int_as_float(atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v)));
}

// in: six correlation images per pair (xc, ov, mcp, mcc, psq, csq)
// out: numerator (in place of xc), denom (in place of ov slot 1 -> stored in mcp),
//      overlap (in place of ov).  maxima[0] = max |denom|, maxima[1] = max overlap.
__global__ void __launch_bounds__(kThreads)
padfield_terms_kernel(float* xc, float* ov, float* mcp, const float* mcc, const float* psq,
                      const float* csq, long long n, float* maxima) {
  const float eps = 1.1920929e-07f;
  float mden = 0.f, mov = 0.f;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n;
       i += (long long)gridDim.x * kThreads) {
    const float o = fmaxf(rintf(ov[i]), eps);  // round-half-even, fmax(., eps)
    const float oi = 1.0f / o;
    const float p = mcp[i], c = mcc[i];
    const float num = xc[i] - p * c * oi;
    const float pd = fmaxf(psq[i] - p * p * oi, 0.f);
    const float cd = fmaxf(csq[i] - c * c * oi, 0.f);
    const float den = sqrtf(pd * cd);
    xc[i] = num;
    mcp[i] = den;
    ov[i] = o;
    mden = fmaxf(mden, fabsf(den));
    mov = fmaxf(mov, o);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mden = fmaxf(mden, __shfl_xor_sync(0xffffffffu, mden, o));
    mov = fmaxf(mov, __shfl_xor_sync(0xffffffffu, mov, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomic_max_nonneg(&maxima[0], mden);
    atomic_max_nonneg(&maxima[1], mov);
  }
}

__global__ void __launch_bounds__(kThreads)
padfield_normalise_kernel(float* xc, const float* ov, const float* den, long long n,
                          const float* maxima) {
  const float eps = 1.1920929e-07f;
  const float tol = 1e3f * eps * maxima[0];  // batch-global max, flow_field.py:137
  const float px_thr = 0.3f * maxima[1];     // batch-global max, flow_field.py:151
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n;
       i += (long long)gridDim.x * kThreads) {
    const float d = den[i];
    float out = (d > tol) ? xc[i] / d : 0.f;
    out = fminf(fmaxf(out, -1.f), 1.f);
    xc[i] = (ov[i] < px_thr) ? 0.f : out;
  }
}

// ---------------------------------------------------------------------------------
// Peaks (flow_field.py:205-275, :178-202).
// ---------------------------------------------------------------------------------
struct PeakParams {
  int ndim;          // 2 or 3
  int sz, sy, sx;    // correlation image extent (sz = 1 in 2-d)
  int md;            // min_distance
  float thr_rel;
  int rz, ry, rx;    // peak_radius
  int cz, cy, cx;    // center_offset
};

// order-preserving key: larger value first, then smaller flat index.
__device__ __forceinline__ unsigned long long peak_key(float v, unsigned idx) {
  unsigned u = __float_as_uint(v);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ((unsigned long long)u << 32) | (unsigned long long)(0xFFFFFFFFu - idx);
}
__device__ __forceinline__ void key_decode(unsigned long long k, float* v, unsigned* idx) {
  unsigned u = (unsigned)(k >> 32);
  u = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
  *v = __uint_as_float(u);
  *idx = 0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull);
}

__device__ unsigned long long block_max_key(unsigned long long k, unsigned long long* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, k, o);
    k = other > k ? other : k;
  }
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = k;
  __syncthreads();
  unsigned long long r = 0;
  for (int w = 0; w < kThreads / 32; ++w) r = sm[w] > r ? sm[w] : r;
  __syncthreads();
  return r;
}

// Pass 1 (generic path; the fast path folds this into rows_inv_fast): global max and
// first argmax per image as an order-preserving key, plus a NaN flag.
__global__ void __launch_bounds__(kThreads)
peak1_kernel(const float* __restrict__ img, PeakParams pp, unsigned long long* keys,
             int* nanflag) {
  __shared__ unsigned long long sm[kThreads / 32];
  __shared__ int nan_flag;
  if (threadIdx.x == 0) nan_flag = 0;
  __syncthreads();
  const long long n = (long long)pp.sz * pp.sy * pp.sx;
  const float* im = img + blockIdx.x * n;
  // gridDim.y blocks share one image (3-d correlation volumes have millions of voxels); the
  // keys are ordered by (value, lowest index), so atomicMax gives the same result as one block.
  // keys / nanflag are zeroed by the caller.
  const long long i0 = n * blockIdx.y / gridDim.y, i1 = n * (blockIdx.y + 1) / gridDim.y;
  unsigned long long best = 0;
  int has_nan = 0;
  for (long long i = i0 + threadIdx.x; i < i1; i += kThreads) {
    const float v = im[i];
    if (v != v) has_nan = 1;
    const unsigned long long k = peak_key(v, (unsigned)i);
    best = k > best ? k : best;
  }
  if (has_nan) nan_flag = 1;
  best = block_max_key(best, sm);
  if (threadIdx.x == 0) {
    if (best != 0) atomicMax(&keys[blockIdx.x], best);
    if (nan_flag) atomicOr(&nanflag[blockIdx.x], 1);
  }
}

// Decodes the keys into (v1, p1) and marks every first peak in the batch bitmap.
__global__ void peak1_decode_kernel(const unsigned long long* keys, const int* nanflag,
                                    long long B, PeakParams pp, float* v1, int* p1,
                                    unsigned* bitmap) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float v = -INFINITY;
  unsigned idx = 0;
  if (keys[b] != 0) key_decode(keys[b], &v, &idx);
  // The global maximum is a peak iff it exceeds threshold_rel * max (it always
  // equals its own zero-padded neighbourhood maximum), i.e. iff it is > 0.
  const bool ok = !nanflag[b] && keys[b] != 0 && (v > pp.thr_rel * v);
  v1[b] = ok ? v : -INFINITY;
  p1[b] = ok ? (int)idx : 0;  // argmax of an all -inf row is 0
  atomicOr(&bitmap[p1[b] >> 5], 1u << (p1[b] & 31));
}

__device__ __forceinline__ bool is_peak(const float* im, const PeakParams& pp, int z, int y,
                                        int x, float v) {
  // img == separable zero-padded max filter of width 2*md+1 (flow_field.py:237-254).
  float m = -INFINITY;
  const int mz = pp.ndim == 3 ? pp.md : 0;
  for (int dz = -mz; dz <= mz; ++dz) {
    const int zz = z + dz;
    for (int dy = -pp.md; dy <= pp.md; ++dy) {
      const int yy = y + dy;
      for (int dx = -pp.md; dx <= pp.md; ++dx) {
        const int xx = x + dx;
        const bool out = zz < 0 || zz >= pp.sz || yy < 0 || yy >= pp.sy || xx < 0 || xx >= pp.sx;
        const float w = out ? 0.f : im[((long long)zz * pp.sy + yy) * pp.sx + xx];
        m = fmaxf(m, w);
      }
    }
  }
  return v == m;
}

// Pass 2: second peak under the batch-coupled exclusion rule + statistics.
// Second-peak candidates of large (3-d) correlation volumes: gridDim.y blocks share one image
// and combine their best (value, lowest index) key with atomicMax -- the same result as the
// scan inside peak2_kernel, which then only reads the key.  keys2 is zeroed by the caller.
__global__ void __launch_bounds__(kThreads)
peak2_scan_kernel(const float* __restrict__ img, PeakParams pp, const float* v1a,
                  const unsigned* __restrict__ bitmap, unsigned long long* keys2) {
  __shared__ unsigned long long sm[kThreads / 32];
  const long long n = (long long)pp.sz * pp.sy * pp.sx;
  const float* im = img + blockIdx.x * n;
  const float v1 = v1a[blockIdx.x];
  if (v1 == -INFINITY) return;
  const float thr = pp.thr_rel * v1;
  const long long i0 = n * blockIdx.y / gridDim.y, i1 = n * (blockIdx.y + 1) / gridDim.y;
  unsigned long long best = 0;
  for (long long i = i0 + threadIdx.x; i < i1; i += kThreads) {
    const float v = __ldg(im + i);
    if (!(v > thr)) continue;
    if ((bitmap[i >> 5] >> (i & 31)) & 1u) continue;  // erased for every row (:263-265)
    const int x = (int)(i % pp.sx), r = (int)(i / pp.sx);
    const int y = r % pp.sy, z = r / pp.sy;
    if (!is_peak(im, pp, z, y, x, v)) continue;
    const unsigned long long k = peak_key(v, (unsigned)i);
    best = k > best ? k : best;
  }
  best = block_max_key(best, sm);
  if (threadIdx.x == 0 && best != 0) atomicMax(&keys2[blockIdx.x], best);
}

__global__ void __launch_bounds__(kThreads)
peak2_kernel(const float* __restrict__ img, PeakParams pp, const float* v1a, const int* p1a,
             const unsigned* __restrict__ bitmap, int ndim_out, float* out,
             const float* __restrict__ bandmax, int band_rows, int nbands,
             const unsigned long long* __restrict__ keys2) {
  __shared__ unsigned long long sm[kThreads / 32];
  __shared__ float smin[kThreads / 32];
  const long long n = (long long)pp.sz * pp.sy * pp.sx;
  const float* im = img + blockIdx.x * n;
  float* o = out + (long long)blockIdx.x * ndim_out;
  const float v1 = v1a[blockIdx.x];
  const int p1 = p1a[blockIdx.x];
  if (v1 == -INFINITY) {  // no peak: the whole row is NaN (flow_field.py:194-196)
    if (threadIdx.x < ndim_out) o[threadIdx.x] = NAN;
    return;
  }
  const float thr = pp.thr_rel * v1;
  unsigned long long best = 0;
  auto consider = [&](int i, float v) {
    if (!(v > thr)) return;
    if ((bitmap[i >> 5] >> (i & 31)) & 1u) return;  // erased for every row (:263-265)
    const int x = i % pp.sx, r = i / pp.sx;
    const int y = r % pp.sy, z = r / pp.sy;
    if (!is_peak(im, pp, z, y, x, v)) return;
    const unsigned long long k = peak_key(v, (unsigned)i);
    best = k > best ? k : best;
  };
  auto scan = [&](int lo, int hi) {  // flat pixel range [lo, hi)
    constexpr int U = 8;  // independent loads in flight per thread
    int i = lo + threadIdx.x;
    for (; i + (U - 1) * kThreads < hi; i += U * kThreads) {
      float v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = __ldg(im + i + u * kThreads);
#pragma unroll
      for (int u = 0; u < U; ++u) consider(i + u * kThreads, v[u]);
    }
    for (; i < hi; i += kThreads) consider(i, __ldg(im + i));
  };
  if (keys2 != nullptr) {
    best = keys2[blockIdx.x];  // found by peak2_scan_kernel
  } else if (bandmax != nullptr) {
    // rows_inv_fast recorded the maximum of every band of `band_rows` rows: bands
    // that cannot hold a pixel above the threshold are never read.
    const float* bmx = bandmax + (long long)blockIdx.x * nbands;
    for (int bnd = 0; bnd < nbands; ++bnd) {
      if (!(__ldg(bmx + bnd) > thr)) continue;
      const int lo = bnd * band_rows * pp.sx;
      const int hi = min((int)n, lo + band_rows * pp.sx);
      scan(lo, hi);
    }
  } else {
    scan(0, (int)n);
  }
  best = block_max_key(best, sm);

  // Sharpness: peak / min over a (2r+1) window with clamped start (:190-192).
  const int px = p1 % pp.sx, pr = p1 / pp.sx;
  const int py = pr % pp.sy, pz = pr / pp.sy;
  const int wz = pp.ndim == 3 ? 2 * pp.rz + 1 : 1, wy = 2 * pp.ry + 1, wx = 2 * pp.rx + 1;
  const int z0 = pp.ndim == 3 ? clamp_start(pz - pp.rz, wz, pp.sz) : 0;
  const int y0 = clamp_start(py - pp.ry, wy, pp.sy), x0 = clamp_start(px - pp.rx, wx, pp.sx);
  float mn = INFINITY;
  for (int i = threadIdx.x; i < wz * wy * wx; i += kThreads) {
    const int xx = x0 + i % wx, rr = i / wx;
    const int yy = y0 + rr % wy, zz = z0 + rr / wy;
    mn = fminf(mn, im[((long long)zz * pp.sy + yy) * pp.sx + xx]);
  }
#pragma unroll
  for (int of = 16; of > 0; of >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, of));
  if ((threadIdx.x & 31) == 0) smin[threadIdx.x >> 5] = mn;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 0; w < kThreads / 32; ++w) mn = fminf(mn, smin[w]);
    float v2;
    if (best != 0) {
      unsigned idx;
      key_decode(best, &v2, &idx);
    } else {
      // argmax of an all -inf row is index 0; the value is read from the
      // un-erased array (flow_field.py:266-268).
      const float v0 = im[0];
      v2 = (v0 > thr && is_peak(im, pp, 0, 0, 0, v0)) ? v0 : -INFINITY;
    }
    int c = 0;
    o[c++] = (float)px - (float)pp.cx;
    o[c++] = (float)py - (float)pp.cy;
    if (pp.ndim == 3) o[c++] = (float)pz - (float)pp.cz;
    o[c++] = v1 / mn;
    o[c++] = (v2 == -INFINITY || v2 == INFINITY) ? 0.0f : v1 / v2;
  }
}

}  // namespace flow
}  // namespace sofima

#include "flow_fast.cuh"

namespace sofima {
namespace flow {

// ---------------------------------------------------------------------------------
// Host side.
// ---------------------------------------------------------------------------------
static int next_fast_len(int n) {
  for (;; ++n) {
    int m = n;
    for (int p : {2, 3, 5}) while (m % p == 0) m /= p;
    if (m == 1) return n;
  }
}

static int make_plan(sofima_ctx* ctx, int L, FftPlan* P) {
  if (L > kMaxFftLen)
    return fail(ctx, SOFIMA_EUNSUPPORTED,
                "FFT length %d exceeds the shared-memory FFT limit %d (whole-strip "
                "correlations are not built yet)", L, kMaxFftLen);
  memset(P, 0, sizeof(*P));
  P->L = L;
  int n = L, s = 1, st = 0;
  auto push = [&](int r) {
    P->radix[st] = r;
    P->s[st] = s;
    P->m[st] = n / r;
    P->bf[st] = L / r;
    P->mg_bf[st] = magic((unsigned)(L / r));
    P->mg_s[st] = magic((unsigned)s);
    n /= r;
    s *= r;
    ++st;
  };
  while (n % 4 == 0) push(4);
  while (n % 2 == 0) push(2);
  while (n % 3 == 0) push(3);
  while (n % 5 == 0) push(5);
  if (n != 1 || st > kMaxStages)
    return fail(ctx, SOFIMA_EINVAL, "FFT length %d is not 5-smooth", L);
  P->nstages = st;
  char name[64];
  snprintf(name, sizeof(name), "flow.tw.%d", L);
  const bool fresh = ctx->scratch.find(name) == ctx->scratch.end();
  void* tw = nullptr;
  int rc = scratch(ctx, name, sizeof(float2) * (size_t)L, &tw);
  if (rc) return rc;
  if (fresh) {
    std::vector<float2> h(L);
    for (int k = 0; k < L; ++k) {
      const double a = -2.0 * M_PI * (double)k / (double)L;
      h[k] = make_float2((float)cos(a), (float)sin(a));
    }
    SOFIMA_CUDA(ctx, cudaMemcpyAsync(tw, h.data(), sizeof(float2) * L, cudaMemcpyHostToDevice,
                                     ctx->stream));
    SOFIMA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  P->tw = static_cast<const float2*>(tw);
  return SOFIMA_OK;
}

// keys_ready: the (max, argmax) keys and NaN flags of the batch were already
// produced by rows_inv_fast into the "flow.keys" / "flow.nan" scratch buffers.
static int run_peaks(sofima_ctx* ctx, const float* images, long long B, const PeakParams& pp,
                     float* out_peaks, bool keys_ready, int band_rows = 0, int nbands = 0) {
  if (B == 0) return SOFIMA_OK;
  const long long n = (long long)pp.sz * pp.sy * pp.sx;
  if (n > INT32_MAX) return fail(ctx, SOFIMA_EINVAL, "correlation image too large");
  if (2 * pp.ry + 1 > pp.sy || 2 * pp.rx + 1 > pp.sx ||
      (pp.ndim == 3 && 2 * pp.rz + 1 > pp.sz))
    return fail(ctx, SOFIMA_EINVAL, "peak_radius window larger than the correlation image");
  void *v1 = nullptr, *p1 = nullptr, *bm = nullptr, *keys = nullptr, *nanf = nullptr;
  int rc;
  if ((rc = scratch(ctx, "flow.v1", sizeof(float) * B, &v1))) return rc;
  if ((rc = scratch(ctx, "flow.p1", sizeof(int) * B, &p1))) return rc;
  if ((rc = scratch(ctx, "flow.keys", sizeof(unsigned long long) * B, &keys))) return rc;
  if ((rc = scratch(ctx, "flow.nan", sizeof(int) * B, &nanf))) return rc;
  const size_t words = (size_t)((n + 31) / 32);
  if ((rc = scratch(ctx, "flow.bitmap", sizeof(unsigned) * words, &bm))) return rc;
  SOFIMA_CUDA(ctx, cudaMemsetAsync(bm, 0, sizeof(unsigned) * words, ctx->stream));
  if (!keys_ready) {
    LaunchTimer timer(ctx, "flow_peak1");
    SOFIMA_CUDA(ctx, cudaMemsetAsync(keys, 0, sizeof(unsigned long long) * B, ctx->stream));
    SOFIMA_CUDA(ctx, cudaMemsetAsync(nanf, 0, sizeof(int) * B, ctx->stream));
    long long slices = n / 65536;
    slices = slices < 1 ? 1 : (slices > 64 ? 64 : slices);
    peak1_kernel<<<dim3((unsigned)B, (unsigned)slices), kThreads, 0, ctx->stream>>>(
        images, pp, (unsigned long long*)keys, (int*)nanf);
    SOFIMA_CHECK_LAUNCH(ctx);
  }
  peak1_decode_kernel<<<(unsigned)ceil_div<long long>(B, 256), 256, 0, ctx->stream>>>(
      (const unsigned long long*)keys, (const int*)nanf, B, pp, (float*)v1, (int*)p1,
      (unsigned*)bm);
  SOFIMA_CHECK_LAUNCH(ctx);
  {
    LaunchTimer timer(ctx, "flow_peak2");
    const float* bandmax = nullptr;
    if (keys_ready && nbands > 0) {
      void* bmx = nullptr;
      if ((rc = scratch(ctx, "flow.bandmax", sizeof(float) * B * nbands, &bmx))) return rc;
      bandmax = static_cast<const float*>(bmx);
    }
    const unsigned long long* keys2 = nullptr;
    if (n >= (1ll << 20) && bandmax == nullptr) {  // 3-d volumes: spread the scan over the GPU
      void* k2 = nullptr;
      if ((rc = scratch(ctx, "flow.keys2", sizeof(unsigned long long) * B, &k2))) return rc;
      SOFIMA_CUDA(ctx, cudaMemsetAsync(k2, 0, sizeof(unsigned long long) * B, ctx->stream));
      long long slices = n / 65536;
      slices = slices > 64 ? 64 : slices;
      peak2_scan_kernel<<<dim3((unsigned)B, (unsigned)slices), kThreads, 0, ctx->stream>>>(
          images, pp, (const float*)v1, (const unsigned*)bm, (unsigned long long*)k2);
      SOFIMA_CHECK_LAUNCH(ctx);
      keys2 = static_cast<const unsigned long long*>(k2);
    }
    peak2_kernel<<<(unsigned)B, kThreads, 0, ctx->stream>>>(images, pp, (const float*)v1,
                                                            (const int*)p1, (const unsigned*)bm,
                                                            pp.ndim + 2, out_peaks, bandmax,
                                                            band_rows, nbands, keys2);
    SOFIMA_CHECK_LAUNCH(ctx);
  }
  return SOFIMA_OK;
}

static int check_params(sofima_ctx* ctx, const sofima_xcorr_params* p) {
  if (!p) return fail(ctx, SOFIMA_EINVAL, "params is NULL");
  if (p->ndim != 2 && p->ndim != 3) return fail(ctx, SOFIMA_EINVAL, "ndim must be 2 or 3");
  if (p->img_dtype != SOFIMA_U8 && p->img_dtype != SOFIMA_F32)
    return fail(ctx, SOFIMA_EINVAL, "img_dtype must be SOFIMA_U8 or SOFIMA_F32");
  for (int d = 0; d < p->ndim; ++d) {
    if (p->pre_patch[d] < 1 || p->post_patch[d] < 1)
      return fail(ctx, SOFIMA_EINVAL, "patch sizes must be positive");
    if (p->pre_patch[d] > p->pre_shape[d] || p->post_patch[d] > p->post_shape[d])
      return fail(ctx, SOFIMA_EINVAL, "patch larger than image");
    if (p->pre_shape[d] > INT32_MAX || p->post_shape[d] > INT32_MAX)
      return fail(ctx, SOFIMA_EINVAL, "image extent out of range");
  }
  if (p->min_distance < 0) return fail(ctx, SOFIMA_EINVAL, "min_distance < 0");
  return SOFIMA_OK;
}

// ---- fast path (flow_fast.cuh): unmasked, both FFT lengths of the form 16 * N2 ----
static bool fast_n2(int L, int* n2) {
  if (L % kN1) return false;
  const int v = L / kN1;
  for (int ok : {8, 10, 12, 15, 16, 20, 24, 25, 32})
    if (v == ok) { *n2 = v; return true; }
  return false;
}

// Transforms per block (row pairs resp. spectral columns): 8, or 4 for the long
// transforms so that the exchange buffers stay below the 48 KB static limit.
template <int N2>
struct FastLines {
  static constexpr int n = N2 <= 20 ? 8 : 4;
};

template <int N2>
static void launch_rows_fwd_fast(sofima_ctx* ctx, const Problem& P, const float2* tw, float2* T,
                                 int rp_max) {
  constexpr int TR = FastLines<N2>::n;
  const dim3 grid(ceil_div(rp_max, TR), P.nslots, P.nb);
  const bool half = 2 * P.img[0].pw <= FastDims<N2>::L && 2 * P.img[1].pw <= FastDims<N2>::L;
  if (half)
    rows_fwd_fast<N2, TR, true><<<grid, TR * FastDims<N2>::G, 0, ctx->stream>>>(P, tw, T);
  else
    rows_fwd_fast<N2, TR, false><<<grid, TR * FastDims<N2>::G, 0, ctx->stream>>>(P, tw, T);
}
template <int N2>
static void launch_cols_fast(sofima_ctx* ctx, const Problem& P, const float2* tw, const float2* T,
                             float2* U, const RowCacheView* rc) {
  constexpr int C = FastLines<N2>::n;
  constexpr int NT = C * FastDims<N2>::G;
  const dim3 grid(ceil_div(P.nkx, C), P.nb);
  const bool half = 2 * P.img[0].ph <= FastDims<N2>::L && 2 * P.img[1].ph <= FastDims<N2>::L;
  RowCacheView none;
  memset(&none, 0, sizeof(none));
  if (rc) {
    if (half)
      cols_fast<N2, C, true, true><<<grid, NT, 0, ctx->stream>>>(P, tw, T, U, *rc);
    else
      cols_fast<N2, C, false, true><<<grid, NT, 0, ctx->stream>>>(P, tw, T, U, *rc);
  } else if (half) {
    cols_fast<N2, C, true, false><<<grid, NT, 0, ctx->stream>>>(P, tw, T, U, none);
  } else {
    cols_fast<N2, C, false, false><<<grid, NT, 0, ctx->stream>>>(P, tw, T, U, none);
  }
}
template <int N2>
static void launch_rowspec_fast(sofima_ctx* ctx, const RowSpecJob& J, int slots, const float2* tw,
                                float2* out) {
  constexpr int TR = FastLines<N2>::n;
  const dim3 grid(ceil_div((J.h + 1) / 2, TR), slots);
  if (2 * J.pw <= FastDims<N2>::L)
    rowspec_fast<N2, TR, true><<<grid, TR * FastDims<N2>::G, 0, ctx->stream>>>(J, tw, out);
  else
    rowspec_fast<N2, TR, false><<<grid, TR * FastDims<N2>::G, 0, ctx->stream>>>(J, tw, out);
}
template <int N2>
static void launch_rows_inv_fast(sofima_ctx* ctx, const Problem& P, const float2* tw,
                                 const float2* U, float* images, float scale,
                                 unsigned long long* keys, int* nanflag, float* bandmax) {
  constexpr int TR = FastLines<N2>::n;
  rows_inv_fast<N2, TR><<<dim3(ceil_div((P.sy + 1) / 2, TR), 1, P.nb), TR * FastDims<N2>::G, 0,
                         ctx->stream>>>(P, tw, U, images, scale, keys, nanflag, bandmax);
}

#define SOFIMA_N2_SWITCH(n2, CALL)          \
  switch (n2) {                             \
    case 8: CALL(8); break;                 \
    case 10: CALL(10); break;               \
    case 12: CALL(12); break;               \
    case 15: CALL(15); break;               \
    case 16: CALL(16); break;               \
    case 20: CALL(20); break;               \
    case 24: CALL(24); break;               \
    case 25: CALL(25); break;               \
    case 32: CALL(32); break;               \
    default: break;                         \
  }

// Computes the correlation images of one batch into `images` [B][sy][sx].
static int run_xcorr(sofima_ctx* ctx, const sofima_xcorr_params* p, const void* pre_img,
                     const void* post_img, const uint8_t* pre_mask, const uint8_t* post_mask,
                     const int32_t* pre_starts, const int32_t* post_starts, long long B,
                     float* images, bool* keys_ready, int* band_rows = nullptr,
                     int* nbands = nullptr) {
  const bool masked = pre_mask != nullptr || post_mask != nullptr;
  if (keys_ready) *keys_ready = false;
  Problem P;
  memset(&P, 0, sizeof(P));
  P.dtype = p->img_dtype;
  const void* datas[2] = {pre_img, post_img};
  const uint8_t* masks[2] = {pre_mask, post_mask};
  const int64_t* shapes[2] = {p->pre_shape, p->post_shape};
  const int64_t* mshapes[2] = {p->pre_mask_shape, p->post_mask_shape};
  const int32_t* patches[2] = {p->pre_patch, p->post_patch};
  for (int i = 0; i < 2; ++i) {
    P.img[i].data = datas[i];
    P.img[i].mask = masks[i];
    P.img[i].h = (int)shapes[i][0];
    P.img[i].w = (int)shapes[i][1];
    P.img[i].ph = patches[i][0];
    P.img[i].pw = patches[i][1];
    if (masks[i]) {
      P.img[i].mh = (int)mshapes[i][0];
      P.img[i].mw = (int)mshapes[i][1];
      if (P.img[i].mh < P.img[i].ph || P.img[i].mw < P.img[i].pw)
        return fail(ctx, SOFIMA_EINVAL, "mask smaller than patch");
    }
  }
  P.starts[0] = pre_starts;
  P.starts[1] = post_starts;
  P.sy = p->pre_patch[0] + p->post_patch[0] - 1;
  P.sx = p->pre_patch[1] + p->post_patch[1] - 1;
  P.PY = p->pre_patch[0] > p->post_patch[0] ? p->pre_patch[0] : p->post_patch[0];
  const int Ly = next_fast_len(P.sy), Lx = next_fast_len(P.sx);
  P.nkx = Lx / 2 + 1;
  P.nslots = masked ? 6 : 2;
  const Slot slots[6] = {{0, 0}, {1, 0}, {0, 1}, {1, 1}, {0, 2}, {1, 2}};
  for (int i = 0; i < 6; ++i) P.slot[i] = slots[i];
  Products pr;
  memset(&pr, 0, sizeof(pr));
  if (masked) {
    // xcorr, overlap, mc_p, mc_c, p_sq, c_sq  (flow_field.py:85,114,118,119,124,128)
    const int a[6] = {0, 3, 3, 2, 3, 2}, b[6] = {1, 2, 0, 1, 4, 5};
    pr.nout = 6;
    for (int i = 0; i < 6; ++i) { pr.a[i] = a[i]; pr.b[i] = b[i]; }
  } else {
    pr.nout = 1;
    pr.a[0] = 0;
    pr.b[0] = 1;
  }

  FftPlan Fx, Fy;
  int rc;
  if ((rc = make_plan(ctx, Lx, &Fx))) return rc;
  if ((rc = make_plan(ctx, Ly, &Fy))) return rc;

  int n2x = 0, n2y = 0;
  const bool fast = !masked && fast_n2(Lx, &n2x) && fast_n2(Ly, &n2y);
  void *keys = nullptr, *nanf = nullptr;
  if (fast && keys_ready) {
    if ((rc = scratch(ctx, "flow.keys", sizeof(unsigned long long) * B, &keys))) return rc;
    if ((rc = scratch(ctx, "flow.nan", sizeof(int) * B, &nanf))) return rc;
    SOFIMA_CUDA(ctx, cudaMemsetAsync(keys, 0, sizeof(unsigned long long) * B, ctx->stream));
    SOFIMA_CUDA(ctx, cudaMemsetAsync(nanf, 0, sizeof(int) * B, ctx->stream));
    *keys_ready = true;
  }
  // Band maxima written by rows_inv_fast: one value per (pair, block of 2 TR rows).
  void* bandmax = nullptr;
  const int fast_tr = n2x <= 20 ? 8 : 4;  // == FastLines<N2>::n
  const int nb_bands = fast ? ceil_div((P.sy + 1) / 2, fast_tr) : 0;
  if (fast) {
    if ((rc = scratch(ctx, "flow.bandmax", sizeof(float) * B * nb_bands, &bandmax))) return rc;
    if (band_rows) *band_rows = 2 * fast_tr;
    if (nbands) *nbands = nb_bands;
  }

  // Shared-memory budgets.
  const size_t smem_cap = 200 * 1024;
  int R = 8;
  while (R > 1 && (size_t)(2 * R + 1) * Lx * sizeof(float2) > 96 * 1024) R /= 2;
  const size_t smem_rows = (size_t)(2 * R + 1) * Lx * sizeof(float2);
  int C = 8;
  while (C > 1 && (size_t)((2 + P.nslots) * C + 1) * Ly * sizeof(float2) > smem_cap) C /= 2;
  size_t smem_cols = (size_t)((2 + P.nslots) * C + 1) * Ly * sizeof(float2);
  // Long columns: one line + work buffer + twiddles per block, spectra through global.
  const bool long_cols = smem_cols > smem_cap;
  if (long_cols) smem_cols = (size_t)3 * Ly * sizeof(float2);
  if (smem_rows > smem_cap || smem_cols > smem_cap)
    return fail(ctx, SOFIMA_EUNSUPPORTED, "patch too large for the shared-memory FFT");
  if (long_cols) {
    SOFIMA_CUDA(ctx, cudaFuncSetAttribute(cols_fwd_long_kernel,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem_cap));
    SOFIMA_CUDA(ctx, cudaFuncSetAttribute(cols_inv_long_kernel,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem_cap));
  }
  SOFIMA_CUDA(ctx, cudaFuncSetAttribute(rows_fwd_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem_cap));
  SOFIMA_CUDA(ctx, cudaFuncSetAttribute(rows_inv_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem_cap));
  SOFIMA_CUDA(ctx, cudaFuncSetAttribute(cols_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem_cap));

  // Sub-batches of at most `scratch_mb` of spectra (T, U and the six masked outputs).
  const size_t per_pair = sizeof(float2) * ((size_t)P.nslots * P.PY * P.nkx +
                                            (size_t)pr.nout * P.sy * P.nkx) +
                          (masked ? sizeof(float) * 6 * (size_t)P.sy * P.sx : 0) +
                          (long_cols ? sizeof(float2) * (size_t)P.nslots * P.nkx * Ly : 0);
  // Sub-batches bound the scratch memory only; measured on B200 the pipeline is fastest
  // when a whole reference batch (1024 pairs = 0.84 GB of spectra) goes through each
  // stage in one launch -- streaming T / U through HBM costs less than small grids.
  size_t scratch_mb = 1536;
  if (const char* e = getenv("SOFIMA_FLOW_SCRATCH_MB")) scratch_mb = (size_t)atoi(e);
  long long nsub = (long long)((scratch_mb << 20) / per_pair);
  if (nsub < 1) nsub = 1;
  if (nsub > B) nsub = B;
  if (nsub > 65535) nsub = 65535;
  void *Tbuf = nullptr, *Ubuf = nullptr, *means = nullptr, *m6 = nullptr, *maxima = nullptr;
  if ((rc = scratch(ctx, "flow.T", sizeof(float2) * (size_t)P.nslots * nsub * P.PY * P.nkx,
                    &Tbuf))) return rc;
  if ((rc = scratch(ctx, "flow.U", sizeof(float2) * (size_t)pr.nout * nsub * P.sy * P.nkx,
                    &Ubuf))) return rc;
  if ((rc = scratch(ctx, "flow.means", sizeof(MeanPartial) * 2 * kMeanGroups * B, &means)))
    return rc;
  void* Sbuf = nullptr;
  if (long_cols &&
      (rc = scratch(ctx, "flow.S", sizeof(float2) * (size_t)P.nslots * nsub * P.nkx * Ly, &Sbuf)))
    return rc;
  P.parts = static_cast<const MeanPartial*>(means);
  P.has_mean = p->has_mean;
  P.mean = p->mean;
  const size_t img_elems = (size_t)P.sy * P.sx;
  float *den_all = nullptr, *ov_all = nullptr;
  if (masked) {
    // per sub-batch: mcc, psq, csq; whole batch: numerator (= images), denom, overlap.
    if ((rc = scratch(ctx, "flow.m3", sizeof(float) * 3 * nsub * img_elems, &m6))) return rc;
    void *d = nullptr, *o = nullptr;
    if ((rc = scratch(ctx, "flow.den", sizeof(float) * B * img_elems, &d))) return rc;
    if ((rc = scratch(ctx, "flow.ov", sizeof(float) * B * img_elems, &o))) return rc;
    den_all = static_cast<float*>(d);
    ov_all = static_cast<float*>(o);
    if ((rc = scratch(ctx, "flow.maxima", sizeof(float) * 2, &maxima))) return rc;
    SOFIMA_CUDA(ctx, cudaMemsetAsync(maxima, 0, sizeof(float) * 2, ctx->stream));
  }
  const float scale = (float)(1.0 / ((double)Lx * (double)Ly));

  // Row-spectra cache built by sofima_xcorr_rowcache for exactly these images?
  RowCacheView rcv;
  memset(&rcv, 0, sizeof(rcv));
  const sofima_ctx::RowCache& rcache = ctx->rowcache;
  bool cached = fast && rcache.valid && rcache.L == Lx && rcache.dtype == P.dtype;
  for (int i = 0; i < 2 && cached; ++i)
    cached = rcache.img[i] == P.img[i].data && rcache.h[i] == P.img[i].h &&
             rcache.w[i] == P.img[i].w && rcache.pw[i] == P.img[i].pw;
  if (cached) {
    for (int i = 0; i < 2; ++i) {
      rcv.spec[i] = rcache.spec[i];
      rcv.xindex[i] = rcache.xindex[i];
      rcv.h[i] = rcache.h[i];
    }
    rcv.pitch = rcache.pitch;
    rcv.fix = rcache.fix;
    void* mb = nullptr;
    if ((rc = scratch(ctx, "flow.rc_meta", sizeof(int4) * 2 * B, &mb))) return rc;
    rcv.meta = static_cast<const int4*>(mb);
  }
  const bool dc_mean = cached && !p->has_mean && P.dtype == SOFIMA_U8;
  if (!p->has_mean && !dc_mean) {  // patch sums of the whole batch in one launch
    for (long long b0 = 0; b0 < B; b0 += 65535) {
      P.b0 = b0;
      P.nb = (int)((B - b0 < 65535) ? (B - b0) : 65535);
      LaunchTimer timer(ctx, "flow_mean");
      if (P.dtype == SOFIMA_U8 && !P.img[0].mask && !P.img[1].mask)
        patch_sum_u8_kernel<<<dim3(P.nb, 2), kThreads, 0, ctx->stream>>>(
            P, (MeanPartial*)means);
      else
        patch_mean_kernel<<<dim3(P.nb, 2, kMeanGroups), kThreads, 0, ctx->stream>>>(
            P, (MeanPartial*)means);
      SOFIMA_CHECK_LAUNCH(ctx);
    }
  }
  if (cached) {
    LaunchTimer timer(ctx, "flow_mean");
    const unsigned int mblocks = (unsigned int)ceil_div<long long>(2 * B * 32, 256);
    if (dc_mean)
      rowcache_meta_kernel<true><<<mblocks, 256, 0, ctx->stream>>>(
          P, rcv, B, const_cast<int4*>(rcv.meta));
    else
      rowcache_meta_kernel<false><<<mblocks, 256, 0, ctx->stream>>>(
          P, rcv, B, const_cast<int4*>(rcv.meta));
    SOFIMA_CHECK_LAUNCH(ctx);
  }
  for (long long b0 = 0; b0 < B; b0 += nsub) {
    const int nb = (int)((B - b0 < nsub) ? (B - b0) : nsub);
    P.b0 = b0;
    P.nb = nb;
    const int rp_max = (P.PY + 1) / 2;
    if (fast) {
      if (!cached) {
        LaunchTimer timer(ctx, "flow_rows_fwd");
#define CALL(N) launch_rows_fwd_fast<N>(ctx, P, Fx.tw, (float2*)Tbuf, rp_max)
        SOFIMA_N2_SWITCH(n2x, CALL)
#undef CALL
        SOFIMA_CHECK_LAUNCH(ctx);
      }
      {
        LaunchTimer timer(ctx, "flow_cols");
#define CALL(N) \
  launch_cols_fast<N>(ctx, P, Fy.tw, (const float2*)Tbuf, (float2*)Ubuf, cached ? &rcv : nullptr)
        SOFIMA_N2_SWITCH(n2y, CALL)
#undef CALL
        SOFIMA_CHECK_LAUNCH(ctx);
      }
      {
        LaunchTimer timer(ctx, "flow_rows_inv");
        unsigned long long* kp = static_cast<unsigned long long*>(keys);
        int* np = static_cast<int*>(nanf);
        if (!kp) {  // caller does not want the fused peak search: throw-away slots
          void* tmp = nullptr;
          if ((rc = scratch(ctx, "flow.keys_tmp", (sizeof(unsigned long long) + sizeof(int)) * B,
                            &tmp))) return rc;
          kp = static_cast<unsigned long long*>(tmp);
          np = reinterpret_cast<int*>(kp + B);
        }
#define CALL(N) \
  launch_rows_inv_fast<N>(ctx, P, Fx.tw, (const float2*)Ubuf, images, scale, kp, np, (float*)bandmax)
        SOFIMA_N2_SWITCH(n2x, CALL)
#undef CALL
        SOFIMA_CHECK_LAUNCH(ctx);
      }
      continue;
    }
    {
      LaunchTimer timer(ctx, "flow_rows_fwd");
      rows_fwd_kernel<<<dim3(ceil_div(rp_max, R), P.nslots, nb), kThreads, smem_rows,
                        ctx->stream>>>(P, Fx, R, (float2*)Tbuf);
      SOFIMA_CHECK_LAUNCH(ctx);
    }
    if (long_cols) {
      LaunchTimer timer(ctx, "flow_cols");
      cols_fwd_long_kernel<<<dim3(P.nkx, P.nslots, nb), kThreads, smem_cols, ctx->stream>>>(
          P, Fy, (const float2*)Tbuf, (float2*)Sbuf);
      SOFIMA_CHECK_LAUNCH(ctx);
      cols_inv_long_kernel<<<dim3(P.nkx, pr.nout, nb), kThreads, smem_cols, ctx->stream>>>(
          P, Fy, pr, (const float2*)Sbuf, (float2*)Ubuf);
      SOFIMA_CHECK_LAUNCH(ctx);
    } else {
      LaunchTimer timer(ctx, "flow_cols");
      cols_kernel<<<dim3(ceil_div(P.nkx, C), nb), kThreads, smem_cols, ctx->stream>>>(
          P, Fy, C, pr, (const float2*)Tbuf, (float2*)Ubuf);
      SOFIMA_CHECK_LAUNCH(ctx);
    }
    Outputs outs;
    memset(&outs, 0, sizeof(outs));
    if (!masked) {
      outs.dst[0] = images;
    } else {
      // Outputs of this sub-batch are addressed with the global pair index b0 + i,
      // so sub-batch-local buffers are offset back by b0.
      float* loc = static_cast<float*>(m6);
      outs.dst[0] = images;                                   // xcorr -> numerator
      outs.dst[1] = ov_all;                                   // overlap
      outs.dst[2] = den_all;                                  // mc_p -> denom
      outs.dst[3] = loc + 0 * nsub * img_elems - b0 * img_elems;  // mc_c
      outs.dst[4] = loc + 1 * nsub * img_elems - b0 * img_elems;  // p_sq
      outs.dst[5] = loc + 2 * nsub * img_elems - b0 * img_elems;  // c_sq
    }
    {
      LaunchTimer timer(ctx, "flow_rows_inv");
      rows_inv_kernel<<<dim3(ceil_div((P.sy + 1) / 2, R), pr.nout, nb), kThreads, smem_rows,
                        ctx->stream>>>(P, Fx, R, (const float2*)Ubuf, outs, scale);
      SOFIMA_CHECK_LAUNCH(ctx);
    }
    if (masked) {
      const long long n = (long long)nb * img_elems;
      const size_t off = (size_t)b0 * img_elems;
      float* loc = static_cast<float*>(m6);
      {
        LaunchTimer timer(ctx, "flow_padfield_terms");
        padfield_terms_kernel<<<ctx->num_sms * 4, kThreads, 0, ctx->stream>>>(
            images + off, ov_all + off, den_all + off, loc, loc + nsub * img_elems,
            loc + 2 * nsub * img_elems, n, (float*)maxima);
        SOFIMA_CHECK_LAUNCH(ctx);
      }
    }
  }
  if (masked) {
    {
      LaunchTimer timer(ctx, "flow_padfield_norm");
      padfield_normalise_kernel<<<ctx->num_sms * 4, kThreads, 0, ctx->stream>>>(
          images, ov_all, den_all, (long long)B * img_elems, (const float*)maxima);
      SOFIMA_CHECK_LAUNCH(ctx);
    }
  }
  return SOFIMA_OK;
}

}  // namespace flow
}  // namespace sofima

#include "flow_fused.cuh"
#include "flow_rowspec_tc.cuh"

namespace sofima {
namespace flow {

template <int N2, int NG>
static int launch_pair_fused(sofima_ctx* ctx, int grid, const Problem& P, const float2* tw,
                             const RowCacheView& rc, float2* slots, int upitch, float scale,
                             const PeakParams& pp, PairPeaks* recs, int mode,
                             const unsigned* bitmap, const int* fixlist, const int* nfix,
                             float* out_peaks) {
  const size_t smem = FusedDims<N2, NG>::smem_bytes;
  SOFIMA_CUDA(ctx, cudaFuncSetAttribute(pair_fused<N2, NG>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pair_fused<N2, NG><<<grid, FusedDims<N2, NG>::NT, smem, ctx->stream>>>(
      P, tw, rc, slots, upitch, scale, pp, recs, mode, bitmap, fixlist, nfix, out_peaks);
  return SOFIMA_OK;
}

#define SOFIMA_FUSED_N2_SWITCH(n2, CALL) \
  switch (n2) {                          \
    case 8: CALL(8); break;              \
    case 10: CALL(10); break;            \
    case 12: CALL(12); break;            \
    case 15: CALL(15); break;            \
    case 16: CALL(16); break;            \
    case 20: CALL(20); break;            \
    default: break;                      \
  }

// Whole batch through the fused kernel (flow_fused.cuh).  *handled = false: the batch is not
// eligible (no row cache for these images, masks, unequal / long transforms) and nothing was
// launched -- the caller takes the three-kernel path.
static int run_fused(sofima_ctx* ctx, const sofima_xcorr_params* p, const void* pre_img,
                     const void* post_img, const int32_t* pre_starts,
                     const int32_t* post_starts, long long B, const PeakParams& pp,
                     float* out_peaks, bool* handled) {
  *handled = false;
  // Opt-in (SOFIMA_FLOW_FUSED=1).  Measured on a B200 (profiles/flow_fused_ab_r2.json): the
  // fused kernel is bit-identical to the three-kernel path and removes its HBM round trips,
  // but one pair per SM leaves 15-20 dependent warps to hide every latency, and the pipeline
  // is bound by instruction issue, not by HBM: 8.8 ms vs 5.05 ms per 9801 pairs.
  const char* fe = getenv("SOFIMA_FLOW_FUSED");
  if (!fe || fe[0] != '1') return SOFIMA_OK;
  if (p->ndim != 2 || B > INT32_MAX) return SOFIMA_OK;
  Problem P;
  memset(&P, 0, sizeof(P));
  P.dtype = p->img_dtype;
  const void* datas[2] = {pre_img, post_img};
  const int64_t* shapes[2] = {p->pre_shape, p->post_shape};
  const int32_t* patches[2] = {p->pre_patch, p->post_patch};
  for (int i = 0; i < 2; ++i) {
    P.img[i].data = datas[i];
    P.img[i].h = (int)shapes[i][0];
    P.img[i].w = (int)shapes[i][1];
    P.img[i].ph = patches[i][0];
    P.img[i].pw = patches[i][1];
  }
  P.starts[0] = pre_starts;
  P.starts[1] = post_starts;
  P.sy = p->pre_patch[0] + p->post_patch[0] - 1;
  P.sx = p->pre_patch[1] + p->post_patch[1] - 1;
  P.PY = p->pre_patch[0] > p->post_patch[0] ? p->pre_patch[0] : p->post_patch[0];
  const int Ly = next_fast_len(P.sy), Lx = next_fast_len(P.sx);
  P.nkx = Lx / 2 + 1;
  P.nslots = 2;
  P.has_mean = p->has_mean;
  P.mean = p->mean;
  P.b0 = 0;
  P.nb = (int)B;
  int n2 = 0;
  if (Lx != Ly || !fast_n2(Lx, &n2) || n2 > 20) return SOFIMA_OK;
  for (int i = 0; i < 2; ++i)
    if (2 * P.img[i].ph > Lx || 2 * P.img[i].pw > Lx) return SOFIMA_OK;
  if (2 * pp.ry + 1 > P.sy || 2 * pp.rx + 1 > P.sx) return SOFIMA_OK;  // run_peaks reports it
  const sofima_ctx::RowCache& rcache = ctx->rowcache;
  bool cached = rcache.valid && rcache.L == Lx && rcache.dtype == P.dtype;
  for (int i = 0; i < 2 && cached; ++i)
    cached = rcache.img[i] == P.img[i].data && rcache.h[i] == P.img[i].h &&
             rcache.w[i] == P.img[i].w && rcache.pw[i] == P.img[i].pw;
  if (!cached) return SOFIMA_OK;

  int rc;
  FftPlan Fx;
  if ((rc = make_plan(ctx, Lx, &Fx))) return rc;
  RowCacheView rcv;
  memset(&rcv, 0, sizeof(rcv));
  for (int i = 0; i < 2; ++i) {
    rcv.spec[i] = rcache.spec[i];
    rcv.xindex[i] = rcache.xindex[i];
    rcv.h[i] = rcache.h[i];
  }
  rcv.pitch = rcache.pitch;
  rcv.fix = rcache.fix;
  void *mb = nullptr, *means = nullptr, *slots = nullptr, *recs = nullptr, *bm = nullptr,
       *fix = nullptr;
  if ((rc = scratch(ctx, "flow.rc_meta", sizeof(int4) * 2 * B, &mb))) return rc;
  rcv.meta = static_cast<const int4*>(mb);
  const bool dc_mean = !p->has_mean && P.dtype == SOFIMA_U8;
  if (!p->has_mean && !dc_mean) {
    if ((rc = scratch(ctx, "flow.means", sizeof(MeanPartial) * 2 * kMeanGroups * B, &means)))
      return rc;
    P.parts = static_cast<const MeanPartial*>(means);
    for (long long b0 = 0; b0 < B; b0 += 65535) {
      Problem Q = P;
      Q.b0 = b0;
      Q.nb = (int)((B - b0 < 65535) ? (B - b0) : 65535);
      LaunchTimer timer(ctx, "flow_mean");
      patch_mean_kernel<<<dim3(Q.nb, 2, kMeanGroups), kThreads, 0, ctx->stream>>>(
          Q, (MeanPartial*)means);
      SOFIMA_CHECK_LAUNCH(ctx);
    }
  }
  {
    LaunchTimer timer(ctx, "flow_mean");
    const unsigned int mblocks = (unsigned int)ceil_div<long long>(2 * B * 32, 256);
    if (dc_mean)
      rowcache_meta_kernel<true><<<mblocks, 256, 0, ctx->stream>>>(P, rcv, B, (int4*)mb);
    else
      rowcache_meta_kernel<false><<<mblocks, 256, 0, ctx->stream>>>(P, rcv, B, (int4*)mb);
    SOFIMA_CHECK_LAUNCH(ctx);
  }
  const int upitch = (P.nkx + 7) & ~7;
  int grid = ctx->num_sms;
  if ((long long)grid > B) grid = (int)B;
  const int fix_grid = grid < 8 ? grid : 8;
  const size_t words = ((size_t)P.sy * P.sx + 31) / 32;
  if ((rc = scratch(ctx, "flow.fused_slots",
                    sizeof(float2) * (size_t)ctx->num_sms * P.sy * upitch, &slots))) return rc;
  if ((rc = scratch(ctx, "flow.fused_recs", sizeof(PairPeaks) * (size_t)B, &recs))) return rc;
  if ((rc = scratch(ctx, "flow.bitmap", sizeof(unsigned) * words, &bm))) return rc;
  if ((rc = scratch(ctx, "flow.fused_fix", sizeof(int) * ((size_t)B + 1), &fix))) return rc;
  int* nfix = static_cast<int*>(fix);
  int* fixlist = nfix + 1;
  SOFIMA_CUDA(ctx, cudaMemsetAsync(bm, 0, sizeof(unsigned) * words, ctx->stream));
  SOFIMA_CUDA(ctx, cudaMemsetAsync(nfix, 0, sizeof(int), ctx->stream));
  const float scale = (float)(1.0 / ((double)Lx * (double)Ly));
  int ng = 3;  // thread groups per block (flow_fused.cuh)
  if (const char* e = getenv("SOFIMA_FLOW_FUSED_GROUPS")) ng = atoi(e) == 4 ? 4 : 3;
  {
    LaunchTimer timer(ctx, "flow_fused");
#define CALL(N)                                                                               \
  rc = ng == 4 ? launch_pair_fused<N, 4>(ctx, grid, P, Fx.tw, rcv, (float2*)slots, upitch,    \
                                         scale, pp, (PairPeaks*)recs, 0, nullptr, nullptr,    \
                                         nullptr, nullptr)                                    \
               : launch_pair_fused<N, 3>(ctx, grid, P, Fx.tw, rcv, (float2*)slots, upitch,    \
                                         scale, pp, (PairPeaks*)recs, 0, nullptr, nullptr,    \
                                         nullptr, nullptr)
    SOFIMA_FUSED_N2_SWITCH(n2, CALL)
#undef CALL
    if (rc) return rc;
    SOFIMA_CHECK_LAUNCH(ctx);
  }
  {
    LaunchTimer timer(ctx, "flow_fused_select");
    const unsigned nblk = (unsigned)ceil_div<long long>(B, 256);
    fused_bitmap_kernel<<<nblk, 256, 0, ctx->stream>>>((const PairPeaks*)recs, B, (unsigned*)bm);
    SOFIMA_CHECK_LAUNCH(ctx);
    fused_finalize_kernel<<<nblk, 256, 0, ctx->stream>>>((const PairPeaks*)recs, B, pp,
                                                         (const unsigned*)bm, out_peaks, fixlist,
                                                         nfix);
    SOFIMA_CHECK_LAUNCH(ctx);
  }
  {
    LaunchTimer timer(ctx, "flow_fused_fixup");
#define CALL(N)                                                                               \
  rc = launch_pair_fused<N, 3>(ctx, fix_grid, P, Fx.tw, rcv, (float2*)slots, upitch, scale,   \
                               pp, (PairPeaks*)recs, 1, (const unsigned*)bm, fixlist, nfix,   \
                               out_peaks)
    SOFIMA_FUSED_N2_SWITCH(n2, CALL)
#undef CALL
    if (rc) return rc;
    SOFIMA_CHECK_LAUNCH(ctx);
  }
  *handled = true;
  return SOFIMA_OK;
}

}  // namespace flow
}  // namespace sofima

#include "flow3d.cuh"
#include "flow3d_masked.cuh"

namespace sofima {
namespace flow {
// Register-codelet axis pass (flow3d.cuh) for L = 16 * N2; false if the length has none.
template <int N2>
static bool launch_axis_fast_n2(sofima_ctx* ctx, bool inverse, float2* data, long long nlines,
                                long long inner, long long istride, long long ostride,
                                long long es, const float2* tw, const LinePrune& pr) {
  using D = AxisFast<N2>;
  if (D::smem > 48 * 1024) {  // opt-in is per device: set it on every call (cheap)
    const cudaError_t e = inverse
        ? cudaFuncSetAttribute(axis_fft_fast_kernel<N2, true>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)D::smem)
        : cudaFuncSetAttribute(axis_fft_fast_kernel<N2, false>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)D::smem);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return false;  // the generic pass takes over
    }
  }
  const unsigned grid = (unsigned)ceil_div<long long>(nlines, D::C);
  if (inverse)
    axis_fft_fast_kernel<N2, true><<<grid, D::NT, D::smem, ctx->stream>>>(
        data, nlines, inner, istride, ostride, es, tw, pr);
  else
    axis_fft_fast_kernel<N2, false><<<grid, D::NT, D::smem, ctx->stream>>>(
        data, nlines, inner, istride, ostride, es, tw, pr);
  return true;
}

static bool launch_axis_fast(sofima_ctx* ctx, int L, bool inverse, float2* data, long long nlines,
                             long long inner, long long istride, long long ostride, long long es,
                             const float2* tw, const LinePrune& pr) {
  if (const char* e = getenv("SOFIMA_FLOW3D_FAST"))
    if (e[0] == '0') return false;
  int n2 = 0;
  if (!fast_n2(L, &n2)) return false;
  switch (n2) {
    case 8: return launch_axis_fast_n2<8>(ctx, inverse, data, nlines, inner, istride, ostride, es, tw, pr);
    case 10: return launch_axis_fast_n2<10>(ctx, inverse, data, nlines, inner, istride, ostride, es, tw, pr);
    case 12: return launch_axis_fast_n2<12>(ctx, inverse, data, nlines, inner, istride, ostride, es, tw, pr);
    case 15: return launch_axis_fast_n2<15>(ctx, inverse, data, nlines, inner, istride, ostride, es, tw, pr);
    case 16: return launch_axis_fast_n2<16>(ctx, inverse, data, nlines, inner, istride, ostride, es, tw, pr);
    case 20: return launch_axis_fast_n2<20>(ctx, inverse, data, nlines, inner, istride, ostride, es, tw, pr);
    default: return false;
  }
}
}  // namespace flow
}  // namespace sofima

namespace sofima {
namespace flow {

static int run_xcorr3_masked(sofima_ctx* ctx, const sofima_xcorr_params* p, const void* pre_img,
                             const void* post_img, const uint8_t* pre_mask,
                             const uint8_t* post_mask, const int32_t* pre_starts,
                             const int32_t* post_starts, long long B, float* images);

// 3-d correlation images of one batch into `images` [B][sz][sy][sx] (unmasked).
static int run_xcorr3(sofima_ctx* ctx, const sofima_xcorr_params* p, const void* pre_img,
                      const void* post_img, const uint8_t* pre_mask, const uint8_t* post_mask,
                      const int32_t* pre_starts, const int32_t* post_starts, long long B,
                      float* images) {
  if (pre_mask || post_mask) {
    return run_xcorr3_masked(ctx, p, pre_img, post_img, pre_mask, post_mask, pre_starts,
                             post_starts, B, images);
  }
  Problem3 P;
  memset(&P, 0, sizeof(P));
  P.dtype = p->img_dtype;
  const void* datas[2] = {pre_img, post_img};
  const int64_t* shapes[2] = {p->pre_shape, p->post_shape};
  const int32_t* patches[2] = {p->pre_patch, p->post_patch};
  for (int i = 0; i < 2; ++i) {
    P.img[i].data = datas[i];
    P.img[i].d = (int)shapes[i][0]; P.img[i].h = (int)shapes[i][1]; P.img[i].w = (int)shapes[i][2];
    P.img[i].pd = patches[i][0]; P.img[i].ph = patches[i][1]; P.img[i].pw = patches[i][2];
  }
  P.starts[0] = pre_starts;
  P.starts[1] = post_starts;
  P.has_mean = p->has_mean;
  P.mean = p->mean;
  P.sz = p->pre_patch[0] + p->post_patch[0] - 1;
  P.sy = p->pre_patch[1] + p->post_patch[1] - 1;
  P.sx = p->pre_patch[2] + p->post_patch[2] - 1;
  P.Lz = next_fast_len(P.sz); P.Ly = next_fast_len(P.sy); P.Lx = next_fast_len(P.sx);
  FftPlan Fz, Fy, Fx;
  int rc;
  if ((rc = make_plan(ctx, P.Lz, &Fz))) return rc;
  if ((rc = make_plan(ctx, P.Ly, &Fy))) return rc;
  if ((rc = make_plan(ctx, P.Lx, &Fx))) return rc;
  const long long vol = (long long)P.Lz * P.Ly * P.Lx;
  long long nsub = (long long)((512ull << 20) / (2 * vol * sizeof(float2)));
  if (nsub < 1) nsub = 1;
  if (nsub > B) nsub = B;
  void *Z = nullptr, *means = nullptr;
  if ((rc = scratch(ctx, "flow3.Z", sizeof(float2) * 2 * nsub * vol, &Z))) return rc;
  if ((rc = scratch(ctx, "flow3.mean_parts", sizeof(double) * 2 * B * kMeanSlices, &means)))
    return rc;
  const size_t smem_cap = 200 * 1024;
  SOFIMA_CUDA(ctx, cudaFuncSetAttribute(axis_fft_kernel<false>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem_cap));
  SOFIMA_CUDA(ctx, cudaFuncSetAttribute(axis_fft_kernel<true>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem_cap));
  auto lines_per_block = [&](int L) {
    int C = 16;
    while (C > 1 && ((size_t)2 * C * (L | 1) + L) * sizeof(float2) > 96 * 1024) C /= 2;
    return C;
  };
  const float scale = (float)(1.0 / ((double)P.Lx * P.Ly * P.Lz));
  for (long long b0 = 0; b0 < B; b0 += nsub) {
    const int nb = (int)((B - b0 < nsub) ? (B - b0) : nsub);
    P.b0 = b0;
    P.nb = nb;
    float2* Zp = static_cast<float2*>(Z);
    {
      LaunchTimer timer(ctx, "flow3_pack");
      patch_mean3_kernel<<<dim3(nb, 2, kMeanSlices), kThreads, 0, ctx->stream>>>(
          P, (double*)means);
      SOFIMA_CHECK_LAUNCH(ctx);
      pack3_kernel<<<dim3(ctx->num_sms * 2, 2, nb), kThreads, 0, ctx->stream>>>(
          P, (const double*)means, Zp);
      SOFIMA_CHECK_LAUNCH(ctx);
    }
    // transforms over x, y, z of all 2 * nb volumes (forward), then of the nb products.
    auto transform = [&](bool inverse, long long nvol) -> int {
      struct Axis {
        const FftPlan* F;
        long long nlines, inner, istride, ostride, es;
        LinePrune pr;
      };
      const long long plane = (long long)P.Ly * P.Lx;
      const LinePrune all = {1, 1, 1, 1};
      // Forward: the patches fill z < pd, y < ph, x < pw of the padded volumes.  The x pass
      // only has to touch the lines with z < pd and y < ph, the y pass those with z < pd (all
      // other lines are zero and their transform is zero): 1.75 instead of 3 passes' worth of
      // traffic for 80^3 patches in 160^3 volumes.  The inverse needs every line.
      const long long pd = P.img[0].pd > P.img[1].pd ? P.img[0].pd : P.img[1].pd;
      const long long ph = P.img[0].ph > P.img[1].ph ? P.img[0].ph : P.img[1].ph;
      const LinePrune px = {P.Lz, pd, P.Ly, ph}, py = {P.Lz, pd, P.Lx, P.Lx};
      const Axis axes[3] = {
          {&Fx, inverse ? nvol * P.Lz * P.Ly : nvol * pd * ph, 1, 0, P.Lx, 1,
           inverse ? all : px},                                           // x: contiguous lines
          {&Fy, inverse ? nvol * P.Lz * P.Lx : nvol * pd * P.Lx, P.Lx, 1, plane, P.Lx,
           inverse ? all : py},                                           // y: lines (z, x)
          {&Fz, nvol * plane, plane, 1, vol, plane, all},                 // z: lines (y, x)
      };
      for (int a = 0; a < 3; ++a) {
        const Axis& A = axes[inverse ? 2 - a : a];
        const int C = lines_per_block(A.F->L);
        const size_t smem = ((size_t)2 * C * (A.F->L | 1) + A.F->L) * sizeof(float2);
        const unsigned grid = (unsigned)ceil_div<long long>(A.nlines, C);
        LaunchTimer timer(ctx, "flow3_fft");
        if (launch_axis_fast(ctx, A.F->L, inverse, Zp, A.nlines, A.inner, A.istride, A.ostride,
                             A.es, A.F->tw, A.pr)) {
          SOFIMA_CHECK_LAUNCH(ctx);
          continue;
        }
        if (inverse)
          axis_fft_kernel<true><<<grid, kThreads, smem, ctx->stream>>>(
              Zp, A.nlines, A.inner, A.istride, A.ostride, A.es, *A.F, C, A.pr);
        else
          axis_fft_kernel<false><<<grid, kThreads, smem, ctx->stream>>>(
              Zp, A.nlines, A.inner, A.istride, A.ostride, A.es, *A.F, C, A.pr);
        SOFIMA_CHECK_LAUNCH(ctx);
      }
      return SOFIMA_OK;
    };
    if ((rc = transform(false, 2ll * nb))) return rc;
    {
      LaunchTimer timer(ctx, "flow3_mul");
      multiply3_kernel<<<ctx->num_sms * 4, kThreads, 0, ctx->stream>>>(Zp, Zp + (long long)nb * vol,
                                                                        (long long)nb * vol);
      SOFIMA_CHECK_LAUNCH(ctx);
    }
    if ((rc = transform(true, nb))) return rc;
    {
      LaunchTimer timer(ctx, "flow3_crop");
      crop3_kernel<<<dim3(ctx->num_sms * 2, nb), kThreads, 0, ctx->stream>>>(P, Zp, images, scale);
      SOFIMA_CHECK_LAUNCH(ctx);
    }
  }
  return SOFIMA_OK;
}

static void peak_params(const sofima_xcorr_params* p, PeakParams* pp) {
  memset(pp, 0, sizeof(*pp));
  const int nd = p->ndim, o = 3 - nd;  // o: offset of the y axis slot in (z, y, x)
  int s[3] = {1, 1, 1}, r[3] = {0, 0, 0}, c[3] = {0, 0, 0};
  for (int d = 0; d < nd; ++d) {
    s[o + d] = p->pre_patch[d] + p->post_patch[d] - 1;
    r[o + d] = p->peak_radius[d];
    // center_offset = (pre + post) // 2 - 1, flow_field.py:357-360
    c[o + d] = (p->pre_patch[d] + p->post_patch[d]) / 2 - 1;
  }
  pp->ndim = nd;
  pp->sz = s[0]; pp->sy = s[1]; pp->sx = s[2];
  pp->rz = r[0]; pp->ry = r[1]; pp->rx = r[2];
  pp->cz = c[0]; pp->cy = c[1]; pp->cx = c[2];
  pp->md = p->min_distance;
  pp->thr_rel = p->threshold_rel;
}

// Masked 3-d correlation images of one batch (flow_field.py:91-155, dim = 3); see
// flow3d_masked.cuh.  Same contract as run_xcorr3.
static int run_xcorr3_masked(sofima_ctx* ctx, const sofima_xcorr_params* p, const void* pre_img,
                             const void* post_img, const uint8_t* pre_mask,
                             const uint8_t* post_mask, const int32_t* pre_starts,
                             const int32_t* post_starts, long long B, float* images) {
  Problem3M P;
  memset(&P, 0, sizeof(P));
  P.dtype = p->img_dtype;
  const void* datas[2] = {pre_img, post_img};
  const uint8_t* masks[2] = {pre_mask, post_mask};
  const int64_t* shapes[2] = {p->pre_shape, p->post_shape};
  const int64_t* mshapes[2] = {p->pre_mask_shape, p->post_mask_shape};
  const int32_t* patches[2] = {p->pre_patch, p->post_patch};
  for (int i = 0; i < 2; ++i) {
    Vol3M& I = P.img[i];
    I.data = datas[i];
    I.mask = masks[i];
    I.d = (int)shapes[i][0]; I.h = (int)shapes[i][1]; I.w = (int)shapes[i][2];
    I.pd = patches[i][0]; I.ph = patches[i][1]; I.pw = patches[i][2];
    I.md = masks[i] ? (int)mshapes[i][0] : I.d;
    I.mh = masks[i] ? (int)mshapes[i][1] : I.h;
    I.mw = masks[i] ? (int)mshapes[i][2] : I.w;
    if (masks[i] && (I.md < I.pd || I.mh < I.ph || I.mw < I.pw))
      return fail(ctx, SOFIMA_EINVAL, "mask smaller than the patch");
  }
  P.starts[0] = pre_starts;
  P.starts[1] = post_starts;
  P.has_mean = p->has_mean;
  P.mean = p->mean;
  P.sz = p->pre_patch[0] + p->post_patch[0] - 1;
  P.sy = p->pre_patch[1] + p->post_patch[1] - 1;
  P.sx = p->pre_patch[2] + p->post_patch[2] - 1;
  P.Lz = next_fast_len(P.sz); P.Ly = next_fast_len(P.sy); P.Lx = next_fast_len(P.sx);
  FftPlan Fz, Fy, Fx;
  int rc;
  if ((rc = make_plan(ctx, P.Lz, &Fz))) return rc;
  if ((rc = make_plan(ctx, P.Ly, &Fy))) return rc;
  if ((rc = make_plan(ctx, P.Lx, &Fx))) return rc;
  const long long vol = (long long)P.Lz * P.Ly * P.Lx;
  const size_t img_elems = (size_t)P.sz * P.sy * P.sx;
  // per pair: 6 forward + 6 product volumes
  long long nsub = (long long)((1024ull << 20) / (12 * vol * sizeof(float2)));
  if (nsub < 1) nsub = 1;
  if (nsub > B) nsub = B;
  void *Z = nullptr, *W = nullptr, *means = nullptr, *ov = nullptr, *den = nullptr,
       *loc = nullptr, *maxima = nullptr;
  if ((rc = scratch(ctx, "flow3m.Z", sizeof(float2) * 6 * nsub * vol, &Z))) return rc;
  if ((rc = scratch(ctx, "flow3m.W", sizeof(float2) * 6 * nsub * vol, &W))) return rc;
  if ((rc = scratch(ctx, "flow3m.means", sizeof(float) * 2 * B, &means))) return rc;
  if ((rc = scratch(ctx, "flow3m.ov", sizeof(float) * B * img_elems, &ov))) return rc;
  if ((rc = scratch(ctx, "flow3m.den", sizeof(float) * B * img_elems, &den))) return rc;
  if ((rc = scratch(ctx, "flow3m.loc", sizeof(float) * 3 * nsub * img_elems, &loc))) return rc;
  if ((rc = scratch(ctx, "flow3m.maxima", sizeof(float) * 2, &maxima))) return rc;
  SOFIMA_CUDA(ctx, cudaMemsetAsync(maxima, 0, sizeof(float) * 2, ctx->stream));
  const size_t smem_cap = 200 * 1024;
  SOFIMA_CUDA(ctx, cudaFuncSetAttribute(axis_fft_kernel<false>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem_cap));
  SOFIMA_CUDA(ctx, cudaFuncSetAttribute(axis_fft_kernel<true>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem_cap));
  auto lines_per_block = [&](int L) {
    int C = 16;
    while (C > 1 && ((size_t)2 * C * (L | 1) + L) * sizeof(float2) > 96 * 1024) C /= 2;
    return C;
  };
  // in-place transforms over x, y, z (forward) / z, y, x (inverse) of nvol volumes
  auto transform = [&](float2* data, bool inverse, long long nvol) -> int {
    struct Axis { const FftPlan* F; long long nlines, inner, istride, ostride, es; };
    const long long plane = (long long)P.Ly * P.Lx;
    const Axis axes[3] = {
        {&Fx, nvol * P.Lz * P.Ly, 1, 0, P.Lx, 1},
        {&Fy, nvol * P.Lz * P.Lx, P.Lx, 1, plane, P.Lx},
        {&Fz, nvol * plane, plane, 1, vol, plane},
    };
    for (int a = 0; a < 3; ++a) {
      const Axis& A = axes[inverse ? 2 - a : a];
      const int C = lines_per_block(A.F->L);
      const size_t smem = ((size_t)2 * C * (A.F->L | 1) + A.F->L) * sizeof(float2);
      const unsigned grid = (unsigned)ceil_div<long long>(A.nlines, C);
      LaunchTimer timer(ctx, "flow3_fft");
      if (inverse)
        axis_fft_kernel<true><<<grid, kThreads, smem, ctx->stream>>>(
            data, A.nlines, A.inner, A.istride, A.ostride, A.es, *A.F, C, LinePrune{1, 1, 1, 1});
      else
        axis_fft_kernel<false><<<grid, kThreads, smem, ctx->stream>>>(
            data, A.nlines, A.inner, A.istride, A.ostride, A.es, *A.F, C, LinePrune{1, 1, 1, 1});
      SOFIMA_CHECK_LAUNCH(ctx);
    }
    return SOFIMA_OK;
  };
  const float scale = (float)(1.0 / ((double)P.Lx * P.Ly * P.Lz));
  float* ov_all = static_cast<float*>(ov);
  float* den_all = static_cast<float*>(den);
  float* locf = static_cast<float*>(loc);
  for (long long b0 = 0; b0 < B; b0 += nsub) {
    const int nb = (int)((B - b0 < nsub) ? (B - b0) : nsub);
    P.b0 = b0;
    P.nb = nb;
    float2* Zp = static_cast<float2*>(Z);
    float2* Wp = static_cast<float2*>(W);
    {
      LaunchTimer timer(ctx, "flow3_pack");
      patch_mean3m_kernel<<<dim3(nb, 2), kThreads, 0, ctx->stream>>>(P, (float*)means);
      SOFIMA_CHECK_LAUNCH(ctx);
      pack3m_kernel<<<dim3(ctx->num_sms * 2, 6, nb), kThreads, 0, ctx->stream>>>(
          P, (const float*)means, Zp);
      SOFIMA_CHECK_LAUNCH(ctx);
    }
    if ((rc = transform(Zp, false, 6ll * nb))) return rc;
    {
      LaunchTimer timer(ctx, "flow3_mul");
      multiply3m_kernel<<<ctx->num_sms * 4, kThreads, 0, ctx->stream>>>(Zp, Wp,
                                                                        (long long)nb * vol);
      SOFIMA_CHECK_LAUNCH(ctx);
    }
    if ((rc = transform(Wp, true, 6ll * nb))) return rc;
    Outputs3M outs;
    // whole-batch arrays are addressed with the global pair index, sub-batch-local ones
    // with the index inside the sub-batch
    outs.dst[0] = images;          outs.first[0] = b0;   // numerator -> result
    outs.dst[1] = ov_all;          outs.first[1] = b0;   // overlap
    outs.dst[2] = den_all;         outs.first[2] = b0;   // mc_p -> denominator
    outs.dst[3] = locf;                              outs.first[3] = 0;  // mc_c
    outs.dst[4] = locf + (size_t)nsub * img_elems;   outs.first[4] = 0;  // p_sq
    outs.dst[5] = locf + 2 * (size_t)nsub * img_elems; outs.first[5] = 0;  // c_sq
    {
      LaunchTimer timer(ctx, "flow3_crop");
      crop3m_kernel<<<dim3(ctx->num_sms, 6, nb), kThreads, 0, ctx->stream>>>(P, Wp, outs, scale);
      SOFIMA_CHECK_LAUNCH(ctx);
    }
    {
      const long long n = (long long)nb * img_elems;
      const size_t off = (size_t)b0 * img_elems;
      LaunchTimer timer(ctx, "flow_padfield_terms");
      padfield_terms_kernel<<<ctx->num_sms * 4, kThreads, 0, ctx->stream>>>(
          images + off, ov_all + off, den_all + off, locf, locf + (size_t)nsub * img_elems,
          locf + 2 * (size_t)nsub * img_elems, n, (float*)maxima);
      SOFIMA_CHECK_LAUNCH(ctx);
    }
  }
  {
    LaunchTimer timer(ctx, "flow_padfield_norm");
    padfield_normalise_kernel<<<ctx->num_sms * 4, kThreads, 0, ctx->stream>>>(
        images, ov_all, den_all, (long long)B * img_elems, (const float*)maxima);
    SOFIMA_CHECK_LAUNCH(ctx);
  }
  return SOFIMA_OK;
}

}  // namespace flow
}  // namespace sofima

namespace sofima {
namespace flow {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) !=
            cudaSuccess || q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// Small integer tables (patch x starts, slot lists) travel to the device as kernel
// arguments: a cudaMemcpyAsync from pageable memory synchronises the stream first, which
// would serialise back-to-back flow_field calls on short strips (BASELINE config 2).
struct SmallInts {
  int v[240];
};
__global__ void upload_small_kernel(SmallInts s, int n, int* dst) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = s.v[i];
}
// table[x] = slot of x start x, -1 elsewhere
__global__ void xindex_kernel(const int* __restrict__ xstarts, int n, int* __restrict__ table,
                              int w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < w) {
    int slot = -1;
    for (int j = 0; j < n; ++j) slot = xstarts[j] == i ? j : slot;
    table[i] = slot;
  }
}
static int upload_ints(sofima_ctx* ctx, const int* host, int n, int* dst) {
  if (n <= 240) {
    SmallInts s;
    memcpy(s.v, host, sizeof(int) * n);
    upload_small_kernel<<<1, 256, 0, ctx->stream>>>(s, n, dst);
    SOFIMA_CHECK_LAUNCH(ctx);
    return SOFIMA_OK;
  }
  SOFIMA_CUDA(ctx, cudaMemcpyAsync(dst, host, sizeof(int) * n, cudaMemcpyHostToDevice,
                                   ctx->stream));
  SOFIMA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the caller reuses `host`
  return SOFIMA_OK;
}

// Twiddle digits of one transform length / patch width for rowspec_tc_kernel:
// [ndelta][nchunks][kTcN x K] s8 in the MMA's no-swizzle K-major layout
// (flow_rowspec_tc.cuh); matrix row x' of the table for `delta` is pixel x' - delta.
static std::vector<int8_t> rowspec_tc_table(int L, int pw, int K, int nkx, int nchunks,
                                            const std::vector<int>& deltas) {
  std::vector<int8_t> tab(deltas.size() * (size_t)nchunks * kTcN * K, 0);
  const size_t lbo = 128, sbo = 128 * (size_t)(K / 16);
  // digits of every twiddle exp(-2 pi i m / L), m < L
  std::vector<int8_t> dig((size_t)L * 8);
  for (int m = 0; m < L; ++m) {
    const double ph = -2.0 * M_PI * (double)m / (double)L;
    for (int reim = 0; reim < 2; ++reim) {
      long long T = llround((reim == 0 ? cos(ph) : sin(ph)) * 134217728.0);  // 2^27
      int d[4];
      for (int i = 3; i >= 1; --i) {
        const long long r = ((T + 64) % 128 + 128) % 128 - 64;  // balanced digit in [-64, 63]
        d[i] = (int)r;
        T = (T - r) / 128;
      }
      d[0] = (int)T;  // |d0| <= 64 (+1 from a carry)
      for (int i = 0; i < 4; ++i) dig[(size_t)m * 8 + reim * 4 + i] = (int8_t)d[i];
    }
  }
  for (size_t di = 0; di < deltas.size(); ++di) {
    for (int c = 0; c < nchunks; ++c) {
      int8_t* t = tab.data() + (di * nchunks + c) * (size_t)kTcN * K;
      for (int n = 0; n < kTcN; ++n) {
        const int bin = c * kTcBins + n / 8;
        if (bin >= nkx) continue;
        for (int x = 0; x < pw; ++x) {
          const int xp = x + deltas[di];
          const int m = (int)((long long)x * bin % L);
          t[(n / 8) * sbo + (xp / 16) * lbo + (n % 8) * 16 + (xp % 16)] = dig[(size_t)m * 8 + n % 8];
        }
      }
    }
  }
  return tab;
}

// Row spectra of one image on the tensor cores.  *done = false: not eligible (dtype, image
// pitch, patch width, no driver entry point) -- the caller runs rowspec_fast instead.
static int rowspec_tc(sofima_ctx* ctx, int which, const void* img, int dtype, int h, int w,
                      int pw, int L, const int32_t* xstarts_host, const int* xstarts_dev,
                      int nslots, int pitch, float2* out, bool* done) {
  *done = false;
  if (const char* e = getenv("SOFIMA_FLOW_ROWSPEC_TC"))
    if (e[0] == '0') return SOFIMA_OK;
  if (dtype != SOFIMA_U8 || (w % 16) != 0 || ((uintptr_t)img & 15) != 0 || nslots < 1 || h < 1)
    return SOFIMA_OK;
  // distinct misalignments of the x starts and the slots of each
  std::vector<int> deltas, didx(nslots);
  for (int j = 0; j < nslots; ++j) {
    const int d = xstarts_host[j] & 15;
    size_t k = 0;
    while (k < deltas.size() && deltas[k] != d) ++k;
    if (k == deltas.size()) deltas.push_back(d);
    didx[j] = (int)k;
  }
  int dmax = 0, dmask = 0;
  for (int d : deltas) { dmax = d > dmax ? d : dmax; dmask |= 1 << d; }
  const int K = (dmax + pw + 31) & ~31;
  const int nd = (int)deltas.size();
  if (K > kTcMaxK) return SOFIMA_OK;
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return SOFIMA_OK;
  const int nkx = L / 2 + 1;
  const int nchunks = (nkx + kTcBins - 1) / kTcBins;
  if (nd * nchunks > ctx->num_sms) return SOFIMA_OK;
  // digit tables, cached per (L, pw, K, set of deltas) and image slot
  char name[64];
  snprintf(name, sizeof(name), "flow.tc_tab%d", which);
  void *tab = nullptr, *dsl = nullptr;
  const size_t tbytes = (size_t)nd * nchunks * kTcN * K;
  int rc;
  if ((rc = scratch(ctx, name, tbytes, &tab))) return rc;
  int* key = ctx->tc_tab_key[which];
  if (key[0] != L || key[1] != pw || key[2] != K || key[3] != dmask) {
    const std::vector<int8_t> host = rowspec_tc_table(L, pw, K, nkx, nchunks, deltas);
    SOFIMA_CUDA(ctx, cudaMemcpyAsync(tab, host.data(), tbytes, cudaMemcpyHostToDevice,
                                     ctx->stream));
    SOFIMA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    key[0] = L; key[1] = pw; key[2] = K; key[3] = dmask;
  }
  // [nd + 1] offsets, then the slot indices grouped by delta
  std::vector<int> dslots(nd + 1 + nslots, 0);
  for (int j = 0; j < nslots; ++j) dslots[didx[j] + 1]++;
  for (int k = 0; k < nd; ++k) dslots[k + 1] += dslots[k];
  {
    std::vector<int> fill(dslots.begin(), dslots.begin() + nd);
    for (int j = 0; j < nslots; ++j) dslots[nd + 1 + fill[didx[j]]++] = j;
  }
  snprintf(name, sizeof(name), "flow.tc_dslots%d", which);
  if ((rc = scratch(ctx, name, sizeof(int) * dslots.size(), &dsl))) return rc;
  if ((rc = upload_ints(ctx, dslots.data(), (int)dslots.size(), static_cast<int*>(dsl))))
    return rc;
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  const cuuint64_t dims[2] = {(cuuint64_t)w, (cuuint64_t)h};
  const cuuint64_t strides[1] = {(cuuint64_t)w};
  const cuuint32_t box[2] = {32, (cuuint32_t)kTcRows}, es[2] = {1, 1};
  const CUresult cr = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(img), dims,
                          strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return SOFIMA_OK;  // odd geometry: plain path
  RowSpecTcJob J;
  J.h = h; J.nslots = nslots; J.pitch = pitch; J.nkx = nkx; J.K = K; J.nchunks = nchunks;
  J.ndelta = nd;
  J.xstarts = xstarts_dev;
  J.dslots = static_cast<const int*>(dsl);
  J.btab = static_cast<const int8_t*>(tab);
  // > half of the SM's shared memory: one block per SM (a block owns all 512 TMEM columns)
  // (more than half of an SM's shared memory for the usual K: one block per SM, which owns
  // all 512 TMEM columns)
  const size_t smem = tc_smem_bytes(K);
  if (smem > 220 * 1024) return SOFIMA_OK;
  SOFIMA_CUDA(ctx, cudaFuncSetAttribute(rowspec_tc_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int ncombos = nd * nchunks;
  const int grid = (ctx->num_sms / ncombos) * ncombos;
  {
    LaunchTimer timer(ctx, "flow_rowspec");
    rowspec_tc_kernel<<<grid, kTcThreads, smem, ctx->stream>>>(map, J, out);
    SOFIMA_CHECK_LAUNCH(ctx);
  }
  *done = true;
  return SOFIMA_OK;
}

}  // namespace flow
}  // namespace sofima

extern "C" {

int sofima_xcorr_rowcache(sofima_ctx* ctx, const sofima_xcorr_params* p, const void* pre_img,
                          const void* post_img, const int32_t* pre_xstarts, int32_t n_pre,
                          const int32_t* post_xstarts, int32_t n_post) {
  using namespace sofima;
  using namespace sofima::flow;
  if (!ctx) return fail(nullptr, SOFIMA_EINVAL, "ctx is NULL");
  ctx->rowcache.valid = false;
  if (!p) return SOFIMA_OK;  // clear only
  int rc = check_params(ctx, p);
  if (rc) return rc;
  if (p->ndim != 2) return fail(ctx, SOFIMA_EUNSUPPORTED, "row cache: 2-d patches only");
  if (!pre_img || !post_img || !pre_xstarts || !post_xstarts || n_pre < 1 || n_post < 1)
    return fail(ctx, SOFIMA_EINVAL, "row cache: NULL / empty argument");
  const int sx = p->pre_patch[1] + p->post_patch[1] - 1;
  const int Lx = next_fast_len(sx);
  const int Ly = next_fast_len(p->pre_patch[0] + p->post_patch[0] - 1);
  int n2x = 0, n2y = 0;
  if (!fast_n2(Lx, &n2x) || !fast_n2(Ly, &n2y))
    return fail(ctx, SOFIMA_EUNSUPPORTED, "row cache: transform length %d x %d is not on the "
                "two-pass path", Ly, Lx);
  DeviceGuard guard(ctx->device);
  const int nkx = Lx / 2 + 1;
  const void* datas[2] = {pre_img, post_img};
  const int64_t* shapes[2] = {p->pre_shape, p->post_shape};
  const int pws[2] = {p->pre_patch[1], p->post_patch[1]};
  const int32_t* xs[2] = {pre_xstarts, post_xstarts};
  const int ns[2] = {n_pre, n_post};
  size_t need = 0;
  const int pitch = (nkx + 7) & ~7;  // rows start 64-byte aligned (TMA: 16-byte strides)
  for (int i = 0; i < 2; ++i) need += (size_t)ns[i] * shapes[i][0] * pitch * sizeof(float2);
  size_t have = 0;
  for (const char* nm : {"flow.rowcache0", "flow.rowcache1"}) {
    auto it = ctx->scratch.find(nm);
    if (it != ctx->scratch.end()) have += it->second.bytes;
  }
  if (need > have) {
    size_t free_b = 0, total_b = 0;
    SOFIMA_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
    if (need - have > free_b / 2)
      return fail(ctx, SOFIMA_ENOMEM, "row cache of %zu MB does not fit", need >> 20);
  }
  FftPlan Fx;
  if ((rc = make_plan(ctx, Lx, &Fx))) return rc;

  sofima_ctx::RowCache c;
  c.dtype = p->img_dtype;
  c.L = Lx;
  c.pitch = pitch;
  for (int i = 0; i < 2; ++i) c.nslots[i] = ns[i];
  std::vector<int> table;
  std::vector<float2> fix((size_t)3 * nkx);
  const bool fix_ready = ctx->rowfix_key[0] == Lx && ctx->rowfix_key[1] == pws[0] &&
                         ctx->rowfix_key[2] == pws[1];
  for (int i = 0; i < 2; ++i) {
    const int h = (int)shapes[i][0], w = (int)shapes[i][1], pw = pws[i];
    c.img[i] = datas[i]; c.h[i] = h; c.w[i] = w; c.pw[i] = pw;
    table.assign((size_t)w, -1);
    for (int j = 0; j < ns[i]; ++j) {
      if (xs[i][j] < 0 || xs[i][j] > w - pw)
        return fail(ctx, SOFIMA_EINVAL, "row cache: x start %d out of range", xs[i][j]);
      table[xs[i][j]] = j;
    }
    void *tb = nullptr, *xb = nullptr, *sb = nullptr;
    const char* names[3][2] = {{"flow.rc_xindex0", "flow.rc_xindex1"},
                               {"flow.rc_xstarts0", "flow.rc_xstarts1"},
                               {"flow.rowcache0", "flow.rowcache1"}};
    if ((rc = scratch(ctx, names[0][i], sizeof(int) * w, &tb))) return rc;
    if ((rc = scratch(ctx, names[1][i], sizeof(int) * ns[i], &xb))) return rc;
    if ((rc = scratch(ctx, names[2][i], (size_t)ns[i] * h * pitch * sizeof(float2), &sb))) return rc;
    if ((rc = upload_ints(ctx, xs[i], ns[i], static_cast<int*>(xb)))) return rc;
    xindex_kernel<<<ceil_div(w, 256), 256, 0, ctx->stream>>>(static_cast<const int*>(xb), ns[i],
                                                             static_cast<int*>(tb), w);
    SOFIMA_CHECK_LAUNCH(ctx);
    c.xindex[i] = static_cast<const int*>(tb);
    c.spec[i] = static_cast<float2*>(sb);
    RowSpecJob J;
    J.data = datas[i]; J.dtype = p->img_dtype; J.h = h; J.w = w; J.pw = pw;
    J.xstarts = static_cast<const int*>(xb);
    J.pitch = pitch;
    bool tc_done = false;
    if ((rc = rowspec_tc(ctx, i, datas[i], p->img_dtype, h, w, pw, Lx, xs[i], J.xstarts, ns[i],
                         pitch, c.spec[i], &tc_done)))
      return rc;
    if (!tc_done) {
      LaunchTimer timer(ctx, "flow_rowspec");
#define CALL(N) launch_rowspec_fast<N>(ctx, J, ns[i], Fx.tw, c.spec[i])
      SOFIMA_N2_SWITCH(n2x, CALL)
#undef CALL
      SOFIMA_CHECK_LAUNCH(ctx);
    }
    // FFT of the pw-wide rect, and for the flipped post patch W[k] = exp(-2 pi i k (pw-1) / L)
    for (int k = 0; k < nkx && !fix_ready; ++k) {
      double re = 0.0, im = 0.0;
      for (int n = 0; n < pw; ++n) {
        const double ph = -2.0 * M_PI * (double)((long long)k * n % Lx) / Lx;
        re += cos(ph);
        im += sin(ph);
      }
      fix[(size_t)i * nkx + k] = make_float2((float)re, (float)im);
      if (i == 1) {
        const double ph = -2.0 * M_PI * (double)((long long)k * (pw - 1) % Lx) / Lx;
        fix[(size_t)2 * nkx + k] = make_float2((float)cos(ph), (float)sin(ph));
      }
    }
  }
  void* fb = nullptr;
  if ((rc = scratch(ctx, "flow.rc_fix", sizeof(float2) * fix.size(), &fb))) return rc;
  if (!fix_ready) {
    SOFIMA_CUDA(ctx, cudaMemcpyAsync(fb, fix.data(), sizeof(float2) * fix.size(),
                                     cudaMemcpyHostToDevice, ctx->stream));
    SOFIMA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->rowfix_key[0] = Lx; ctx->rowfix_key[1] = pws[0]; ctx->rowfix_key[2] = pws[1];
  }
  c.fix = static_cast<const float2*>(fb);
  c.valid = true;
  ctx->rowcache = c;
  return SOFIMA_OK;
}

int sofima_xcorr_images(sofima_ctx* ctx, const sofima_xcorr_params* p, const void* pre_img,
                        const void* post_img, const uint8_t* pre_mask,
                        const uint8_t* post_mask, const int32_t* pre_starts,
                        const int32_t* post_starts, int64_t batch, float* out_xcorr) {
  using namespace sofima;
  if (!ctx) return fail(nullptr, SOFIMA_EINVAL, "ctx is NULL");
  int rc = flow::check_params(ctx, p);
  if (rc) return rc;
  if (batch < 0) return fail(ctx, SOFIMA_EINVAL, "batch < 0");
  if (batch == 0) return SOFIMA_OK;
  if (!pre_img || !post_img || !pre_starts || !post_starts || !out_xcorr)
    return fail(ctx, SOFIMA_EINVAL, "NULL array argument");
  DeviceGuard guard(ctx->device);
  if (p->ndim == 3)
    return flow::run_xcorr3(ctx, p, pre_img, post_img, pre_mask, post_mask, pre_starts,
                            post_starts, batch, out_xcorr);
  return flow::run_xcorr(ctx, p, pre_img, post_img, pre_mask, post_mask, pre_starts,
                         post_starts, batch, out_xcorr, nullptr);
}

int sofima_xcorr_peaks(sofima_ctx* ctx, const sofima_xcorr_params* p, const void* pre_img,
                       const void* post_img, const uint8_t* pre_mask, const uint8_t* post_mask,
                       const int32_t* pre_starts, const int32_t* post_starts, int64_t batch,
                       float* out_peaks) {
  using namespace sofima;
  if (!ctx) return fail(nullptr, SOFIMA_EINVAL, "ctx is NULL");
  int rc = flow::check_params(ctx, p);
  if (rc) return rc;
  if (batch < 0) return fail(ctx, SOFIMA_EINVAL, "batch < 0");
  if (batch == 0) return SOFIMA_OK;
  if (!pre_img || !post_img || !pre_starts || !post_starts || !out_peaks)
    return fail(ctx, SOFIMA_EINVAL, "NULL array argument");
  DeviceGuard guard(ctx->device);
  flow::PeakParams pp;
  flow::peak_params(p, &pp);
  if (p->ndim == 2 && !pre_mask && !post_mask) {
    bool handled = false;
    rc = flow::run_fused(ctx, p, pre_img, post_img, pre_starts, post_starts, batch, pp,
                         out_peaks, &handled);
    if (rc || handled) return rc;
  }
  const size_t img_elems = (size_t)pp.sz * pp.sy * pp.sx;
  void* images = nullptr;
  if ((rc = scratch(ctx, "flow.images", sizeof(float) * (size_t)batch * img_elems, &images)))
    return rc;
  bool keys_ready = false;
  int band_rows = 0, nbands = 0;
  if (p->ndim == 3)
    rc = flow::run_xcorr3(ctx, p, pre_img, post_img, pre_mask, post_mask, pre_starts,
                          post_starts, batch, static_cast<float*>(images));
  else
    rc = flow::run_xcorr(ctx, p, pre_img, post_img, pre_mask, post_mask, pre_starts,
                         post_starts, batch, static_cast<float*>(images), &keys_ready,
                         &band_rows, &nbands);
  if (rc) return rc;
  return flow::run_peaks(ctx, static_cast<const float*>(images), batch, pp, out_peaks,
                         keys_ready, band_rows, nbands);
}

int sofima_batched_peaks(sofima_ctx* ctx, int ndim, const float* img, const int64_t* img_shape,
                         int64_t batch, const int32_t* center_offset, int min_distance,
                         float threshold_rel, const int32_t* peak_radius, float* out_peaks) {
  using namespace sofima;
  if (!ctx) return fail(nullptr, SOFIMA_EINVAL, "ctx is NULL");
  if (ndim != 2 && ndim != 3) return fail(ctx, SOFIMA_EINVAL, "ndim must be 2 or 3");
  if (batch < 0) return fail(ctx, SOFIMA_EINVAL, "batch < 0");
  if (batch == 0) return SOFIMA_OK;
  if (!img || !img_shape || !center_offset || !peak_radius || !out_peaks)
    return fail(ctx, SOFIMA_EINVAL, "NULL array argument");
  if (min_distance < 0) return fail(ctx, SOFIMA_EINVAL, "min_distance < 0");
  DeviceGuard guard(ctx->device);
  flow::PeakParams pp;
  memset(&pp, 0, sizeof(pp));
  const int o = 3 - ndim;
  int sh[3] = {1, 1, 1}, r[3] = {0, 0, 0}, c[3] = {0, 0, 0};
  for (int d = 0; d < ndim; ++d) {
    if (img_shape[d] > INT32_MAX) return fail(ctx, SOFIMA_EINVAL, "image too large");
    sh[o + d] = (int)img_shape[d];
    r[o + d] = peak_radius[d];
    c[o + d] = center_offset[d];
  }
  pp.ndim = ndim;
  pp.sz = sh[0]; pp.sy = sh[1]; pp.sx = sh[2];
  pp.rz = r[0]; pp.ry = r[1]; pp.rx = r[2];
  pp.cz = c[0]; pp.cy = c[1]; pp.cx = c[2];
  pp.md = min_distance;
  pp.thr_rel = threshold_rel;
  return flow::run_peaks(ctx, img, batch, pp, out_peaks, false);
}

}  // extern "C"
