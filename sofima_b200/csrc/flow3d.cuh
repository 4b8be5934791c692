// 3-d patch correlation (reference flow_field.py:36-89 with dim = 3, unmasked).
//
// Every patch is packed into a zero-padded complex volume [Lz][Ly][Lx], transformed along
// x, y, z (forward passes skip the all-zero lines), multiplied, transformed back and
// cropped.  Lengths of the form 16 * N2 use the register codelets of the 2-d fast path
// (axis_fft_fast_kernel: 16 neighbouring lines per block, coalesced strided passes); other
// 5-smooth lengths the generic shared-memory Stockham pass (axis_fft_kernel).  The peak
// kernels of flow.cu are dimension-generic.  BASELINE config 5.
#pragma once

namespace sofima {
namespace flow {

struct Vol3 {
  const void* data;
  int d, h, w;     // image extent (z, y, x)
  int pd, ph, pw;  // patch extent
};

struct Problem3 {
  Vol3 img[2];
  int dtype;
  const int32_t* starts[2];  // [B][3] (z, y, x)
  int has_mean;
  float mean;
  int Lz, Ly, Lx;
  int sz, sy, sx;
  long long b0;
  int nb;
};

// Sum of one 3-d patch in kMeanSlices partial sums; grid = (pair, image, slice), deterministic
// fp64 accumulation (one block per patch took 0.8 ms per sub-batch: 14 blocks on 148 SMs).
constexpr int kMeanSlices = 16;

__global__ void __launch_bounds__(kThreads)
patch_mean3_kernel(Problem3 P, double* parts) {
  const int which = blockIdx.y;
  const long long b = P.b0 + blockIdx.x;
  if (P.has_mean) return;
  const Vol3& I = P.img[which];
  const int z0 = clamp_start(P.starts[which][b * 3 + 0], I.pd, I.d);
  const int y0 = clamp_start(P.starts[which][b * 3 + 1], I.ph, I.h);
  const int x0 = clamp_start(P.starts[which][b * 3 + 2], I.pw, I.w);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rows = I.pd * I.ph;
  const int r0 = (int)((long long)rows * blockIdx.z / kMeanSlices);
  const int r1 = (int)((long long)rows * (blockIdx.z + 1) / kMeanSlices);
  double sum = 0.0;
  for (int r = r0 + warp; r < r1; r += kThreads / 32) {
    const int z = r / I.ph, y = r - z * I.ph;
    const long long row = ((long long)(z0 + z) * I.h + (y0 + y)) * I.w + x0;
    for (int x = lane; x < I.pw; x += 32) sum += (double)load_px(I.data, P.dtype, row + x);
  }
  __shared__ double rs[kThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) rs[warp] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) t += rs[w];
    parts[(b * 2 + which) * kMeanSlices + blockIdx.z] = t;
  }
}

__device__ __forceinline__ float patch_mean3(const Problem3& P, const double* parts, long long b,
                                             int which) {
  if (P.has_mean) return P.mean;
  const Vol3& I = P.img[which];
  double t = 0.0;
  for (int s = 0; s < kMeanSlices; ++s) t += parts[(b * 2 + which) * kMeanSlices + s];
  return __fdiv_rn((float)t, (float)(I.pd * I.ph * I.pw));
}

// Z[slot][pair] = zero-padded (patch - mean), the post patch flipped on all axes.
__global__ void __launch_bounds__(kThreads)
pack3_kernel(Problem3 P, const double* __restrict__ parts, float2* __restrict__ Z) {
  const int which = blockIdx.y;
  const long long b = P.b0 + blockIdx.z;
  const Vol3& I = P.img[which];
  const int z0 = clamp_start(P.starts[which][b * 3 + 0], I.pd, I.d);
  const int y0 = clamp_start(P.starts[which][b * 3 + 1], I.ph, I.h);
  const int x0 = clamp_start(P.starts[which][b * 3 + 2], I.pw, I.w);
  const float mean = patch_mean3(P, parts, b, which);
  const bool flip = which == 1;
  const long long vol = (long long)P.Lz * P.Ly * P.Lx;
  float2* out = Z + ((long long)which * P.nb + blockIdx.z) * vol;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < vol;
       i += (long long)gridDim.x * kThreads) {
    const int x = (int)(i % P.Lx);
    const long long r = i / P.Lx;
    const int y = (int)(r % P.Ly), z = (int)(r / P.Ly);
    float v = 0.f;
    if (z < I.pd && y < I.ph && x < I.pw) {
      const int zz = flip ? I.pd - 1 - z : z, yy = flip ? I.ph - 1 - y : y,
                xx = flip ? I.pw - 1 - x : x;
      v = load_px(I.data, P.dtype, ((long long)(z0 + zz) * I.h + (y0 + yy)) * I.w + x0 + xx) -
          mean;
    }
    out[i] = make_float2(v, 0.f);
  }
}

// In-place complex FFT of `nlines` strided lines of length F.L:
//   line l starts at (l / inner) * outer_stride + (l % inner) * inner_stride,
//   consecutive elements are `es` apart.  C lines per block.
// Lines of a pass are numbered (volume, a, b) with a < An, b < Bn; a pass that only needs the
// lines with a < Au and b < Bu (forward passes over zero-padded volumes: every other line is
// all zeros and stays all zeros) numbers exactly those.  An = Au = Bn = Bu = 1 means "all".
struct LinePrune {
  long long An, Au, Bn, Bu;
};

template <bool INV>
__global__ void __launch_bounds__(kThreads)
axis_fft_kernel(float2* __restrict__ data, long long nlines, long long inner,
                long long inner_stride, long long outer_stride, long long es, FftPlan F,
                int C, LinePrune pr) {
  extern __shared__ float2 smem[];
  __shared__ long long base_s[16];
  const int L = F.L;
  const int LP = L | 1;  // odd line pitch: the 16 lines of a block fall on 16 different banks
  float2* w0 = smem;
  float2* w1 = w0 + (size_t)C * LP;
  float2* tw_s = w1 + (size_t)C * LP;
  load_twiddles(tw_s, F);
  const long long l0 = (long long)blockIdx.x * C;
  const int nc = (int)min((long long)C, nlines - l0);
  if ((int)threadIdx.x < nc) {  // one address computation per line, not per element
    const long long lp = l0 + threadIdx.x;
    const long long per = pr.Au * pr.Bu;
    const long long v = lp / per, r = lp - v * per;
    const long long a = r / pr.Bu, b = r - a * pr.Bu;
    const long long l = (v * pr.An + a) * pr.Bn + b;
    base_s[threadIdx.x] = (l / inner) * outer_stride + (l % inner) * inner_stride;
  }
  __syncthreads();
  if (es == 1) {
    for (int i = threadIdx.x; i < C * L; i += kThreads) {
      const int c = i / L, k = i - c * L;
      w0[c * LP + k] = c < nc ? data[base_s[c] + k] : make_float2(0.f, 0.f);
    }
  } else {
    for (int i = threadIdx.x; i < C * L; i += kThreads) {
      const int k = i / C, c = i - k * C;  // c fastest: neighbouring lines are adjacent
      w0[c * LP + k] = c < nc ? data[base_s[c] + (long long)k * es] : make_float2(0.f, 0.f);
    }
  }
  __syncthreads();
  const float2* res = block_fft<INV>(w0, w1, C, F, tw_s, LP);
  if (es == 1) {
    for (int i = threadIdx.x; i < C * L; i += kThreads) {
      const int c = i / L, k = i - c * L;
      if (c < nc) data[base_s[c] + k] = res[c * LP + k];
    }
  } else {
    for (int i = threadIdx.x; i < C * L; i += kThreads) {
      const int k = i / C, c = i - k * C;
      if (c < nc) data[base_s[c] + (long long)k * es] = res[c * LP + k];
    }
  }
}

// The same pass for lengths L = 16 * N2 with all butterflies in registers (the codelets of the
// 2-d fast path): thread (line c, r) runs a 16-point DFT over the elements N2 n1 + r, the
// results are twiddled and exchanged through shared memory, thread (c, k1) runs the N2-point
// DFT and owns the bins k1 + 16 k2.  16 neighbouring lines per block (c fastest: strided
// passes read and write 128-byte runs); contiguous lines (es == 1) are staged through shared
// memory so that their global accesses are coalesced too.  The inverse is the forward
// transform with real and imaginary parts swapped on the way in and out.
template <int N2>
struct AxisFast {
  static constexpr int L = kN1 * N2;
  static constexpr int C = 16;
  static constexpr int G = N2 > kN1 ? N2 : kN1;
  static constexpr int NT = C * G;
  static constexpr int N2P = N2 | 1;
  static constexpr int EX = kN1 * N2P + 1;  // odd: the 16 lines fall on 16 different banks
  static constexpr int XP = L + 1;
  static constexpr size_t smem = sizeof(float2) * ((size_t)C * EX + (size_t)C * XP + L);
};

template <int N2, bool INV>
__global__ void __launch_bounds__(AxisFast<N2>::NT)
axis_fft_fast_kernel(float2* __restrict__ data, long long nlines, long long inner,
                     long long inner_stride, long long outer_stride, long long es,
                     const float2* __restrict__ tw, LinePrune pr) {
  using D = AxisFast<N2>;
  constexpr int L = D::L, C = D::C;
  extern __shared__ float2 smem[];
  __shared__ long long base_s[C];
  float2* ex = smem;
  float2* xs = ex + C * D::EX;
  float2* tw_s = xs + C * D::XP;
  stage_twiddles<L, D::NT>(tw_s, tw);
  const long long l0 = (long long)blockIdx.x * C;
  const int nc = (int)min((long long)C, nlines - l0);
  if ((int)threadIdx.x < nc) {
    const long long lp = l0 + threadIdx.x;
    const long long per = pr.Au * pr.Bu;
    const long long v = lp / per, rr = lp - v * per;
    const long long a = rr / pr.Bu, b = rr - a * pr.Bu;
    const long long l = (v * pr.An + a) * pr.Bn + b;
    base_s[threadIdx.x] = (l / inner) * outer_stride + (l % inner) * inner_stride;
  }
  __syncthreads();
  const int c = threadIdx.x % C, r = threadIdx.x / C;
  const bool contiguous = es == 1;
  if (contiguous) {
    for (int i = threadIdx.x; i < C * L; i += D::NT) {
      const int cc = i / L, k = i - cc * L;
      xs[cc * D::XP + k] = cc < nc ? data[base_s[cc] + k] : make_float2(0.f, 0.f);
    }
    __syncthreads();
  }
  float2 a[kN1];
  if (r < N2) {
#pragma unroll
    for (int n1 = 0; n1 < kN1; ++n1) {
      const int k = N2 * n1 + r;
      float2 v;
      if (contiguous) v = xs[c * D::XP + k];
      else v = c < nc ? data[base_s[c] + (long long)k * es] : make_float2(0.f, 0.f);
      a[n1] = INV ? swap_ri(v) : v;
    }
    Dft<kN1>::run(a);
  }
  stage_twiddles_wait();
  __syncthreads();  // twiddles staged
  if (r < N2) {
#pragma unroll
    for (int k1 = 0; k1 < kN1; ++k1) ex[c * D::EX + k1 * D::N2P + r] = cmul(a[k1], tw_s[r * k1]);
  }
  __syncthreads();
  if (r < kN1) {
    float2 bq[N2];
#pragma unroll
    for (int n2 = 0; n2 < N2; ++n2) bq[n2] = ex[c * D::EX + r * D::N2P + n2];
    Dft<N2>::run(bq);
#pragma unroll
    for (int k2 = 0; k2 < N2; ++k2) {
      const int k = r + kN1 * k2;
      const float2 v = INV ? swap_ri(bq[k2]) : bq[k2];
      if (contiguous) xs[c * D::XP + k] = v;
      else if (c < nc) data[base_s[c] + (long long)k * es] = v;
    }
  }
  if (contiguous) {
    __syncthreads();
    for (int i = threadIdx.x; i < C * L; i += D::NT) {
      const int cc = i / L, k = i - cc * L;
      if (cc < nc) data[base_s[cc] + k] = xs[cc * D::XP + k];
    }
  }
}

__global__ void __launch_bounds__(kThreads)
multiply3_kernel(float2* __restrict__ a, const float2* __restrict__ b, long long n) {
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n;
       i += (long long)gridDim.x * kThreads)
    a[i] = cmul(a[i], b[i]);
}

__global__ void __launch_bounds__(kThreads)
crop3_kernel(Problem3 P, const float2* __restrict__ Z, float* __restrict__ images,
             float scale) {
  const long long vol = (long long)P.Lz * P.Ly * P.Lx;
  const long long n = (long long)P.sz * P.sy * P.sx;
  const float2* in = Z + (long long)blockIdx.y * vol;
  float* out = images + (P.b0 + blockIdx.y) * n;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n;
       i += (long long)gridDim.x * kThreads) {
    const int x = (int)(i % P.sx);
    const long long r = i / P.sx;
    const int y = (int)(r % P.sy), z = (int)(r / P.sy);
    out[i] = in[((long long)z * P.Ly + y) * P.Lx + x].x * scale;
  }
}

}  // namespace flow
}  // namespace sofima
