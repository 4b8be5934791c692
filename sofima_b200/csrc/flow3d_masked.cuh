// Masked 3-d patch correlation (reference flow_field.py:91-155 with dim = 3): Padfield's
// masked normalised cross-correlation on the zero-padded complex volumes of flow3d.cuh.
// Six forward transforms per pair (patch, valid mask and squared patch of both images), six
// products, six inverse transforms; the Padfield terms and the normalisation with its two
// batch-global maxima are the dimension-agnostic kernels of the 2-d path
// (padfield_terms_kernel / padfield_normalise_kernel).
//
// Checked on a B200 against golden vectors of the reference's NumPy branch
// (tests/test_masked3d_gpu.py).
#pragma once

namespace sofima {
namespace flow {

struct Vol3M {
  const void* data;
  const uint8_t* mask;  // may be NULL: every voxel valid
  int d, h, w;          // image extent (z, y, x)
  int pd, ph, pw;       // patch extent
  int md, mh, mw;       // mask extent (masks may be larger than the image)
};

struct Problem3M {
  Vol3M img[2];
  int dtype;
  const int32_t* starts[2];  // [B][3] (z, y, x)
  int has_mean;
  float mean;
  int Lz, Ly, Lx;
  int sz, sy, sx;
  long long b0;
  int nb;
};

// Mean over the unmasked voxels of one patch (flow_field.py:340-347: nanmean of the patch
// with masked voxels set to NaN); grid = (pair, image), fp64 accumulation in a fixed order.
__global__ void __launch_bounds__(kThreads)
patch_mean3m_kernel(Problem3M P, float* means) {
  const int which = blockIdx.y;
  const long long b = P.b0 + blockIdx.x;
  if (P.has_mean) {
    if (threadIdx.x == 0) means[b * 2 + which] = P.mean;
    return;
  }
  const Vol3M& I = P.img[which];
  const int z0 = clamp_start(P.starts[which][b * 3 + 0], I.pd, I.d);
  const int y0 = clamp_start(P.starts[which][b * 3 + 1], I.ph, I.h);
  const int x0 = clamp_start(P.starts[which][b * 3 + 2], I.pw, I.w);
  int mz0 = 0, my0 = 0, mx0 = 0;
  if (I.mask) {
    mz0 = clamp_start(P.starts[which][b * 3 + 0], I.pd, I.md);
    my0 = clamp_start(P.starts[which][b * 3 + 1], I.ph, I.mh);
    mx0 = clamp_start(P.starts[which][b * 3 + 2], I.pw, I.mw);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double sum = 0.0;
  int cnt = 0;
  for (int r = warp; r < I.pd * I.ph; r += kThreads / 32) {
    const int z = r / I.ph, y = r - z * I.ph;
    const long long row = ((long long)(z0 + z) * I.h + (y0 + y)) * I.w + x0;
    const uint8_t* mrow =
        I.mask ? I.mask + ((long long)(mz0 + z) * I.mh + (my0 + y)) * I.mw + mx0 : nullptr;
    for (int x = lane; x < I.pw; x += 32) {
      const bool valid = mrow ? (mrow[x] == 0) : true;
      sum += valid ? (double)load_px(I.data, P.dtype, row + x) : 0.0;
      cnt += valid;
    }
  }
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
    double t = 0.0;
    int c = 0;
    for (int w = 0; w < kThreads / 32; ++w) { t += rs[w]; c += rc[w]; }
    means[b * 2 + which] = __fdiv_rn((float)t, (float)c);  // 0 / 0 = NaN: patch fully masked
  }
}

// Z[slot][pair] zero-padded volumes; slot = 2 * kind + which with which = 0 (pre), 1 (post,
// flipped on all axes) and kind = 0: (patch - mean) with masked voxels zeroed
// (flow_field.py:73-76), 1: valid-mask indicator, 2: square of kind 0.
// grid = (blocks, 6, pair).
__global__ void __launch_bounds__(kThreads)
pack3m_kernel(Problem3M P, const float* __restrict__ means, float2* __restrict__ Z) {
  const int slot = blockIdx.y;
  const int which = slot & 1, kind = slot >> 1;
  const long long b = P.b0 + blockIdx.z;
  const Vol3M& I = P.img[which];
  const int z0 = clamp_start(P.starts[which][b * 3 + 0], I.pd, I.d);
  const int y0 = clamp_start(P.starts[which][b * 3 + 1], I.ph, I.h);
  const int x0 = clamp_start(P.starts[which][b * 3 + 2], I.pw, I.w);
  int mz0 = 0, my0 = 0, mx0 = 0;
  if (I.mask) {
    mz0 = clamp_start(P.starts[which][b * 3 + 0], I.pd, I.md);
    my0 = clamp_start(P.starts[which][b * 3 + 1], I.ph, I.mh);
    mx0 = clamp_start(P.starts[which][b * 3 + 2], I.pw, I.mw);
  }
  const float mean = means[b * 2 + which];
  const bool flip = which == 1;
  const long long vol = (long long)P.Lz * P.Ly * P.Lx;
  float2* out = Z + ((long long)slot * P.nb + blockIdx.z) * vol;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < vol;
       i += (long long)gridDim.x * kThreads) {
    const int x = (int)(i % P.Lx);
    const long long r = i / P.Lx;
    const int y = (int)(r % P.Ly), z = (int)(r / P.Ly);
    float v = 0.f;
    if (z < I.pd && y < I.ph && x < I.pw) {
      const int zz = flip ? I.pd - 1 - z : z, yy = flip ? I.ph - 1 - y : y,
                xx = flip ? I.pw - 1 - x : x;
      const bool valid =
          I.mask ? I.mask[((long long)(mz0 + zz) * I.mh + (my0 + yy)) * I.mw + mx0 + xx] == 0
                 : true;
      if (kind == 1) {
        v = valid ? 1.f : 0.f;
      } else {
        v = load_px(I.data, P.dtype,
                    ((long long)(z0 + zz) * I.h + (y0 + yy)) * I.w + x0 + xx) - mean;
        v = valid ? v : 0.f;
        if (kind == 2) v = v * v;
      }
    }
    out[i] = make_float2(v, 0.f);
  }
}

// The six spectra products of flow_field.py:81-129.  Z slots: 0 P, 1 C, 2 MP, 3 MC, 4 P2,
// 5 C2 (each [nb][vol]); W outputs: 0 P C (numerator), 1 MC MP (overlap), 2 MC P (mc_p),
// 3 MP C (mc_c), 4 MC P2 (p_sq), 5 MP C2 (c_sq).
__global__ void __launch_bounds__(kThreads)
multiply3m_kernel(const float2* __restrict__ Z, float2* __restrict__ W, long long nvol) {
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < nvol;
       i += (long long)gridDim.x * kThreads) {
    const float2 p = Z[i], c = Z[nvol + i], mp = Z[2 * nvol + i], mc = Z[3 * nvol + i];
    const float2 p2 = Z[4 * nvol + i], c2 = Z[5 * nvol + i];
    W[i] = cmul(p, c);
    W[nvol + i] = cmul(mc, mp);
    W[2 * nvol + i] = cmul(mc, p);
    W[3 * nvol + i] = cmul(mp, c);
    W[4 * nvol + i] = cmul(mc, p2);
    W[5 * nvol + i] = cmul(mp, c2);
  }
}

struct Outputs3M {
  float* dst[6];  // per output j: base pointer such that pair index q addresses dst[j] + q * n
  long long first[6];  // pair index of the first pair of this sub-batch in dst[j]'s numbering
};

// Crop + scale of the six inverse transforms; grid = (blocks, 6, pair).
__global__ void __launch_bounds__(kThreads)
crop3m_kernel(Problem3M P, const float2* __restrict__ W, Outputs3M outs, float scale) {
  const int j = blockIdx.y;
  const long long vol = (long long)P.Lz * P.Ly * P.Lx;
  const long long n = (long long)P.sz * P.sy * P.sx;
  const float2* in = W + ((long long)j * P.nb + blockIdx.z) * vol;
  float* out = outs.dst[j] + (outs.first[j] + blockIdx.z) * n;
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
