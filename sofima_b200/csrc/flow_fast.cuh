// Two-pass register-resident FFT kernels for the unmasked correlation path.
//
// For transform lengths L = 16 * N2 (N2 in {8,10,12,15,16,20,24,25,32}; L = 320 for
// the EM default patch 160) each 1-d FFT is done in exactly two passes with all
// butterflies in registers (generated codelets on packed fp32x2 instructions,
// fft_codelets.cuh):
//
//   forward  x[N2 n1 + n2] -> X[k1 + 16 k2]:
//     pass 1  thread n2 : 16-point DFT over n1, twiddle W_L^(n2 k1)
//     (exchange through shared memory, conflict-free padded layout)
//     pass 2  thread k1 : N2-point DFT over n2
//   inverse  X[k1 + 16 k2] -> x[N2 n1 + n2]  (re/im swap trick, same codelets):
//     pass 1  thread k1 : N2-point DFT over k2, twiddle W_L^(k1 n2)
//     pass 2  thread n2 : 16-point DFT over k1
//
// The three stages are those of flow.cu (rows_fwd / cols / rows_inv); the cols
// stage keeps the spectrum of the first patch in shared memory, multiplies in
// registers and runs the inverse without leaving the SM.  rows_inv also folds the
// first peak search (global max + first argmax per image) into its epilogue.
#pragma once

#include "fft_codelets.cuh"

namespace sofima {
namespace flow {

constexpr int kN1 = 16;

template <int N2>
struct FastDims {
  static constexpr int L = kN1 * N2;
  static constexpr int N2P = N2 | 1;           // odd pitch: conflict-free exchange
  static constexpr int EX = kN1 * N2P + 2;     // float2 per exchange line (== 2 mod 16)
  // Row kernels map threads transform-major (groups of N2 consecutive threads): the
  // lines of consecutive transforms must continue the bank sequence, i.e. the line
  // pitch is == N2 (mod 16) in float2 units ...
  static constexpr int EXR = kN1 * N2P + ((N2 - kN1 * N2P) % 16 + 16) % 16;
  // ... and two staged pixel rows (floats) advance the banks by N2 (mod 32).
  static constexpr int PXP = L + (N2 / 2) % 16;
  static constexpr int LP = L + 2;             // padded spectrum line (== 2 mod 16)
  static constexpr int G = N2 > kN1 ? N2 : kN1;  // threads per transform
};

__device__ __forceinline__ float2 swap_ri(float2 a) { return make_float2(a.y, a.x); }

// Twiddle table -> shared memory with asynchronous copies (LDGSTS): a plain load ->
// shared-store pair at the top of a kernel stalls the thread for one L2 round trip before
// it can request its spectra from HBM; the asynchronous copy needs no register and no
// scoreboard wait.  stage_twiddles_wait() + __syncthreads() must precede the first use.
template <int L, int NT>
__device__ __forceinline__ void stage_twiddles(float2* tw_s, const float2* __restrict__ tw) {
  for (int i = threadIdx.x; i < L; i += NT) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(tw_s + i);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(tw + i) : "memory");
  }
}
// The same table as the inverse-row kernel reads it -- thread (f, k1) needs W^(k1 n2) for all
// n2 -- transposed to [n2][k1]: the 16 lanes of a half-warp then read 16 consecutive entries
// instead of a stride of n2 (a 16-way bank conflict for n2 = 16).
template <int N2, int NT>
__device__ __forceinline__ void stage_twiddles_by_n2(float2* tw_s, const float2* __restrict__ tw) {
  for (int i = threadIdx.x; i < N2 * kN1; i += NT) {
    const int n2 = i / kN1, k1 = i - n2 * kN1;
    const unsigned dst = (unsigned)__cvta_generic_to_shared(tw_s + i);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(tw + k1 * n2)
                 : "memory");
  }
}
__device__ __forceinline__ void stage_twiddles_wait() {
  asm volatile("cp.async.wait_all;" ::: "memory");
}

// ---------------------------------------------------------------------------------
// Stage 1: forward row FFTs (two real rows per complex transform).
// grid = (row-pair groups, slot, pair), block = TR * N2 threads.
// ---------------------------------------------------------------------------------
template <int N2, int TR, bool HALF>
__global__ void __launch_bounds__(TR * FastDims<N2>::G)
rows_fwd_fast(Problem P, const float2* __restrict__ tw, float2* __restrict__ T) {
  using D = FastDims<N2>;
  constexpr int L = D::L;
  constexpr int NT = TR * D::G;
  constexpr int NKX = L / 2 + 1;
  __shared__ float2 ex[TR * D::EXR];
  __shared__ float2 xs[TR * D::PXP];   // first the staged pixel rows (as float), then X
  __shared__ float2 tw_s[L];
  float* px = reinterpret_cast<float*>(xs);  // [2 TR][PXP]
  const Slot sl = P.slot[blockIdx.y];
  const Image& I = P.img[sl.src];
  const long long b = P.b0 + blockIdx.z;
  const int rp0 = blockIdx.x * TR;
  const int nrows = I.ph, pw = I.pw;
  if (2 * rp0 >= nrows) return;
  stage_twiddles<L, NT>(tw_s, tw);

  const int y0 = clamp_start(P.starts[sl.src][b * 2 + 0], I.ph, I.h);
  const int x0 = clamp_start(P.starts[sl.src][b * 2 + 1], I.pw, I.w);
  int my0 = 0, mx0 = 0;
  if (I.mask) {
    my0 = clamp_start(P.starts[sl.src][b * 2 + 0], I.ph, I.mh);
    mx0 = clamp_start(P.starts[sl.src][b * 2 + 1], I.pw, I.mw);
  }
  const float mean = patch_mean(P, b, sl.src);
  const bool flip = sl.src == 1;  // curr[::-1, ::-1], flow_field.py:78-79

  // Stage the 2 TR patch rows of this block: coalesced reads, one value per pixel
  // = (pixel - mean) with masked pixels zeroed (flow_field.py:73-76), or the
  // valid-mask indicator / the square for the Padfield terms.
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < 2 * TR; r += NT / 32) {
      const int y = 2 * rp0 + r;
      float* dst = px + r * D::PXP;
      if (y >= nrows) {
        for (int x = lane; x < pw; x += 32) dst[x] = 0.f;
        continue;
      }
      const int yy = flip ? nrows - 1 - y : y;
      const long long row = (long long)(y0 + yy) * I.w + x0;
      const uint8_t* mrow = I.mask ? I.mask + (long long)(my0 + yy) * I.mw + mx0 : nullptr;
      for (int x = lane; x < pw; x += 32) {
        const int xx = flip ? pw - 1 - x : x;
        const bool valid = mrow ? (mrow[xx] == 0) : true;
        float v;
        if (sl.xform == 1) {
          v = valid ? 1.f : 0.f;
        } else {
          v = load_px(I.data, P.dtype, row + xx) - mean;
          v = valid ? v : 0.f;
          if (sl.xform == 2) v = v * v;
        }
        dst[x] = v;
      }
    }
  }
  stage_twiddles_wait();
  __syncthreads();

  if (threadIdx.x < TR * N2) {  // pass 1: thread (f, n2)
    const int f = threadIdx.x / N2, n2 = threadIdx.x - f * N2;
    const float* r0 = px + (2 * f) * D::PXP;
    const float* r1 = r0 + D::PXP;
    float2 a[kN1];
#pragma unroll
    for (int n1 = 0; n1 < kN1; ++n1) {
      const int x = N2 * n1 + n2;
      if (HALF && n1 >= kN1 / 2) {
        a[n1] = make_float2(0.f, 0.f);  // pw <= L / 2: compile-time zeros prune the DFT
      } else {
        a[n1] = (x < pw) ? make_float2(r0[x], r1[x]) : make_float2(0.f, 0.f);
      }
    }
    Dft<kN1>::run(a);
#pragma unroll
    for (int k1 = 0; k1 < kN1; ++k1)
      ex[f * D::EXR + k1 * D::N2P + n2] = cmul(a[k1], tw_s[n2 * k1]);
  }
  __syncthreads();
  if (threadIdx.x < TR * kN1) {  // pass 2: thread (f, k1)
    const int f = threadIdx.x / kN1, k1 = threadIdx.x % kN1;
    float2 bq[N2];
#pragma unroll
    for (int n2 = 0; n2 < N2; ++n2) bq[n2] = ex[f * D::EXR + k1 * D::N2P + n2];
    Dft<N2>::run(bq);
#pragma unroll
    for (int k2 = 0; k2 < N2; ++k2) xs[f * L + k1 + kN1 * k2] = bq[k2];
  }
  __syncthreads();
  // separate the two real rows: X_even = (Z[k] + conj Z[L-k]) / 2,
  //                             X_odd  = (Z[k] - conj Z[L-k]) / (2i)
  float2* Tb = T + ((size_t)blockIdx.y * P.nb + blockIdx.z) * P.PY * NKX;
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int f = warp; f < TR; f += NT / 32) {
      const int y = 2 * (rp0 + f);
      if (y >= nrows) continue;
      const bool two = y + 1 < nrows;
      for (int k = lane; k < NKX; k += 32) {
        const float2 a = xs[f * L + k];
        const float2 c = xs[f * L + (k == 0 ? 0 : L - k)];
        Tb[(size_t)y * NKX + k] = make_float2(0.5f * (a.x + c.x), 0.5f * (a.y - c.y));
        if (two)
          Tb[(size_t)(y + 1) * NKX + k] = make_float2(0.5f * (a.y + c.y), 0.5f * (c.x - a.x));
      }
    }
  }
}

// ---------------------------------------------------------------------------------
// Stage 1, shared form: patches of a regular flow grid overlap, so the same image row
// segment [x0, x0 + pw) appears in ph / step patches.  By linearity
//   FFT(row - mean * rect) = FFT(row) - mean * FFT(rect)
// the forward row transform of the RAW pixels can be computed once per (image row,
// distinct x0) and the per-patch mean applied when the column stage loads it; the
// flipped post patch follows from FFT(reversed real row)[k] = W^k(pw-1) conj(FFT(row)[k]).
// grid = (row-pair groups of the image, x-start slot), block = TR * G threads.
// out: [slot][h][L / 2 + 1].
// ---------------------------------------------------------------------------------
struct RowSpecJob {
  const void* data;
  int dtype, h, w, pw;
  const int* xstarts;  // [slots] device
  int pitch;           // float2 elements per cached row (>= L / 2 + 1, a multiple of 8)
};

template <int N2, int TR, bool HALF>
__global__ void __launch_bounds__(TR * FastDims<N2>::G)
rowspec_fast(RowSpecJob J, const float2* __restrict__ tw, float2* __restrict__ out) {
  using D = FastDims<N2>;
  constexpr int L = D::L;
  constexpr int NT = TR * D::G;
  constexpr int NKX = L / 2 + 1;
  __shared__ float2 ex[TR * D::EXR];
  __shared__ float2 xs[TR * D::PXP];
  __shared__ float2 tw_s[L];
  float* px = reinterpret_cast<float*>(xs);
  const int rp0 = blockIdx.x * TR;
  const int nrows = J.h, pw = J.pw;
  if (2 * rp0 >= nrows) return;
  stage_twiddles<L, NT>(tw_s, tw);
  const int x0 = __ldg(&J.xstarts[blockIdx.y]);
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < 2 * TR; r += NT / 32) {
      const int y = 2 * rp0 + r;
      float* dst = px + r * D::PXP;
      if (y >= nrows) {
        for (int x = lane; x < pw; x += 32) dst[x] = 0.f;
        continue;
      }
      const long long row = (long long)y * J.w + x0;
      for (int x = lane; x < pw; x += 32) dst[x] = load_px(J.data, J.dtype, row + x);
    }
  }
  stage_twiddles_wait();
  __syncthreads();
  if (threadIdx.x < TR * N2) {
    const int f = threadIdx.x / N2, n2 = threadIdx.x - f * N2;
    const float* r0 = px + (2 * f) * D::PXP;
    const float* r1 = r0 + D::PXP;
    float2 a[kN1];
#pragma unroll
    for (int n1 = 0; n1 < kN1; ++n1) {
      const int x = N2 * n1 + n2;
      if (HALF && n1 >= kN1 / 2) {
        a[n1] = make_float2(0.f, 0.f);
      } else {
        a[n1] = (x < pw) ? make_float2(r0[x], r1[x]) : make_float2(0.f, 0.f);
      }
    }
    Dft<kN1>::run(a);
#pragma unroll
    for (int k1 = 0; k1 < kN1; ++k1)
      ex[f * D::EXR + k1 * D::N2P + n2] = cmul(a[k1], tw_s[n2 * k1]);
  }
  __syncthreads();
  if (threadIdx.x < TR * kN1) {
    const int f = threadIdx.x / kN1, k1 = threadIdx.x % kN1;
    float2 bq[N2];
#pragma unroll
    for (int n2 = 0; n2 < N2; ++n2) bq[n2] = ex[f * D::EXR + k1 * D::N2P + n2];
    Dft<N2>::run(bq);
#pragma unroll
    for (int k2 = 0; k2 < N2; ++k2) xs[f * L + k1 + kN1 * k2] = bq[k2];
  }
  __syncthreads();
  const int pitch = J.pitch;
  float2* ob = out + (size_t)blockIdx.y * nrows * pitch;
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int f = warp; f < TR; f += NT / 32) {
      const int y = 2 * (rp0 + f);
      if (y >= nrows) continue;
      const bool two = y + 1 < nrows;
      for (int k = lane; k < NKX; k += 32) {
        const float2 a = xs[f * L + k];
        const float2 c = xs[f * L + (k == 0 ? 0 : L - k)];
        ob[(size_t)y * pitch + k] = make_float2(0.5f * (a.x + c.x), 0.5f * (a.y - c.y));
        if (two)
          ob[(size_t)(y + 1) * pitch + k] = make_float2(0.5f * (a.y + c.y), 0.5f * (c.x - a.x));
      }
    }
  }
}

struct RowCacheView {
  const float2* spec[2];  // [slot][h][pitch]
  const int* xindex[2];
  const float2* fix;  // [3][nkx]
  int h[2];
  int pitch;          // float2 elements per cached row
  const int4* meta;   // [B][2]: {first cached row (lo, hi), mean bits, slot valid}
};

// Per (patch, image) record for the cached column stage, so that the column kernel
// needs ONE load instead of the chain starts -> xindex -> spectra / partial sums.
// One warp per (patch, image).  DC_MEAN: uint8 images -- the patch sum is the sum of the
// DC bins of its cached rows (exact: integer sums below 2^24 survive the fp32 butterflies
// unchanged), so the separate pass over the pixels is not needed.
template <bool DC_MEAN>
__global__ void rowcache_meta_kernel(Problem P, RowCacheView RC, long long B, int4* meta) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= 2 * B) return;
  const long long b = i >> 1;
  const int sl = (int)(i & 1);
  const Image& I = P.img[sl];
  const int y0 = clamp_start(P.starts[sl][b * 2 + 0], I.ph, I.h);
  const int x0 = clamp_start(P.starts[sl][b * 2 + 1], I.pw, I.w);
  const int slot = RC.xindex[sl][x0];
  const long long first = (long long)(slot < 0 ? 0 : slot) * RC.h[sl] + y0;
  float mean;
  if (DC_MEAN) {
    float s = 0.f;
    const float2* col0 = RC.spec[sl] + first * RC.pitch;
    for (int y = lane; y < I.ph; y += 32) s += __ldg(&col0[(long long)y * RC.pitch]).x;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    mean = __fdiv_rn(s, (float)(I.ph * I.pw));
  } else {
    mean = patch_mean(P, b, sl);
  }
  const long long row0 = sl == 0 ? first : first + I.ph - 1;
  if (lane == 0)
    meta[i] = make_int4((int)(row0 & 0xffffffffll), (int)(row0 >> 32), __float_as_int(mean),
                        slot >= 0 ? 1 : 0);
}

// ---------------------------------------------------------------------------------
// Stage 2: forward column FFTs of both patches, product, inverse column FFT.
// grid = (column groups, pair), block = C * N2 threads (c fastest).
// ---------------------------------------------------------------------------------
template <int N2, int C, bool HALF, bool CACHED>
__global__ void __launch_bounds__(C * FastDims<N2>::G)
cols_fast(Problem P, const float2* __restrict__ tw, const float2* __restrict__ T,
          float2* __restrict__ U, const RowCacheView RC) {
  using D = FastDims<N2>;
  constexpr int L = D::L;
  __shared__ float2 ex[C * D::EX];
  __shared__ float2 sa[C * D::LP];
  __shared__ float2 tw_s[L];
  stage_twiddles<L, C * D::G>(tw_s, tw);
  const int k0 = blockIdx.x * C;
  const int c = threadIdx.x % C;
  const int r = threadIdx.x / C;  // n2 in pass-1 role, k1 in pass-2 role (r < 16)
  const bool col_ok = k0 + c < P.nkx;
  float2 prod[N2];
  // cached form: all patch-level loads are issued up front (no dependent chains later)
  int4 meta2[2] = {make_int4(0, 0, 0, 0), make_int4(0, 0, 0, 0)};
  float2 fix3[3] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(1.f, 0.f)};
  if (CACHED) {
    meta2[0] = __ldg(&RC.meta[(P.b0 + blockIdx.y) * 2 + 0]);
    meta2[1] = __ldg(&RC.meta[(P.b0 + blockIdx.y) * 2 + 1]);
    if (col_ok) {
#pragma unroll
      for (int j = 0; j < 3; ++j) fix3[j] = __ldg(&RC.fix[j * P.nkx + k0 + c]);
    }
  }

  // Cached row spectra: row y of the patch is image row y0 + y (pre) resp.
  // y0 + rows - 1 - y (post, flipped), slot = xindex[x0]; the mean and the x flip are
  // applied after the load (see rowspec_fast).  The raw spectra of the post patch are
  // requested while the pre patch is still being transformed (software pipelining: the
  // loads were half of this kernel's stalls).
  constexpr int NLOAD = HALF ? kN1 / 2 : kN1;
  float2 raw[NLOAD];
  auto load_raw = [&](int sl) {
    const int rows = P.img[sl].ph;
    const int4 m = meta2[sl];
    const long long row0 = ((long long)m.y << 32) | (unsigned int)m.x;
    const long long cbase = row0 * RC.pitch + k0 + c;
    const long long cstep = sl == 0 ? RC.pitch : -(long long)RC.pitch;
#pragma unroll
    for (int n1 = 0; n1 < NLOAD; ++n1) {
      const int y = N2 * n1 + r;
      raw[n1] = (col_ok && y < rows && m.w != 0) ? __ldg(RC.spec[sl] + cbase + cstep * y)
                                                : make_float2(0.f, 0.f);
    }
  };
  if (CACHED && r < N2) load_raw(0);
#pragma unroll
  for (int sl = 0; sl < 2; ++sl) {
    const int rows = P.img[sl].ph;
    const float2* Tb = T + ((size_t)sl * P.nb + blockIdx.y) * P.PY * P.nkx + k0 + c;
    float mean = 0.f;
    float2 rect = make_float2(0.f, 0.f), wk = make_float2(1.f, 0.f);
    bool cached_ok = false;
    if (CACHED) {
      cached_ok = meta2[sl].w != 0;
      mean = __int_as_float(meta2[sl].z);
      rect = fix3[sl];
      wk = fix3[2];
    }
    if (sl == 0) stage_twiddles_wait();
    __syncthreads();  // twiddles staged / previous readers of ex done
    if (r < N2) {
      float2 a[kN1];
#pragma unroll
      for (int n1 = 0; n1 < kN1; ++n1) {
        const int y = N2 * n1 + r;
        if (HALF && n1 >= kN1 / 2) {
          a[n1] = make_float2(0.f, 0.f);  // rows <= L / 2: pruned by constant folding
        } else if (CACHED) {
          float2 v = make_float2(0.f, 0.f);
          if (col_ok && y < rows) {
            if (cached_ok) {
              const float2 s = raw[n1 < NLOAD ? n1 : 0];
              // pre: s; post: W conj(s)
              const float2 t = sl == 0 ? s : make_float2(wk.x * s.x + wk.y * s.y,
                                                         wk.y * s.x - wk.x * s.y);
              v = make_float2(t.x - mean * rect.x, t.y - mean * rect.y);
            } else {
              v = make_float2(__int_as_float(0x7fc00000), 0.f);  // x start not cached
            }
          }
          a[n1] = v;
        } else {
          a[n1] = (col_ok && y < rows) ? __ldg(Tb + (size_t)y * P.nkx) : make_float2(0.f, 0.f);
        }
      }
      Dft<kN1>::run(a);
#pragma unroll
      for (int k1 = 0; k1 < kN1; ++k1)
        ex[c * D::EX + k1 * D::N2P + r] = cmul(a[k1], tw_s[r * k1]);
      if (CACHED && sl == 0) load_raw(1);  // in flight during pass 2 of the pre patch
    }
    __syncthreads();
    if (r < kN1) {
      float2 bq[N2];
#pragma unroll
      for (int n2 = 0; n2 < N2; ++n2) bq[n2] = ex[c * D::EX + r * D::N2P + n2];
      Dft<N2>::run(bq);
      if (sl == 0) {
#pragma unroll
        for (int k2 = 0; k2 < N2; ++k2) sa[c * D::LP + r + kN1 * k2] = bq[k2];
      } else {
#pragma unroll
        for (int k2 = 0; k2 < N2; ++k2)
          prod[k2] = swap_ri(cmul(bq[k2], sa[c * D::LP + r + kN1 * k2]));
      }
    }
  }
  // inverse: N2-point DFT over k2 (thread k1), twiddle, exchange, 16-point over k1.
  __syncthreads();
  if (r < kN1) {
    Dft<N2>::run(prod);
#pragma unroll
    for (int n2 = 0; n2 < N2; ++n2)
      ex[c * D::EX + r * D::N2P + n2] = cmul(prod[n2], tw_s[r * n2]);
  }
  __syncthreads();
  if (r < N2) {
    float2 a[kN1];
#pragma unroll
    for (int k1 = 0; k1 < kN1; ++k1) a[k1] = ex[c * D::EX + k1 * D::N2P + r];
    Dft<kN1>::run(a);
    if (col_ok) {
      float2* Ub = U + (size_t)blockIdx.y * P.sy * P.nkx + k0 + c;
#pragma unroll
      for (int n1 = 0; n1 < kN1; ++n1) {
        const int y = N2 * n1 + r;
        if (y < P.sy) Ub[(size_t)y * P.nkx] = swap_ri(a[n1]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------
// Stage 3: inverse row FFTs, crop, scale, and the first-peak search.
// grid = (row-pair groups, 1, pair), block = TR * N2 threads.
// keys[pair]: running (max value, lowest flat index) of the image, nanflag[pair].
// ---------------------------------------------------------------------------------
template <int N2, int TR>
__global__ void __launch_bounds__(TR * FastDims<N2>::G)
rows_inv_fast(Problem P, const float2* __restrict__ tw, const float2* __restrict__ U,
              float* __restrict__ images, float scale, unsigned long long* keys,
              int* nanflag, float* __restrict__ bandmax) {
  using D = FastDims<N2>;
  constexpr int L = D::L;
  __shared__ float2 ex[TR * D::EXR];
  __shared__ float2 tw_s[L];
  constexpr int NT = TR * D::G;
  __shared__ unsigned long long kred[(NT + 31) / 32];
  const int rp0 = blockIdx.x * TR;
  if (2 * rp0 >= P.sy) return;
  constexpr int NKX = L / 2 + 1;
  const float2* Ub = U + (size_t)blockIdx.z * P.sy * NKX;
  stage_twiddles_by_n2<N2, NT>(tw_s, tw);  // needed after the first DFT pass
  const int f = threadIdx.x / kN1, k1 = threadIdx.x % kN1;
  const int y = 2 * (rp0 + f);
  const bool line = threadIdx.x < TR * kN1 && y < P.sy;  // thread (f, k1) of a real row
  float2 bq[N2];
  if (line) {  // rows past the image: their transform is never read back
    {
      const float2* u0p = Ub + (size_t)y * NKX;
      const float2* u1p = u0p + NKX;
      // Bin k = k1 + 16 k2 of the Hermitian line: k2 is a compile-time constant after
      // unrolling, so whole groups of 16 bins are known to lie strictly inside (0, L/2)
      // (plain load) or strictly above L/2 (mirrored, conjugated load); only the groups
      // holding the DC or the Nyquist bin need the per-bin tests.
      auto gather = [&](auto pair_tag) {
        constexpr bool PAIR = decltype(pair_tag)::value;  // row y + 1 exists
#pragma unroll
        for (int k2 = 0; k2 < N2; ++k2) {
          const int lo = kN1 * k2;
          const float2 zero = make_float2(0.f, 0.f);
          if (lo > 0 && lo + kN1 - 1 < L / 2) {
            const float2 u0 = __ldg(u0p + lo + k1);
            const float2 u1 = PAIR ? __ldg(u1p + lo + k1) : zero;
            bq[k2] = make_float2(u0.y + u1.x, u0.x - u1.y);  // swap_ri(u0 + i u1)
          } else if (lo > L / 2) {
            const int kk = L - lo - k1;
            const float2 u0 = __ldg(u0p + kk);
            const float2 u1 = PAIR ? __ldg(u1p + kk) : zero;
            bq[k2] = make_float2(-u0.y + u1.x, u0.x + u1.y);  // conjugated bins
          } else {
            const int k = lo + k1;
            const bool mirror = k >= NKX;
            const int kk = mirror ? L - k : k;
            float2 u0 = __ldg(u0p + kk);
            float2 u1 = PAIR ? __ldg(u1p + kk) : zero;
            // c2r ignores the imaginary part of the DC and Nyquist bins.
            if (kk == 0 || kk == L / 2) { u0.y = 0.f; u1.y = 0.f; }
            if (mirror) { u0.y = -u0.y; u1.y = -u1.y; }
            bq[k2] = make_float2(u0.y + u1.x, u0.x - u1.y);
          }
        }
      };
      if (y + 1 < P.sy) gather(std::true_type{}); else gather(std::false_type{});
    }
  }
  if (line) Dft<N2>::run(bq);
  stage_twiddles_wait();
  __syncthreads();
  if (line) {
#pragma unroll
    for (int n2 = 0; n2 < N2; ++n2)
      ex[f * D::EXR + k1 * D::N2P + n2] = cmul(bq[n2], tw_s[n2 * kN1 + k1]);
  }
  __syncthreads();
  unsigned long long best = 0;
  int has_nan = 0;
  if (threadIdx.x < TR * N2) {  // thread (f, n2)
    const int f = threadIdx.x / N2, n2 = threadIdx.x - f * N2;
    const int y = 2 * (rp0 + f);
    if (y < P.sy) {
      float2 a[kN1];
#pragma unroll
      for (int k1 = 0; k1 < kN1; ++k1) a[k1] = ex[f * D::EXR + k1 * D::N2P + n2];
      Dft<kN1>::run(a);
      const bool pair = y + 1 < P.sy;
      float* out0 = images + ((size_t)(P.b0 + blockIdx.z) * P.sy + y) * P.sx + n2;
      float* out1 = out0 + P.sx;
      // Running maximum per output row with a strict compare in increasing x (keeps the
      // first of equal values); the 64-bit order-preserving key is built once per row.
      float bv0 = -INFINITY, bv1 = -INFINITY;
      int bx0 = n2, bx1 = n2;
      bool nan0 = false, nan1 = false;
#pragma unroll
      for (int n1 = 0; n1 < kN1; ++n1) {
        const int x = N2 * n1 + n2;
        if (x >= P.sx) continue;
        // after the re/im swap: real part -> row y, imaginary part -> row y + 1
        const float v0 = a[n1].y * scale, v1 = a[n1].x * scale;
        out0[N2 * n1] = v0;
        if (pair) out1[N2 * n1] = v1;
        nan0 |= v0 != v0;
        nan1 |= v1 != v1;
        if (v0 > bv0) { bv0 = v0; bx0 = x; }
        if (v1 > bv1) { bv1 = v1; bx1 = x; }
      }
      has_nan = (nan0 || (pair && nan1)) ? 1 : 0;
      best = peak_key(bv0, (unsigned)(y * P.sx + bx0));
      if (pair) {
        const unsigned long long k1key = peak_key(bv1, (unsigned)((y + 1) * P.sx + bx1));
        best = k1key > best ? k1key : best;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
    best = other > best ? other : best;
    has_nan |= __shfl_xor_sync(0xffffffffu, has_nan, o);
  }
  if ((threadIdx.x & 31) == 0) {
    kred[threadIdx.x >> 5] = best;
    if (has_nan) atomicOr(&nanflag[P.b0 + blockIdx.z], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (NT + 31) / 32; ++w) best = kred[w] > best ? kred[w] : best;
    atomicMax(&keys[P.b0 + blockIdx.z], best);
    // maximum of this band of 2 TR rows (NaN never compares above a threshold)
    float bv = -INFINITY;
    unsigned bi;
    if (best != 0) key_decode(best, &bv, &bi);
    bandmax[(P.b0 + blockIdx.z) * gridDim.x + blockIdx.x] = bv;
  }
}

}  // namespace flow
}  // namespace sofima
