// One persistent kernel per reference batch for the unmasked correlation path:
// column transforms -> spectral product -> inverse columns -> inverse rows -> peak search,
// one patch pair at a time per thread block, nothing but a 152-byte record per pair leaves
// the SM's neighbourhood.
//
// Replaces the cols_fast -> rows_inv_fast -> peak2_kernel chain of flow_fast.cuh / flow.cu
// (reference flow_field.py:81-89 irfftn(rfftn * rfftn), :205-275 _batched_peaks, :178-202
// _peak_stats) whenever the shared row spectra are available (sofima_xcorr_rowcache).  That
// chain moves the product spectra (421 MB per batch of 1024 pairs) and the correlation
// images (417 MB per batch) through HBM twice each; here
//   * every block owns ONE scratch slot of sy x pitch complex values in global memory that
//     it rewrites for every pair -- 148 slots = 63 MB, resident in the 126 MB L2 (lines that
//     are rewritten in place are never evicted to HBM);
//   * the inverse row transform overwrites its own input rows with the correlation image
//     (319 floats fit in the 161 complex values they came from), so the image of the current
//     pair lives in the same slot and is read back by the peak search of the SAME block;
//   * the batch-coupled second-peak rule (flow_field.py:263-265: the first peaks of ALL batch
//     members are erased in EVERY row) needs the whole batch, so the block emits the 32 best
//     local maxima above the threshold; a one-thread-per-pair kernel applies the erase rule
//     once the batch's first peaks are known.  The rare pair whose 32 candidates are all
//     erased while more existed is recomputed by the same kernel in "fix-up" mode with the
//     finished bitmap (exact, like peak2_kernel).
//
// Thread block = 4 groups of 8 x G threads (G = max(N2, 16); 160 for the 320-point
// transforms of patch 160) plus one helper warp:
//   column phase: group g transforms the column groups g, g + 4, ... (8 spectral columns
//                 each, code of cols_fast); the helper warp does the last column
//                 (L / 2 + 1 = 8 N2 + 1 columns: the odd one out);
//   row phase:    group g transforms the row-pair groups g, g + 4, ... (code of rows_inv_fast);
//   peak phase:   all threads.
// Groups synchronise with named barriers; the block meets three times per pair.
// The arithmetic is instruction-for-instruction that of the three-kernel path: results are
// bit-identical (tests/test_flow_gpu.py::test_fused_pipeline_matches_three_kernel_path).
#pragma once

namespace sofima {
namespace flow {

constexpr int kFusedC = 8;          // spectral columns per column item
constexpr int kFusedTR = 8;         // row pairs per row item
constexpr int kFusedCand = 32;      // candidates stored per pair
constexpr int kFusedListCap = 640;  // local maxima kept in shared memory while scanning

struct PairPeaks {
  float v1;       // first peak value; -inf: no peak (whole row NaN, flow_field.py:194-196)
  int p1;         // flat index of the first peak (0 if none: argmax of an all -inf row)
  float mn;       // minimum of the sharpness window around the first peak
  float v0peak;   // image[0] if pixel 0 is a peak above the threshold, else -inf (:266-268)
  int npeaks;     // local maxima above the threshold; -1: more than the list could hold
  int nstored;    // entries of cv / ci (the best `nstored` by value, then lowest index)
  float cv[kFusedCand];
  int ci[kFusedCand];
};

// NG = 4: four groups + a helper warp for the odd last column (96 registers per thread);
// NG = 3: three groups, the last column is an ordinary (mostly empty) column item -- 21 items
//         = 7 rounds for the 320-point transforms -- and 128 registers per thread.
template <int N2, int NG>
struct FusedDims {
  using D = FastDims<N2>;
  static constexpr bool HELPER = false;
  static constexpr int GT = kFusedC * D::G;               // threads per group
  static constexpr int NT = NG * GT + (HELPER ? 32 : 0);  // + helper warp
  static constexpr int MAXREG = NG == 4 ? 96 : 128;
  static constexpr int EXC = kFusedC * D::EX;             // exchange, column items
  static constexpr int EXRW = kFusedTR * D::EXR;          // exchange, row items
  static constexpr int EXN = EXC > EXRW ? EXC : EXRW;     // float2 per group
  static constexpr int SAN = kFusedC * D::LP;             // float2 per group
  static constexpr size_t smem_bytes =
      sizeof(float2) * (D::L + NG * (EXN + SAN) + D::EX + D::LP) +
      sizeof(float) * kFusedListCap + sizeof(int) * kFusedListCap + 1024;
};

__device__ __forceinline__ void group_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// order-preserving 32-bit key of a float (for atomicMax on band maxima)
__device__ __forceinline__ unsigned f2ord(float v) {
  const unsigned u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// Scratch-slot accesses carry an L2 evict_last policy: the 63 MB of slots are rewritten in
// place for every pair and must not be displaced by the row spectra streaming through L2
// (tests/gpu_micro/l2_rewrite_probe.cu: with the hint the dirty lines stay on chip).
__device__ __forceinline__ uint64_t l2_evict_last_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void st_slot(float2* p, float2 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1,%2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y),
               "l"(pol) : "memory");
}
__device__ __forceinline__ void st_slot(float* p, float v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ float2 ld_slot(const float2* p, uint64_t pol) {
  float2 v;
  asm volatile("ld.global.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;" : "=f"(v.x), "=f"(v.y)
               : "l"(p), "l"(pol) : "memory");
  return v;
}
__device__ __forceinline__ float ld_slot(const float* p, uint64_t pol) {
  float v;
  asm volatile("ld.global.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol)
               : "memory");
  return v;
}

// img == zero-padded 5 x 5 (2 md + 1) maximum (flow_field.py:237-254) on a pitched image.
__device__ __forceinline__ bool is_peak_pitched(const float* im, int pitch, int sy, int sx,
                                                int md, int y, int x, float v, uint64_t pol) {
  float m = -INFINITY;
  for (int dy = -md; dy <= md; ++dy) {
    const int yy = y + dy;
    for (int dx = -md; dx <= md; ++dx) {
      const int xx = x + dx;
      const bool out = yy < 0 || yy >= sy || xx < 0 || xx >= sx;
      const float w = out ? 0.f : ld_slot(im + (size_t)yy * pitch + xx, pol);
      m = fmaxf(m, w);
    }
  }
  return v == m;
}

// One column item: forward transforms of C spectral columns of both patches, product,
// inverse transform, result rows into the block's scratch slot.  Code of cols_fast<N2, C,
// HALF = true, CACHED = true>; `sync` separates the passes of the C * G threads involved.
template <int N2, int C, class Sync>
__device__ __forceinline__ void fused_col_item(const Problem& P, const RowCacheView& RC,
                                               const int4 (&meta2)[2], const float2* tw_s,
                                               float2* ex, float2* sa, int k0, int c, int r,
                                               float2* U, int upitch, uint64_t pol, Sync sync) {
  using D = FastDims<N2>;
  const bool col_ok = k0 + c < P.nkx;
  float2 prod[N2];
  float2 fix3[3] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(1.f, 0.f)};
  if (col_ok) {
#pragma unroll
    for (int j = 0; j < 3; ++j) fix3[j] = __ldg(&RC.fix[j * P.nkx + k0 + c]);
  }
  constexpr int NLOAD = kN1 / 2;
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
  if (r < N2) load_raw(0);
#pragma unroll
  for (int sl = 0; sl < 2; ++sl) {
    const int rows = P.img[sl].ph;
    const bool cached_ok = meta2[sl].w != 0;
    const float mean = __int_as_float(meta2[sl].z);
    const float2 rect = fix3[sl];
    const float2 wk = fix3[2];
    sync();  // previous readers of ex are done
    if (r < N2) {
      float2 a[kN1];
#pragma unroll
      for (int n1 = 0; n1 < kN1; ++n1) {
        const int y = N2 * n1 + r;
        if (n1 >= kN1 / 2) {
          a[n1] = make_float2(0.f, 0.f);  // rows <= L / 2: pruned by constant folding
        } else {
          float2 v = make_float2(0.f, 0.f);
          if (col_ok && y < rows) {
            if (cached_ok) {
              const float2 s = raw[n1 < NLOAD ? n1 : 0];
              const float2 t = sl == 0 ? s : make_float2(wk.x * s.x + wk.y * s.y,
                                                         wk.y * s.x - wk.x * s.y);
              v = make_float2(t.x - mean * rect.x, t.y - mean * rect.y);
            } else {
              v = make_float2(__int_as_float(0x7fc00000), 0.f);  // x start not cached
            }
          }
          a[n1] = v;
        }
      }
      Dft<kN1>::run(a);
#pragma unroll
      for (int k1 = 0; k1 < kN1; ++k1)
        ex[c * D::EX + k1 * D::N2P + r] = cmul(a[k1], tw_s[r * k1]);
      if (sl == 0) load_raw(1);  // in flight during pass 2 of the pre patch
    }
    sync();
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
  sync();
  if (r < kN1) {
    Dft<N2>::run(prod);
#pragma unroll
    for (int n2 = 0; n2 < N2; ++n2)
      ex[c * D::EX + r * D::N2P + n2] = cmul(prod[n2], tw_s[r * n2]);
  }
  sync();
  if (r < N2) {
    float2 a[kN1];
#pragma unroll
    for (int k1 = 0; k1 < kN1; ++k1) a[k1] = ex[c * D::EX + k1 * D::N2P + r];
    Dft<kN1>::run(a);
    if (col_ok) {
      float2* Ub = U + k0 + c;
#pragma unroll
      for (int n1 = 0; n1 < kN1; ++n1) {
        const int y = N2 * n1 + r;
        if (y < P.sy) st_slot(Ub + (size_t)y * upitch, swap_ri(a[n1]), pol);
      }
    }
  }
}

// One row item: inverse transforms of TR Hermitian row pairs read from the scratch slot;
// the correlation image rows overwrite them in place (code of rows_inv_fast).  Returns the
// thread's best (value, lowest flat index) key of the rows it produced and its NaN flag.
template <int N2, class Sync>
__device__ __forceinline__ void fused_row_item(const Problem& P, const float2* tw_s, float2* ex,
                                               float2* U, int upitch, int rp0, int gt,
                                               float scale, unsigned long long& best,
                                               int& has_nan, float& band_best, uint64_t pol,
                                               Sync sync) {
  using D = FastDims<N2>;
  constexpr int L = D::L;
  constexpr int TR = kFusedTR;
  constexpr int NKX = L / 2 + 1;
  const int f = gt / kN1, k1 = gt % kN1;
  const int y = 2 * (rp0 + f);
  const bool line = gt < TR * kN1 && y < P.sy;
  float2 bq[N2];
  if (line) {
    const float2* u0p = U + (size_t)y * upitch;
    const float2* u1p = u0p + upitch;
    auto gather = [&](auto pair_tag) {
      constexpr bool PAIR = decltype(pair_tag)::value;
#pragma unroll
      for (int k2 = 0; k2 < N2; ++k2) {
        const int lo = kN1 * k2;
        const float2 zero = make_float2(0.f, 0.f);
        if (lo > 0 && lo + kN1 - 1 < L / 2) {
          const float2 u0 = ld_slot(u0p + lo + k1, pol);
          const float2 u1 = PAIR ? ld_slot(u1p + lo + k1, pol) : zero;
          bq[k2] = make_float2(u0.y + u1.x, u0.x - u1.y);
        } else if (lo > L / 2) {
          const int kk = L - lo - k1;
          const float2 u0 = ld_slot(u0p + kk, pol);
          const float2 u1 = PAIR ? ld_slot(u1p + kk, pol) : zero;
          bq[k2] = make_float2(-u0.y + u1.x, u0.x + u1.y);
        } else {
          const int k = lo + k1;
          const bool mirror = k >= NKX;
          const int kk = mirror ? L - k : k;
          float2 u0 = ld_slot(u0p + kk, pol);
          float2 u1 = PAIR ? ld_slot(u1p + kk, pol) : zero;
          if (kk == 0 || kk == L / 2) { u0.y = 0.f; u1.y = 0.f; }
          if (mirror) { u0.y = -u0.y; u1.y = -u1.y; }
          bq[k2] = make_float2(u0.y + u1.x, u0.x - u1.y);
        }
      }
    };
    if (y + 1 < P.sy) gather(std::true_type{}); else gather(std::false_type{});
    Dft<N2>::run(bq);
  }
  sync();  // previous readers of ex are done; every input row of this item has been read
  if (line) {
#pragma unroll
    for (int n2 = 0; n2 < N2; ++n2)
      ex[f * D::EXR + k1 * D::N2P + n2] = cmul(bq[n2], tw_s[k1 * n2]);
  }
  sync();
  band_best = -INFINITY;
  if (gt < TR * N2) {
    const int f2 = gt / N2, n2 = gt - f2 * N2;
    const int y2 = 2 * (rp0 + f2);
    if (y2 < P.sy) {
      float2 a[kN1];
#pragma unroll
      for (int kk = 0; kk < kN1; ++kk) a[kk] = ex[f2 * D::EXR + kk * D::N2P + n2];
      Dft<kN1>::run(a);
      const bool pair = y2 + 1 < P.sy;
      float* out0 = reinterpret_cast<float*>(U + (size_t)y2 * upitch) + n2;
      float* out1 = out0 + 2 * upitch;
      float bv0 = -INFINITY, bv1 = -INFINITY;
      int bx0 = n2, bx1 = n2;
      bool nan0 = false, nan1 = false;
#pragma unroll
      for (int n1 = 0; n1 < kN1; ++n1) {
        const int x = N2 * n1 + n2;
        if (x >= P.sx) continue;
        const float v0 = a[n1].y * scale, v1 = a[n1].x * scale;
        st_slot(out0 + N2 * n1, v0, pol);
        if (pair) st_slot(out1 + N2 * n1, v1, pol);
        nan0 |= v0 != v0;
        nan1 |= v1 != v1;
        if (v0 > bv0) { bv0 = v0; bx0 = x; }
        if (v1 > bv1) { bv1 = v1; bx1 = x; }
      }
      has_nan |= (nan0 || (pair && nan1)) ? 1 : 0;
      unsigned long long k = peak_key(bv0, (unsigned)(y2 * P.sx + bx0));
      band_best = bv0;
      if (pair) {
        const unsigned long long k1key = peak_key(bv1, (unsigned)((y2 + 1) * P.sx + bx1));
        k = k1key > k ? k1key : k;
        band_best = fmaxf(bv0, bv1);
      }
      best = k > best ? k : best;
    }
  }
}

// mode 0: pairs [0, B) in contiguous chunks per block, results -> recs.
// mode 1: fix-up of the pairs listed in fixlist[0 .. *nfix) with the finished first-peak
//         bitmap: exact second peak, final rows -> out_peaks.
template <int N2, int NG>
__global__ void __maxnreg__((FusedDims<N2, NG>::MAXREG))
pair_fused(const Problem P, const float2* __restrict__ tw, const RowCacheView RC,
           float2* __restrict__ scratch, int upitch, float scale, const PeakParams pp,
           PairPeaks* __restrict__ recs, int mode, const unsigned* __restrict__ bitmap,
           const int* __restrict__ fixlist, const int* __restrict__ nfix,
           float* __restrict__ out_peaks) {
  using D = FastDims<N2>;
  using F = FusedDims<N2, NG>;
  constexpr int L = D::L, GT = F::GT, NT = F::NT;
  extern __shared__ __align__(16) unsigned char fused_smem[];
  float2* tw_s = reinterpret_cast<float2*>(fused_smem);
  float2* gbase = tw_s + L;
  float2* exh = gbase + NG * (F::EXN + F::SAN);
  float2* sah = exh + D::EX;
  float* lv = reinterpret_cast<float*>(sah + D::LP);
  int* li = reinterpret_cast<int*>(lv + kFusedListCap);
  __shared__ unsigned long long kred[(NT + 31) / 32];
  __shared__ unsigned bandkey[64];
  __shared__ int s_nan, s_npk;
  __shared__ float s_min[(NT + 31) / 32];
  __shared__ unsigned long long s_key;

  const int tid = threadIdx.x;
  const bool helper = F::HELPER && tid >= NG * GT;
  const int g = tid / GT, gt = tid - g * GT;
  float2* ex = gbase + (helper ? 0 : g) * (F::EXN + F::SAN);
  float2* sa = ex + F::EXN;
  stage_twiddles<L, NT>(tw_s, tw);
  stage_twiddles_wait();
  __syncthreads();
  const uint64_t pol = l2_evict_last_policy();

  const long long B = P.nb;
  long long i0, i1;
  if (mode == 0) {
    const long long q = (B + gridDim.x - 1) / gridDim.x;
    i0 = (long long)blockIdx.x * q;
    i1 = i0 + q < B ? i0 + q : B;
  } else {
    i0 = blockIdx.x;
    i1 = *nfix;
  }
  const long long istep = mode == 0 ? 1 : gridDim.x;
  float2* U = scratch + (size_t)blockIdx.x * P.sy * upitch;
  const float* img = reinterpret_cast<const float*>(U);
  const int ipitch = 2 * upitch;
  const int nrp = (P.sy + 1) / 2;                        // row pairs
  const int nitems = (nrp + kFusedTR - 1) / kFusedTR;    // row items (bands of 16 rows)
  const int sx = P.sx, sy = P.sy;

  for (long long it = i0; it < i1; it += istep) {
    const long long b = P.b0 + (mode == 0 ? it : (long long)fixlist[it]);
    int4 meta2[2];
    meta2[0] = __ldg(&RC.meta[b * 2 + 0]);
    meta2[1] = __ldg(&RC.meta[b * 2 + 1]);
    if (tid < 64) bandkey[tid] = 0u;
    if (tid == 0) { s_nan = 0; s_npk = 0; }

    // ---- column phase
    if (!helper) {
      auto sync = [&]() { group_sync(1 + g, GT); };
      const int c = gt % kFusedC, r = gt / kFusedC;
      for (int cg = g; cg < (F::HELPER ? N2 : N2 + 1); cg += NG)
        fused_col_item<N2, kFusedC>(P, RC, meta2, tw_s, ex, sa, cg * kFusedC, c, r, U, upitch,
                                    pol, sync);
    } else {
      auto sync = [&]() { __syncwarp(); };
      // all 32 lanes enter (the passes are separated by __syncwarp); lanes >= G have no work
      const int r = tid - NG * GT;
      fused_col_item<N2, 1>(P, RC, meta2, tw_s, exh, sah, kFusedC * N2, 0, r, U, upitch, pol,
                              sync);
    }
    __syncthreads();  // the whole product spectrum is in the slot

    // ---- row phase
    unsigned long long best = 0;
    int has_nan = 0;
    if (!helper) {
      auto sync = [&]() { group_sync(1 + g, GT); };
      for (int item = g; item < nitems; item += NG) {
        float band_best;
        fused_row_item<N2>(P, tw_s, ex, U, upitch, item * kFusedTR, gt, scale, best, has_nan,
                           band_best, pol, sync);
        // maximum of this band of 16 rows (NaN never compares above a threshold)
        float bb = band_best;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bb = fmaxf(bb, __shfl_xor_sync(0xffffffffu, bb, o));
        if ((tid & 31) == 0 && bb > -INFINITY && item < 64) atomicMax(&bandkey[item], f2ord(bb));
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other > best ? other : best;
      has_nan |= __shfl_xor_sync(0xffffffffu, has_nan, o);
    }
    if ((tid & 31) == 0) {
      kred[tid >> 5] = best;
      if (has_nan) atomicOr(&s_nan, 1);
    }
    __syncthreads();  // the whole image is in the slot
    if (tid == 0) {
      unsigned long long k = 0;
      for (int w = 0; w < (NT + 31) / 32; ++w) k = kred[w] > k ? kred[w] : k;
      s_key = k;
    }
    __syncthreads();

    // ---- peak phase (flow_field.py:251-268, :178-202)
    const unsigned long long key = s_key;
    float v1 = -INFINITY;
    unsigned idx1 = 0;
    if (key != 0) key_decode(key, &v1, &idx1);
    const bool ok = !s_nan && key != 0 && (v1 > pp.thr_rel * v1);
    if (!ok) {
      if (mode == 0) {
        if (tid == 0) {
          PairPeaks* o = recs + b;
          o->v1 = -INFINITY; o->p1 = 0; o->mn = 1.f; o->v0peak = -INFINITY;
          o->npeaks = 0; o->nstored = 0;
        }
      } else if (tid < 4) {
        out_peaks[b * 4 + tid] = NAN;
      }
      __syncthreads();
      continue;
    }
    const int p1 = (int)idx1;
    const float thr = pp.thr_rel * v1;
    unsigned long long best2 = 0;  // mode 1
    for (int item = 0; item < nitems; ++item) {
      if (item < 64 && !(ord2f(bandkey[item]) > thr)) continue;
      const int ylo = item * 2 * kFusedTR;
      const int yhi = min(sy, ylo + 2 * kFusedTR);
      const int n = (yhi - ylo) * sx;
      for (int e = tid; e < n; e += NT) {
        const int yy = ylo + e / sx, xx = e % sx;
        const float v = ld_slot(img + (size_t)yy * ipitch + xx, pol);
        if (!(v > thr)) continue;
        const int flat = yy * sx + xx;
        if (mode == 1 && ((bitmap[flat >> 5] >> (flat & 31)) & 1u)) continue;
        if (!is_peak_pitched(img, ipitch, sy, sx, pp.md, yy, xx, v, pol)) continue;
        if (mode == 0) {
          const int slot = atomicAdd(&s_npk, 1);
          if (slot < kFusedListCap) { lv[slot] = v; li[slot] = flat; }
        } else {
          const unsigned long long k = peak_key(v, (unsigned)flat);
          best2 = k > best2 ? k : best2;
        }
      }
    }
    // sharpness window (clamped start, flow_field.py:190-192)
    const int px = p1 % sx, py = p1 / sx;
    const int wy = 2 * pp.ry + 1, wx = 2 * pp.rx + 1;
    const int y0 = clamp_start(py - pp.ry, wy, sy), x0 = clamp_start(px - pp.rx, wx, sx);
    float mn = INFINITY;
    for (int e = tid; e < wy * wx; e += NT)
      mn = fminf(mn, ld_slot(img + (size_t)(y0 + e / wx) * ipitch + x0 + e % wx, pol));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best2, o);
      best2 = other > best2 ? other : best2;
    }
    if ((tid & 31) == 0) { s_min[tid >> 5] = mn; kred[tid >> 5] = best2; }
    __syncthreads();
    if (tid < 32) {
      for (int w = 1; w < (NT + 31) / 32; ++w) {
        mn = fminf(mn, s_min[w]);
        best2 = kred[w] > best2 ? kred[w] : best2;
      }
      mn = fminf(mn, s_min[0]);
      best2 = kred[0] > best2 ? kred[0] : best2;
      const float v0 = ld_slot(img, pol);
      const float v0peak =
          (v0 > thr && is_peak_pitched(img, ipitch, sy, sx, pp.md, 0, 0, v0, pol)) ? v0 : -INFINITY;
      if (mode == 0) {
        PairPeaks* o = recs + b;
        const int total = s_npk;
        const int n = total < kFusedListCap ? total : kFusedListCap;
        int stored;
        if (n <= kFusedCand) {
          stored = n;
          if (tid < n) { o->cv[tid] = lv[tid]; o->ci[tid] = li[tid]; }
        } else {
          // the kFusedCand best entries: repeated extraction of the maximum key
          stored = kFusedCand;
          for (int round = 0; round < kFusedCand; ++round) {
            unsigned long long bk = 0;
            int bj = -1;
            for (int j = tid; j < n; j += 32) {
              const unsigned long long k = peak_key(lv[j], (unsigned)li[j]);
              if (k > bk) { bk = k; bj = j; }
            }
#pragma unroll
            for (int of = 16; of > 0; of >>= 1) {
              const unsigned long long ok2 = __shfl_xor_sync(0xffffffffu, bk, of);
              const int oj = __shfl_xor_sync(0xffffffffu, bj, of);
              if (ok2 > bk) { bk = ok2; bj = oj; }
            }
            if (tid == 0) {
              o->cv[round] = lv[bj];
              o->ci[round] = li[bj];
              lv[bj] = -INFINITY;  // peaks are > thr > 0: -inf marks "taken"
              li[bj] = 0x7fffffff;
            }
            __syncwarp();
          }
        }
        if (tid == 0) {
          o->v1 = v1; o->p1 = p1; o->mn = mn; o->v0peak = v0peak;
          o->npeaks = total > kFusedListCap ? -1 : total;
          o->nstored = stored;
        }
      } else if (tid == 0) {
        float v2 = v0peak;
        if (best2 != 0) {
          unsigned i2;
          key_decode(best2, &v2, &i2);
        }
        float* o = out_peaks + b * 4;
        o[0] = (float)px - (float)pp.cx;
        o[1] = (float)py - (float)pp.cy;
        o[2] = v1 / mn;
        o[3] = (v2 == -INFINITY || v2 == INFINITY) ? 0.0f : v1 / v2;
      }
    }
    __syncthreads();  // the slot and the candidate list are free for the next pair
  }
}

// First peaks of the batch -> bitmap (peak1_decode_kernel's rule).
__global__ void fused_bitmap_kernel(const PairPeaks* __restrict__ recs, long long B,
                                    unsigned* bitmap) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int p1 = recs[b].p1;
  atomicOr(&bitmap[p1 >> 5], 1u << (p1 & 31));
}

// Second peak under the batch-coupled erase rule + the output row (peak2_kernel's tail).
__global__ void fused_finalize_kernel(const PairPeaks* __restrict__ recs, long long B,
                                      const PeakParams pp, const unsigned* __restrict__ bitmap,
                                      float* __restrict__ out, int* fixlist, int* nfix) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const PairPeaks& r = recs[b];
  float* o = out + b * 4;
  if (r.v1 == -INFINITY) {
    o[0] = o[1] = o[2] = o[3] = NAN;
    return;
  }
  unsigned long long best = 0;
  for (int j = 0; j < r.nstored; ++j) {
    const int i = r.ci[j];
    if ((bitmap[i >> 5] >> (i & 31)) & 1u) continue;  // erased for every row (:263-265)
    const unsigned long long k = peak_key(r.cv[j], (unsigned)i);
    best = k > best ? k : best;
  }
  float v2;
  if (best != 0) {
    unsigned i2;
    key_decode(best, &v2, &i2);
  } else if (r.npeaks >= 0 && r.npeaks == r.nstored) {
    v2 = r.v0peak;  // nothing left: argmax of an all -inf row is index 0 (:266-268)
  } else {
    // every stored candidate is erased but more local maxima existed: exact recomputation
    fixlist[atomicAdd(nfix, 1)] = (int)b;
    return;
  }
  const int px = r.p1 % pp.sx, py = r.p1 / pp.sx;
  o[0] = (float)px - (float)pp.cx;
  o[1] = (float)py - (float)pp.cy;
  o[2] = r.v1 / r.mn;
  o[3] = (v2 == -INFINITY || v2 == INFINITY) ? 0.0f : r.v1 / v2;
}

}  // namespace flow
}  // namespace sofima
