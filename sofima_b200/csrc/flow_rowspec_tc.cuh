// Shared row spectra on the tensor cores (sm_100a: TMA -> tcgen05.mma kind::i8 -> TMEM).
//
// sofima_xcorr_rowcache needs, for every image row Y and every distinct patch x start x0,
// the zero-padded length-L real DFT of the pw pixels [x0, x0 + pw) -- the forward row pass
// of flow_field.py:81-82 (rfftn) hoisted out of the per-patch work.  For uint8 images this
// is a GEMM with an EXACT integer formulation:
//
//     out[Y, k] = sum_x  pixel[Y, x0 + x] * w[x, k],   w = exp(-2 pi i x k / L)
//
//   A = the image itself: [128 rows][K = pw bytes] tiles pulled by TMA straight from the
//       uint8 image into the 32-byte-swizzled K-major layout the MMA reads (no conversion,
//       no register staging);
//   B = the twiddles as FOUR signed base-128 digits, w = d0 2^-6 + d1 2^-13 + d2 2^-20 +
//       d3 2^-27 (|error| <= 2^-28 per twiddle), one s8 matrix column per (bin, re/im, digit);
//   D = s32 accumulators in tensor memory: every product and every sum is exact.
//
// The epilogue reads the four digit sums of an output from TMEM and combines them in fp32.
// The result is the exact DFT of the pixels with twiddles good to 2^-28, i.e. closer to the
// true value than an fp32 FFT (the DC bin is exactly the pixel sum, which the per-patch mean
// relies on, flow_fast.cuh rowcache_meta_kernel).
//
// Block = 320 threads: warp 0 issues TMA, warp 1 issues the MMAs, warps 2..9 drain TMEM
// (warp w may only touch TMEM lanes 32 (w % 4) .. +31; two warps share a lane quarter and
// take 12 of the 24 bins each).  Each block keeps ONE chunk of B
// (24 frequency bins = 192 matrix columns) resident in shared memory and walks over its share
// of the (x start, 128-row tile) items with a two-stage A pipeline and two TMEM accumulators,
// so TMA, MMA and epilogue of consecutive items overlap.
#pragma once

#include <cuda.h>

namespace sofima {
namespace flow {

constexpr int kTcBins = 24;              // frequency bins per chunk
constexpr int kTcN = kTcBins * 8;        // matrix columns per chunk: (bin, re/im, digit)
constexpr int kTcRows = 128;             // image rows per item (= MMA M)
constexpr int kTcMaxK = 256;
constexpr int kTcOutPitch = kTcBins + 2; // float2 per staged output row (16-byte aligned rows)
constexpr int kTcEpiWarps = 8;           // two per TMEM lane quarter, 12 bins each
constexpr int kTcThreads = 64 + 32 * kTcEpiWarps;
constexpr int kTcStages = 4;             // A tiles in flight (TMA runs up to 4 items ahead)
__host__ __device__ constexpr size_t tc_smem_bytes(int K) {
  return (size_t)kTcStages * kTcRows * K + (size_t)kTcN * K +
         2 * (size_t)kTcRows * kTcOutPitch * sizeof(float2);
}

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
// Shared-memory matrix descriptor (K-major).  layout_type 0: no swizzle, core matrices of
// 8 rows x 16 bytes, LBO = distance between the two 16-byte K chunks of one MMA, SBO =
// distance between 8-row groups; layout_type 6: SWIZZLE_32B, rows 32 bytes apart.
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo,
                                            uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version
  d |= (uint64_t)layout_type << 61;
  return d;
}
// kind::i8 instruction descriptor: D = s32, A = u8, B = s8, both K-major, M x N.
__host__ __device__ constexpr uint32_t tc_idesc_u8s8(int M, int N) {
  return (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tc_mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc_smem_u32(bar)),
               "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; "
        "selp.u32 %0, 1, 0, p; }"
        : "=r"(ok) : "r"(tc_smem_u32(bar)), "r"(parity), "r"(20000u) : "memory");
  }
}

struct RowSpecTcJob {
  int h, nslots, pitch, nkx;
  int K;        // MMA K extent: (largest x-start misalignment + patch width) rounded up to 32
  int nchunks;  // ceil(nkx / kTcBins)
  int ndelta;   // distinct values of (x start mod 16)
  const int* xstarts;   // [nslots] device
  // The TMA wants the innermost box coordinate 16-byte aligned, so a row window starts at
  // x0 & ~15 and the twiddle rows are shifted by delta = x0 & 15 instead: one digit table per
  // distinct delta, and every block works on the slots of ONE delta.
  const int* dslots;    // [ndelta + 1] offsets into, then [nslots] slot indices grouped by delta
  const int8_t* btab;   // [ndelta][nchunks][kTcN * K] digits, no-swizzle canonical layout
};

__global__ void __launch_bounds__(kTcThreads, 1)
rowspec_tc_kernel(const __grid_constant__ CUtensorMap imap, const RowSpecTcJob J,
                  float2* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t tc_smem[];
  const int K = J.K;
  // [kTcStages] A tiles of 128 x K bytes (K / 32 swizzled boxes of 4 KB each), B, output tiles
  uint8_t* sB = tc_smem + kTcStages * kTcRows * K;
  float2* sOut = reinterpret_cast<float2*>(sB + kTcN * K);  // [2][128][kTcOutPitch]
  __shared__ __align__(8) uint64_t bar_full[kTcStages], bar_empty[kTcStages], bar_tfull[2],
      bar_tempty[2], bar_b;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const int ncombos = J.ndelta * J.nchunks;
  const int combo = blockIdx.x % ncombos;
  const int di = combo / J.nchunks, chunk = combo - di * J.nchunks;
  const int worker = blockIdx.x / ncombos, nworkers = gridDim.x / ncombos;
  const int ntiles = (J.h + kTcRows - 1) / kTcRows;
  const int* myslots = J.dslots + J.ndelta + 1 + __ldg(&J.dslots[di]);
  const int nitems = (__ldg(&J.dslots[di + 1]) - __ldg(&J.dslots[di])) * ntiles;

  if (tid == 0) {
    for (int s = 0; s < kTcStages; ++s) {
      tc_mbar_init(&bar_full[s], 1);
      tc_mbar_init(&bar_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      tc_mbar_init(&bar_tfull[s], 1);
      tc_mbar_init(&bar_tempty[s], 32 * kTcEpiWarps);
    }
    tc_mbar_init(&bar_b, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // 512 columns: two accumulators of kTcN = 192 columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;"
                 ::"r"(tc_smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      // the block's chunk of B, once
      const uint32_t bbytes = (uint32_t)(kTcN * K);
      tc_mbar_expect_tx(&bar_b, bbytes);
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
          ::"r"(tc_smem_u32(sB)), "l"(J.btab + (size_t)combo * bbytes), "r"(bbytes),
            "r"(tc_smem_u32(&bar_b)) : "memory");
      int n = 0;
      for (int it = worker; it < nitems; it += nworkers, ++n) {
        const int s = n % kTcStages;
        tc_mbar_wait(&bar_empty[s], ((n / kTcStages) & 1) ^ 1);
        const int si = it / ntiles, tile = it - si * ntiles;
        const int xa = __ldg(&J.xstarts[__ldg(&myslots[si])]) & ~15;
        tc_mbar_expect_tx(&bar_full[s], (uint32_t)(kTcRows * K));
        for (int kb = 0; kb < K / 32; ++kb)
          asm volatile(
              "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
              "[%0], [%1, {%2, %3}], [%4];"
              ::"r"(tc_smem_u32(tc_smem + (s * K + kb * 32) * kTcRows)), "l"(&imap),
                "r"(xa + kb * 32),
                "r"(tile * kTcRows), "r"(tc_smem_u32(&bar_full[s])) : "memory");
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      tc_mbar_wait(&bar_b, 0);
      const uint32_t idesc = tc_idesc_u8s8(kTcRows, kTcN);
      const uint32_t lbo = 128, sbo = 128u * (uint32_t)(K / 16);
      int n = 0;
      for (int it = worker; it < nitems; it += nworkers, ++n) {
        const int s = n % kTcStages, a = n & 1;
        tc_mbar_wait(&bar_full[s], (n / kTcStages) & 1);
        tc_mbar_wait(&bar_tempty[a], ((n >> 1) & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t dcol = tbase + (uint32_t)(a * 256);
        for (int kb = 0; kb < K / 32; ++kb) {
          const uint64_t da =
              tc_desc(tc_smem_u32(tc_smem + (s * K + kb * 32) * kTcRows), 16, 256, 6);
          const uint64_t db = tc_desc(tc_smem_u32(sB) + kb * 2 * lbo, lbo, sbo, 0);
          asm volatile(
              "{ .reg .pred p; setp.ne.b32 p, %4, 0; "
              "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p; }"
              ::"r"(dcol), "l"(da), "l"(db), "r"(idesc), "r"((uint32_t)(kb > 0)) : "memory");
        }
        // both commits fire when the MMAs above have completed
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                     ::"r"(tc_smem_u32(&bar_empty[s])) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                     ::"r"(tc_smem_u32(&bar_tfull[a])) : "memory");
      }
    }
  } else {
    const int q = warp & 3;                 // TMEM lane quarter of this warp
    const int half = (warp - 2) >> 2;       // which 12 of the chunk's 24 bins
    const int row_in_tile = q * 32 + lane;
    const int bin0 = chunk * kTcBins;
    const int et = tid - 64;                // 0 .. 255
    // cooperative store: float4 index i = et + 256 j -> (row i / 12, 16-byte column i % 12)
    constexpr int kVec = kTcBins / 2;       // float4 per output row of the chunk
    int n = 0;
    for (int it = worker; it < nitems; it += nworkers, ++n) {
      const int s = n & 1;
      const int si = it / ntiles, tile = it - si * ntiles;
      const int slot = __ldg(&myslots[si]);
      tc_mbar_wait(&bar_tfull[s], (n >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tbase + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * 256 + half * 96);
      float2* srow = sOut + ((size_t)s * kTcRows + row_in_tile) * kTcOutPitch + half * 12;
      // four bins (32 columns) per TMEM load; digits d0..d3 of re, then of im, per bin
#pragma unroll 1
      for (int b4 = 0; b4 < 12; b4 += 4) {
        uint32_t r[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
            "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
              "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]),
              "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
              "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
              "=r"(r[31])
            : "r"(taddr + (uint32_t)(b4 * 8)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (b4 + 4 >= 12) {  // last load of this accumulator: hand it back to the MMA warp
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          tc_mbar_arrive(&bar_tempty[s]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int* d = reinterpret_cast<const int*>(r) + 8 * j;
          // The digit sums are exact integers below 2^23; two of them are merged in integer
          // arithmetic (d0 128 + d1, d2 128 + d3 < 2^31), so that an output costs two
          // int -> float conversions: value = hi 2^-13 + lo 2^-27, good to one fp32 ulp.
          const float re = fmaf((float)(d[0] * 128 + d[1]), 1.220703125e-4f,
                                (float)(d[2] * 128 + d[3]) * 7.450580596923828125e-9f);
          const float im = fmaf((float)(d[4] * 128 + d[5]), 1.220703125e-4f,
                                (float)(d[6] * 128 + d[7]) * 7.450580596923828125e-9f);
          srow[b4 + j] = make_float2(re, im);
        }
      }
      // the 128 x 24 tile goes out in row-major order, 16 bytes per thread
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kTcEpiWarps) : "memory");
      const float4* st = reinterpret_cast<const float4*>(sOut + (size_t)s * kTcRows * kTcOutPitch);
      float4* obase = reinterpret_cast<float4*>(out + ((size_t)slot * J.h + tile * kTcRows) * J.pitch + bin0);
      const int rows_ok = min(kTcRows, J.h - tile * kTcRows);
      const int opitch4 = J.pitch / 2;
      int rr = et / kVec, cc = et - rr * kVec;
#pragma unroll
      for (int j = 0; j < kTcRows * kVec / (32 * kTcEpiWarps); ++j) {
        if (rr < rows_ok && bin0 + 2 * cc < J.pitch)  // the last chunk may overhang the row pitch
          obase[(size_t)rr * opitch4 + cc] = st[rr * (kTcOutPitch / 2) + cc];
        rr += (32 * kTcEpiWarps) / kVec;
        cc += (32 * kTcEpiWarps) % kVec;
        if (cc >= kVec) { cc -= kVec; ++rr; }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase));
}

}  // namespace flow
}  // namespace sofima
