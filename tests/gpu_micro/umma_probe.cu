// Bring-up probe for the Blackwell-only machinery the flow path uses (sm_100a):
//   1. tcgen05.mma kind::i8 (u8 x s8 -> s32 in TMEM) with shared-memory operand descriptors,
//      no-swizzle K-major layout written by threads;
//   2. the same with the A operand staged by TMA (cp.async.bulk.tensor.2d, SWIZZLE_32B boxes of
//      [128 rows][32 bytes] straight from a uint8 image), K = 160, N = 192;
//   3. a TMA box of complex64 values ([160 rows][8 columns] out of a [rows][168] array);
//   4. cp.async.bulk (1-d) of global data written earlier in the SAME kernel by ordinary
//      stores (generic -> async proxy ordering).
// Prints one PASS / FAIL line per case; exit code = number of failures.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(99);                                                                    \
    }                                                                              \
  } while (0)

// ------------------------------------------------------------------ device helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes,
                                              uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)layout_type << 61;
  return d;
}
// kind::i8 instruction descriptor: D = s32, A = u8 (0) / s8 (1), B likewise, both K-major.
__host__ __device__ constexpr uint32_t idesc_i8(int M, int N, int a_signed, int b_signed) {
  return (2u << 4) | ((uint32_t)a_signed << 7) | ((uint32_t)b_signed << 10) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; "
        "selp.u32 %0, 1, 0, p; }"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes,
                                          uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc,
                                       bool accumulate) {
  asm volatile(
      "{ .reg .pred p; setp.ne.b32 p, %4, 0; "
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p; }"
      ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]),
                 "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ------------------------------------------------------------------ cases 1 and 2
// D[128][N] = A[128][K] (u8) * B[N][K]^T (s8).  A_MODE 0: threads write A in the no-swizzle
// canonical layout; A_MODE 1: TMA writes K/32 boxes of [128][32 B] with SWIZZLE_32B.
// B is always written by threads, no-swizzle canonical layout.
template <int N, int K, int A_MODE>
__global__ void __launch_bounds__(128)
gemm_i8_probe(const uint8_t* __restrict__ A, int lda, const int8_t* __restrict__ B,
              int32_t* __restrict__ D, const __grid_constant__ CUtensorMap amap, int xt) {
  constexpr int KC = K / 16;  // 16-byte chunks along K
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                 // 128 * K bytes
  uint8_t* sB = smem + 128 * K;       // N * K bytes
  __shared__ __align__(8) uint64_t bar_tma, bar_mma;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  constexpr int TCOLS = N <= 32 ? 32 : N <= 64 ? 64 : N <= 128 ? 128 : N <= 256 ? 256 : 512;

  if (tid == 0) {
    mbar_init(&bar_tma, 1);
    mbar_init(&bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&tmem_base)), "r"(TCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // no-swizzle K-major canonical layout: element (row r, byte k) lives at
  //   (r / 8) * SBO + (k / 16) * LBO + (r % 8) * 16 + (k % 16)
  // with core matrices (8 rows x 16 B = 128 B) contiguous along K: LBO = 128, SBO = 128 * KC.
  constexpr uint32_t LBO = 128, SBO = 128 * KC;
  if (A_MODE == 0) {
    for (int i = tid; i < 128 * K; i += 128) {
      const int r = i / K, k = i % K;
      sA[(r / 8) * SBO + (k / 16) * LBO + (r % 8) * 16 + (k % 16)] = A[(size_t)r * lda + k];
    }
  }
  for (int i = tid; i < N * K; i += 128) {
    const int r = i / K, k = i % K;
    sB[(r / 8) * SBO + (k / 16) * LBO + (r % 8) * 16 + (k % 16)] = (uint8_t)B[(size_t)r * K + k];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // st.shared -> UMMA reads
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_base;

  if (tid == 0) {
    if (A_MODE == 1) {
      mbar_expect_tx(&bar_tma, 128 * K);
      for (int kb = 0; kb < K / 32; ++kb)  // box kb: image columns [32 kb, 32 kb + 32), rows 0..127
        tma_load_2d(sA + kb * 128 * 32, &amap, xt + kb * 32, 0, &bar_tma);
      mbar_wait(&bar_tma, 0);
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    constexpr uint32_t idesc = idesc_i8(128, N, 0, 1);
    for (int kb = 0; kb < K / 32; ++kb) {
      uint64_t da;
      if (A_MODE == 0)
        da = umma_desc(smem_u32(sA) + kb * 2 * LBO, LBO, SBO, 0);
      else  // SWIZZLE_32B box: rows 32 B apart, 8-row groups 256 B apart
        da = umma_desc(smem_u32(sA + kb * 128 * 32), 16, 256, 6);
      const uint64_t db = umma_desc(smem_u32(sB) + kb * 2 * LBO, LBO, SBO, 0);
      mma_i8(tbase, da, db, idesc, kb > 0);
    }
    mma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // D row = TMEM lane; warp w reads lanes 32 w .. 32 w + 31
  const int row = warp * 32 + (tid & 31);
  for (int c = 0; c < N; c += 8) {
    uint32_t r[8];
    tmem_ld8(tbase + ((uint32_t)(warp * 32) << 16) + c, r);
    for (int j = 0; j < 8; ++j) D[(size_t)row * N + c + j] = (int32_t)r[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(TCOLS));
}

// ------------------------------------------------------------------ case 3
__global__ void tma_c64_probe(const __grid_constant__ CUtensorMap map, int row0, int col0,
                              float2* out /*[160][8]*/) {
  __shared__ __align__(128) float2 tile[160 * 8];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, sizeof(tile));
    tma_load_2d(tile, &map, col0, row0, &bar);
  }
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < 160 * 8; i += blockDim.x) out[i] = tile[i];
}

// ------------------------------------------------------------------ case 4
__global__ void bulk_after_store_probe(float* scratch, int n, float* out, int rounds) {
  extern __shared__ __align__(128) float buf[];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  float acc = 0.f;
  for (int r = 0; r < rounds; ++r) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) scratch[i] = (float)(i * 3 + r);
    asm volatile("fence.proxy.async;" ::: "memory");  // generic stores -> async-proxy read
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_expect_tx(&bar, n * 4);
      bulk_load(buf, scratch, n * 4, &bar);
    }
    mbar_wait(&bar, r & 1);
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += buf[i] - (float)(i * 3 + r);
    __syncthreads();
  }
  out[threadIdx.x] = acc;  // 0 everywhere iff every round saw the fresh values
}

// ------------------------------------------------------------------ host
typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiled get_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (!fn || q != cudaDriverEntryPointSuccess) { printf("no cuTensorMapEncodeTiled\n"); exit(98); }
  return (EncodeTiled)fn;
}

template <int N, int K, int A_MODE>
static int run_gemm(const char* name, EncodeTiled enc) {
  const int W = getenv("PROBE_W") ? atoi(getenv("PROBE_W")) : 512, H = 128;  // A is a window of a [H][W] uint8 image starting at column 40
  std::vector<uint8_t> img((size_t)H * W);
  std::vector<int8_t> b((size_t)N * K);
  srand(1234 + N + K);
  for (auto& v : img) v = (uint8_t)(rand() & 255);
  for (auto& v : b) v = (int8_t)((rand() % 129) - 64);
  uint8_t* dimg; int8_t* db; int32_t* dd;
  CK(cudaMalloc(&dimg, img.size())); CK(cudaMalloc(&db, b.size())); CK(cudaMalloc(&dd, 128 * N * 4));
  CK(cudaMemcpy(dimg, img.data(), img.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, b.data(), b.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dd, 0xff, 128 * N * 4));
  const int x0 = 40;
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  if (A_MODE == 1) {
    // the tensor starts at image column x0 (40 is a multiple of 8 but not of 16: the TMA
    // needs a 16-byte aligned base, so the map covers the whole image and x0 is a coordinate)
    cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H};
    cuuint64_t strides[1] = {(cuuint64_t)W};
    cuuint32_t box[2] = {32, 128}, es[2] = {1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dimg, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B,
                     getenv("PROBE_L2_128") ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%s: FAIL (encode %d)\n", name, (int)r); return 1; }
  }
  const size_t smem = 128 * K + N * K;
  CK(cudaFuncSetAttribute(gemm_i8_probe<N, K, A_MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)smem));
  // A_MODE 1 box coordinates are relative to column 0: shift the image pointer view instead
  CUtensorMap m2 = map;
  const int xt = getenv("PROBE_X0") ? atoi(getenv("PROBE_X0")) : 0;
  gemm_i8_probe<N, K, A_MODE><<<1, 128, smem>>>(dimg + (A_MODE == 0 ? x0 : 0), W, db, dd, m2, xt);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: FAIL (%s)\n", name, cudaGetErrorString(e)); return 1; }
  std::vector<int32_t> d(128 * N);
  CK(cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost));
  const int xa = A_MODE == 0 ? x0 : xt;  // mode 1 reads columns [xt, xt + K) of the image
  long bad = 0;
  for (int r = 0; r < 128; ++r)
    for (int n = 0; n < N; ++n) {
      int32_t s = 0;
      for (int k = 0; k < K; ++k) s += (int32_t)img[(size_t)r * W + xa + k] * (int32_t)b[(size_t)n * K + k];
      if (s != d[(size_t)r * N + n]) {
        if (bad < 4) printf("  %s: D[%d][%d] = %d, want %d\n", name, r, n, d[(size_t)r * N + n], s);
        ++bad;
      }
    }
  printf("%s: %s (%ld mismatches)\n", name, bad ? "FAIL" : "PASS", bad);
  cudaFree(dimg); cudaFree(db); cudaFree(dd);
  return bad ? 1 : 0;
}

static int run_tma_c64(EncodeTiled enc) {
  const int rows = 400, pitch = 168;
  std::vector<float2> h((size_t)rows * pitch);
  for (size_t i = 0; i < h.size(); ++i) h[i] = make_float2((float)i, -(float)i);
  float2 *d, *o;
  CK(cudaMalloc(&d, h.size() * 8)); CK(cudaMalloc(&o, 160 * 8 * 8));
  CK(cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
  CUtensorMap map;
  cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)pitch * 8};
  cuuint32_t box[2] = {8, 160}, es[2] = {1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("tma_c64: FAIL (encode %d)\n", (int)r); return 1; }
  int fails = 0;
  const int cases[3][2] = {{0, 0}, {120, 24}, {240, 160}};  // last: rows 240..399, cols 160..167
  for (auto& c : cases) {
    tma_c64_probe<<<1, 128>>>(map, c[0], c[1], o);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("tma_c64: FAIL (%s)\n", cudaGetErrorString(e)); return 1; }
    std::vector<float2> got(160 * 8);
    CK(cudaMemcpy(got.data(), o, got.size() * 8, cudaMemcpyDeviceToHost));
    long bad = 0;
    for (int y = 0; y < 160; ++y)
      for (int x = 0; x < 8; ++x) {
        const float2 w = h[(size_t)(c[0] + y) * pitch + c[1] + x];
        if (got[y * 8 + x].x != w.x || got[y * 8 + x].y != w.y) ++bad;
      }
    printf("tma_c64 box at (row %d, col %d): %s (%ld mismatches)\n", c[0], c[1],
           bad ? "FAIL" : "PASS", bad);
    fails += bad ? 1 : 0;
  }
  cudaFree(d); cudaFree(o);
  return fails;
}

static int run_bulk() {
  const int n = 5376;  // 21.5 KB, the size of a 16-row product-spectrum tile
  float *s, *o;
  CK(cudaMalloc(&s, n * 4)); CK(cudaMalloc(&o, 256 * 4));
  CK(cudaFuncSetAttribute(bulk_after_store_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, n * 4));
  bulk_after_store_probe<<<1, 256, n * 4>>>(s, n, o, 50);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("bulk_after_store: FAIL (%s)\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> h(256);
  CK(cudaMemcpy(h.data(), o, 256 * 4, cudaMemcpyDeviceToHost));
  long bad = 0;
  for (float v : h) bad += v != 0.f;
  printf("bulk_after_store: %s (%ld threads saw stale data)\n", bad ? "FAIL" : "PASS", bad);
  cudaFree(s); cudaFree(o);
  return bad ? 1 : 0;
}

int main() {
  EncodeTiled enc = get_encode();
  int fails = 0;
  fails += run_gemm<64, 64, 0>("i8 gemm N=64 K=64, thread-written no-swizzle operands", enc);
  fails += run_gemm<192, 160, 0>("i8 gemm N=192 K=160, thread-written no-swizzle operands", enc);
  fails += run_gemm<64, 64, 1>("i8 gemm N=64 K=64, A by TMA SWIZZLE_32B", enc);
  fails += run_gemm<192, 160, 1>("i8 gemm N=192 K=160, A by TMA SWIZZLE_32B", enc);
  fails += run_tma_c64(enc);
  fails += run_bulk();
  printf("umma_probe: %d failing case(s)\n", fails);
  return fails;
}
