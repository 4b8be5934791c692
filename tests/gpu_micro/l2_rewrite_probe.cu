// Does a per-block scratch slot that is rewritten in place stay in L2, or do the dirty lines
// reach HBM anyway?  Each of `grid` blocks writes and reads back its own slot of `slot_kb`
// KB `rounds` times, optionally while streaming `stream_mb` MB of other data per round
// through L2 (like the row spectra of the flow path).  Run under
//   ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
// and compare the DRAM bytes with slot bytes x rounds.
//   variant 0: plain st.global.cg / ld.global.cg
//   variant 1: st / ld with an L2::evict_last policy on the slot, evict_first on the stream
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>
#include <cstdlib>

template <int VARIANT>
__global__ void __launch_bounds__(512, 1)
rewrite(float4* slots, size_t slot_f4, int rounds, const float4* stream, size_t stream_f4_per_block,
        float* sink) {
  float4* s = slots + (size_t)blockIdx.x * slot_f4;
  const float4* st = stream + (size_t)blockIdx.x * stream_f4_per_block;
  uint64_t pol_last = 0, pol_first = 0;
  if (VARIANT == 1) {
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
  }
  float acc = 0.f;
  for (int r = 0; r < rounds; ++r) {
    for (size_t i = threadIdx.x; i < slot_f4; i += blockDim.x) {
      const float4 v = make_float4((float)r, (float)i, 1.f, 2.f);
      if (VARIANT == 1)
        asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(s + i),
                     "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol_last) : "memory");
      else
        __stcg(s + i, v);
    }
    for (size_t i = threadIdx.x; i < stream_f4_per_block; i += blockDim.x) {
      float4 v;
      const float4* p = st + ((i + (size_t)r * 7919) % stream_f4_per_block);
      if (VARIANT == 1)
        asm volatile("ld.global.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol_first));
      else
        v = __ldcg(p);
      acc += v.x;
    }
    __syncthreads();
    for (size_t i = threadIdx.x; i < slot_f4; i += blockDim.x) {
      float4 v;
      if (VARIANT == 1)
        asm volatile("ld.global.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(s + i), "l"(pol_last));
      else
        v = __ldcg(s + i);
      acc += v.y;
    }
    __syncthreads();
  }
  if (acc == 12345.678f) sink[0] = acc;
}

int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int slot_kb = argc > 2 ? atoi(argv[2]) : 420;
  const int stream_kb = argc > 3 ? atoi(argv[3]) : 0;   // per block per round
  const int rounds = argc > 4 ? atoi(argv[4]) : 20;
  const int grid = 148;
  const size_t slot_f4 = (size_t)slot_kb * 1024 / 16, stream_f4 = (size_t)stream_kb * 1024 / 16;
  float4 *slots, *stream; float* sink;
  cudaMalloc(&slots, slot_f4 * 16 * grid);
  cudaMalloc(&stream, (stream_f4 ? stream_f4 : 1) * 16 * grid);
  cudaMalloc(&sink, 4);
  cudaMemset(stream, 0, (stream_f4 ? stream_f4 : 1) * 16 * grid);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    if (variant == 0) rewrite<0><<<grid, 512>>>(slots, slot_f4, rounds, stream, stream_f4, sink);
    else rewrite<1><<<grid, 512>>>(slots, slot_f4, rounds, stream, stream_f4, sink);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
  }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("variant %d slot %d KB x %d blocks = %.1f MB, stream %d KB/block/round, %d rounds: %.3f ms, "
         "slot traffic %.1f MB written + %.1f MB read, stream %.1f MB; err=%s\n", variant, slot_kb, grid,
         slot_kb * grid / 1024.0, stream_kb, rounds, ms, slot_kb * grid / 1024.0 * rounds,
         slot_kb * grid / 1024.0 * rounds, stream_kb * grid / 1024.0 * rounds,
         cudaGetErrorString(cudaGetLastError()));
  return 0;
}
