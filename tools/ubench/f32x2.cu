// Micro-benchmark: throughput of scalar FADD/FMUL/FFMA vs packed FADD2/FMUL2/FFMA2 on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 r;
  asm volatile("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc; }"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  float2 r;
  asm volatile("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mul.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc; }"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 r;
  asm volatile("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd; }"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return r;
}
template <int MODE>
__global__ void k(float* out, float s) {
  float2 a[8];
  for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
  float2 b = make_float2(s, s * 0.5f);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { a[i].x = __fadd_rn(a[i].x, b.x); a[i].y = __fadd_rn(a[i].y, b.y); }
      if (MODE == 1) a[i] = add2(a[i], b);
      if (MODE == 2) { a[i].x = __fmul_rn(a[i].x, b.x); a[i].y = __fmul_rn(a[i].y, b.y); }
      if (MODE == 3) a[i] = mul2(a[i], b);
      if (MODE == 4) { a[i].x = __fmaf_rn(a[i].x, b.x, b.y); a[i].y = __fmaf_rn(a[i].y, b.y, b.x); }
      if (MODE == 5) a[i] = fma2(a[i], b, b);
    }
  }
  float acc = 0;
  for (int i = 0; i < 8; ++i) acc += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE> void run(const char* name, float* d) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int blocks = 148 * 8, threads = 256;
  k<MODE><<<blocks, threads>>>(d, 1.0001f);
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(d, 1.0001f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double elem_ops = (double)blocks * threads * ITERS * 16;  // fp32 results produced
  printf("%-8s %8.3f ms  %7.2f T fp32-results/s\n", name, ms, elem_ops / ms / 1e9);
}
int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  run<0>("FADD", d); run<1>("FADD2", d); run<2>("FMUL", d); run<3>("FMUL2", d); run<4>("FFMA", d); run<5>("FFMA2", d);
  return 0;
}
