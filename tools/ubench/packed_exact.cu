// Checks that the packed fp32x2 forms of mesh.cu produce the same bits as the scalar forms.
#include "../../sofima_b200/csrc/mesh.cu"
#include <cstdio>
#include <cstdlib>
using namespace sofima::mesh;
__global__ void check(const float* in, int n, int* bad) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* v = in + 16 * i;
  // sqrt / div
  float2 x = make_float2(fabsf(v[0]) * 100.f + 1e-3f, fabsf(v[1]) * 1e4f + 1e-3f);
  float2 s2 = sqrt_rn_unguarded2(x);
  if (__float_as_uint(s2.x) != __float_as_uint(sqrt_rn_unguarded(x.x)) || __float_as_uint(s2.y) != __float_as_uint(sqrt_rn_unguarded(x.y))) atomicAdd(&bad[0], 1);
  float2 a = make_float2(40.f, 56.568542f);
  float2 d2 = div_rn_unguarded2(a, s2);
  if (__float_as_uint(d2.x) != __float_as_uint(div_rn_unguarded(a.x, s2.x)) || __float_as_uint(d2.y) != __float_as_uint(div_rn_unguarded(a.y, s2.y))) atomicAdd(&bad[1], 1);
  float2 num = make_float2(v[2], v[3]); float den = fabsf(v[4]) + 1e-6f;
  float2 db = div_rn_unguarded_by(num, den);
  if (__float_as_uint(db.x) != __float_as_uint(div_rn_unguarded(num.x, den)) || __float_as_uint(db.y) != __float_as_uint(div_rn_unguarded(num.y, den))) atomicAdd(&bad[2], 1);
  // links
  Link L0, L1, L2, L3;
  L0.l0v[0] = 40.f; L0.l0v[1] = 0.f; L0.l0 = 40.f; L0.neg_k = -0.1f;
  L1.l0v[0] = 0.f; L1.l0v[1] = 40.f; L1.l0 = 40.f; L1.neg_k = -0.1f;
  L2.l0v[0] = 40.f; L2.l0v[1] = 40.f; L2.l0 = 56.568542f; L2.neg_k = -0.1f / sqrtf(2.f);
  L3.l0v[0] = -40.f; L3.l0v[1] = 40.f; L3.l0 = 56.568542f; L3.neg_k = -0.1f / sqrtf(2.f);
  float2 xf = make_float2(v[5] * 3, v[6] * 3), t0 = make_float2(v[7] * 3, v[8] * 3), t1 = make_float2(v[9] * 3, v[10] * 3);
  float2 t2 = make_float2(v[11] * 3, v[12] * 3), t3 = make_float2(v[13] * 3, v[14] * 3);
  for (int poo = 0; poo < 2; ++poo) {
    float2 f0, f1, f2, f3;
    link_pair<1, 0, 0, 1>(t0, t1, xf, L0, L1, poo, f0, f1);
    link_pair<1, 1, -1, 1>(t2, t3, xf, L2, L3, poo, f2, f3);
    float2 g0 = link2<1, 0>(t0, xf, L0.l0v[0], L0.l0v[1], L0.l0, L0.neg_k, poo);
    float2 g1 = link2<0, 1>(t1, xf, L1.l0v[0], L1.l0v[1], L1.l0, L1.neg_k, poo);
    float2 g2 = link2<1, 1>(t2, xf, L2.l0v[0], L2.l0v[1], L2.l0, L2.neg_k, poo);
    float2 g3 = link2<-1, 1>(t3, xf, L3.l0v[0], L3.l0v[1], L3.l0, L3.neg_k, poo);
    auto ne = [](float2 p, float2 q) { return __float_as_uint(p.x) != __float_as_uint(q.x) || __float_as_uint(p.y) != __float_as_uint(q.y); };
    if (ne(f0, g0)) atomicAdd(&bad[3], 1);
    if (ne(f1, g1)) atomicAdd(&bad[4], 1);
    if (ne(f2, g2)) atomicAdd(&bad[5], 1);
    if (ne(f3, g3)) atomicAdd(&bad[6], 1);
  }
}
int main() {
  const int n = 1 << 20;
  float* h = (float*)malloc(sizeof(float) * 16 * n);
  srand(1);
  for (long i = 0; i < 16L * n; ++i) h[i] = (float)rand() / RAND_MAX * 2.f - 1.f;
  float* d; int* bad; cudaMalloc(&d, sizeof(float) * 16 * n); cudaMalloc(&bad, 8 * sizeof(int));
  cudaMemcpy(d, h, sizeof(float) * 16 * n, cudaMemcpyHostToDevice); cudaMemset(bad, 0, 8 * sizeof(int));
  check<<<n / 256, 256>>>(d, n, bad);
  int hb[8]; cudaMemcpy(hb, bad, sizeof(hb), cudaMemcpyDeviceToHost);
  printf("mismatches: sqrt %d div %d div_by %d link0 %d link1 %d link2 %d link3 %d  (%s)\n", hb[0], hb[1], hb[2], hb[3], hb[4], hb[5], hb[6], cudaGetErrorString(cudaGetLastError()));
  return 0;
}
