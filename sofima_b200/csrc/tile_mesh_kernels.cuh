// Kernels of the rigid tile-grid relaxation (see tile_mesh.cu).  Kept in a header of their
// own so that tests/host/tile_mesh_block_emu.cpp can compile the SAME kernel source for the
// host -- 256 OS threads, __syncthreads() as a std::barrier, ThreadSanitizer on -- and check
// the barrier / ownership structure without a GPU.
#pragma once

#include "tile_mesh_core.cuh"

namespace sofima {
namespace tilemesh {

constexpr int kThreads = 256;

struct Result {
  State st;
  int pad;
  double e_kin;
  float v_max;
  int pad2;
};

__global__ void __launch_bounds__(kThreads)
tile_force_kernel(const float* __restrict__ x, const float* __restrict__ cx,
                  const float* __restrict__ cy, Shape s, float* __restrict__ out) {
  const long long m = s.nodes();
  const long long plane = (long long)s.ny * s.nx;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < m * s.ncomp;
       i += (long long)gridDim.x * kThreads) {
    const int c = (int)(i / m);
    const long long n = i - c * m;
    const int z = (int)(n / plane);
    const int y = (int)((n - z * plane) / s.nx);
    const int xx = (int)(n - z * plane - (long long)y * s.nx);
    out[i] = tile_force(x, cx, cy, s, c, z, y, xx);
  }
}

// One chunk (mesh.py:371-521 with mesh_force = elastic_tile_mesh[_3d], prev = None) in a
// single block.  x, v, a: in / out.
__global__ void __launch_bounds__(kThreads)
tile_chunk_kernel(float* x, float* v, float* a, const float* __restrict__ cx,
                  const float* __restrict__ cy, Shape s, Chunk k, State st0, Result* res) {
  __shared__ double part[kThreads];
  __shared__ float fpart[kThreads];
  __shared__ State st;
  const long long m = s.nodes();
  const long long plane = (long long)s.ny * s.nx;
  if (threadIdx.x == 0) st = st0;
  // a = force(x) at the start of the chunk (mesh.py:427-434 before the loop)
  for (long long i = threadIdx.x; i < m * s.ncomp; i += kThreads) {
    const int c = (int)(i / m);
    const long long n = i - c * m;
    const int z = (int)(n / plane);
    const int y = (int)((n - z * plane) / s.nx);
    const int xx = (int)(n - z * plane - (long long)y * s.nx);
    a[i] = tile_force(x, cx, cy, s, c, z, y, xx);
  }
  __syncthreads();
  for (int it = 0; it < k.num_iters; ++it) {
    const State cur = st;  // every thread reads the scalars of this step
    for (long long n = threadIdx.x; n < m; n += kThreads) advance_node(x, v, a, s, n, k, cur);
    __syncthreads();  // all positions advanced before any force is evaluated
    double p = 0.0;
    for (long long n = threadIdx.x; n < m; n += kThreads)
      p += kick_node(x, v, a, cx, cy, s, n, k, cur);
    if (!k.fire) {  // uniform branch: plain velocity Verlet has no global coupling,
      __syncthreads();  // but every force must be evaluated before positions move again
      continue;
    }
    part[threadIdx.x] = p;
    __syncthreads();
    if (threadIdx.x == 0) {
      double power = 0.0;
      for (int t = 0; t < kThreads; ++t) power += part[t];  // fixed order
      State next = cur;
      fire_update(&next, k, power);
      st = next;
    }
    __syncthreads();
    const float gate = st.gate;
    // node n belongs to thread n % kThreads in every phase, so the next advance (own
    // nodes only) needs no barrier after this loop
    for (long long n = threadIdx.x; n < m; n += kThreads)
      for (int c = 0; c < s.ncomp; ++c) v[c * m + n] = v[c * m + n] * gate;
  }
  __syncthreads();
  // mesh.py:584-586: e_kin = sum |v|^2, v_max = max |v|
  double e = 0.0;
  float vm = 0.f;
  bool any_nan = false;
  for (long long n = threadIdx.x; n < m; n += kThreads) {
    const float sp = speed_node(v, s, n);
    e += (double)(sp * sp);
    if (sp != sp) any_nan = true;
    vm = sp > vm ? sp : vm;
  }
  part[threadIdx.x] = e;
  fpart[threadIdx.x] = any_nan ? NAN : vm;
  __syncthreads();
  if (threadIdx.x == 0) {
    double esum = 0.0;
    float vmax = 0.f;
    bool nan_seen = false;
    for (int t = 0; t < kThreads; ++t) {
      esum += part[t];
      if (fpart[t] != fpart[t]) nan_seen = true;
      vmax = fpart[t] > vmax ? fpart[t] : vmax;
    }
    res->st = st;
    res->e_kin = esum;
    res->v_max = nan_seen ? NAN : vmax;  // np.max propagates NaN
  }
}

}  // namespace tilemesh
}  // namespace sofima
