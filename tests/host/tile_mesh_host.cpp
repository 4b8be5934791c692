// Host build of the tile-mesh arithmetic (sofima_b200/csrc/tile_mesh_core.cuh) for the
// CPU test suite: the SAME __host__ __device__ functions the CUDA kernel runs, driven by
// plain loops instead of a thread block, so that their results can be compared bit for bit
// with the golden vectors of the reference's own run without a GPU.  Test infrastructure:
// built and loaded only by tests/test_tile_mesh_host.py (g++ -O2 -ffp-contract=off).
#include "../../sofima_b200/csrc/tile_mesh_core.cuh"

using namespace sofima::tilemesh;

extern "C" {

void tile_mesh_force_host(const float* x, const float* cx, const float* cy, int ncomp, int nz,
                          int ny, int nx, float* out) {
  const Shape s{ncomp, nz, ny, nx};
  const long long m = s.nodes();
  for (int c = 0; c < ncomp; ++c)
    for (int z = 0; z < nz; ++z)
      for (int y = 0; y < ny; ++y)
        for (int xx = 0; xx < nx; ++xx)
          out[c * m + ((long long)z * ny + y) * nx + xx] = tile_force(x, cx, cy, s, c, z, y, xx);
}

// One chunk = the loop structure of tile_chunk_kernel without threads.
void tile_mesh_chunk_host(float* x, float* v, float* a, const float* cx, const float* cy,
                          int ncomp, int nz, int ny, int nx,
                          const sofima_integration_config* cfg, float* dt, float* alpha,
                          float* cap, int32_t* n_pos, double* e_kin, float* v_max) {
  const Shape s{ncomp, nz, ny, nx};
  const Chunk k = make_chunk(*cfg);
  const long long m = s.nodes();
  State st{*dt, *alpha, *cap, 1.0f, 0};
  tile_mesh_force_host(x, cx, cy, ncomp, nz, ny, nx, a);
  for (int it = 0; it < k.num_iters; ++it) {
    const State cur = st;
    for (long long n = 0; n < m; ++n) advance_node(x, v, a, s, n, k, cur);
    double power = 0.0;
    for (long long n = 0; n < m; ++n) power += kick_node(x, v, a, cx, cy, s, n, k, cur);
    if (!k.fire) continue;
    fire_update(&st, k, power);
    for (long long i = 0; i < m * ncomp; ++i) v[i] = v[i] * st.gate;
  }
  double e = 0.0;
  float vm = 0.f;
  bool nan_seen = false;
  for (long long n = 0; n < m; ++n) {
    const float sp = speed_node(v, s, n);
    e += (double)(sp * sp);
    if (sp != sp) nan_seen = true;
    vm = sp > vm ? sp : vm;
  }
  if (k.fire) {
    *dt = st.dt;
    *alpha = st.alpha;
    *cap = st.cap;
    *n_pos = st.n_pos;
  } else {
    *n_pos = -1;
  }
  *e_kin = e;
  *v_max = nan_seen ? NAN : vm;
}

}  // extern "C"
