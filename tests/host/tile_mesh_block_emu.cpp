// Host emulation of ONE CUDA thread block for the kernels of
// sofima_b200/csrc/tile_mesh_kernels.cuh: the kernel source is compiled unchanged, every CUDA
// thread is an OS thread, __syncthreads() is a std::barrier and __shared__ variables are
// function-local statics.  Built twice by tests/test_tile_mesh_host.py: plain (results are
// compared bit for bit with the reference's run) and with -fsanitize=thread (a missing
// barrier or a node touched by two threads shows up as a data race).  Test infrastructure.
#include <barrier>
#include <thread>
#include <vector>

struct EmuDim3 { unsigned x, y, z; };
static thread_local EmuDim3 threadIdx, blockIdx;
static EmuDim3 blockDim, gridDim;
static std::barrier<>* g_barrier = nullptr;
inline void __syncthreads() { g_barrier->arrive_and_wait(); }

#define __global__
#define __shared__ static
#define __restrict__
#define __launch_bounds__(...)

#include "../../sofima_b200/csrc/tile_mesh_kernels.cuh"

using namespace sofima::tilemesh;

template <typename F>
static void run_block(F&& body) {
  blockDim = {kThreads, 1, 1};
  gridDim = {1, 1, 1};
  std::barrier<> bar(kThreads);
  g_barrier = &bar;
  std::vector<std::thread> threads;
  for (unsigned t = 0; t < kThreads; ++t)
    threads.emplace_back([&body, t] {
      threadIdx = {t, 0, 0};
      blockIdx = {0, 0, 0};
      body();
    });
  for (auto& th : threads) th.join();
}

extern "C" {

void tile_mesh_chunk_emu(float* x, float* v, float* a, const float* cx, const float* cy,
                         int ncomp, int nz, int ny, int nx, const sofima_integration_config* cfg,
                         float* dt, float* alpha, float* cap, int32_t* n_pos, double* e_kin,
                         float* v_max) {
  const Shape s{ncomp, nz, ny, nx};
  const Chunk k = make_chunk(*cfg);
  const State st0{*dt, *alpha, *cap, 1.0f, 0};
  Result res;
  run_block([&] { tile_chunk_kernel(x, v, a, cx, cy, s, k, st0, &res); });
  if (k.fire) {
    *dt = res.st.dt;
    *alpha = res.st.alpha;
    *cap = res.st.cap;
    *n_pos = res.st.n_pos;
  } else {
    *n_pos = -1;
  }
  *e_kin = res.e_kin;
  *v_max = res.v_max;
}

}  // extern "C"

#ifdef EMU_MAIN
// Stand-alone run for ThreadSanitizer: FIRE and plain chunks on a grid smaller and a grid
// larger than the block, exit code 0 = ran to completion (TSAN sets its own on a race).
#include <cstdio>
#include <cstdlib>
#include <cstring>

int main() {
  const int shapes[2][2] = {{3, 4}, {20, 30}};
  for (int fire = 1; fire >= 0; --fire)
    for (const auto& yx : shapes) {
      const int ncomp = 2 + fire, nz = 1, ny = yx[0], nx = yx[1];
      const size_t n = (size_t)ncomp * nz * ny * nx;
      std::vector<float> x(n, 0.f), v(n, 0.f), a(n, 0.f), cx(n), cy(n);
      srand(7);
      for (size_t i = 0; i < n; ++i) {
        cx[i] = -40.f + (float)(rand() % 13 - 6);
        cy[i] = -30.f + (float)(rand() % 13 - 6);
      }
      for (int c = 0; c < ncomp; ++c)
        for (int y = 0; y < ny; ++y) {
          cx[((size_t)c * ny + y) * nx + nx - 1] = NAN;   // no +x neighbour
        }
      for (int c = 0; c < ncomp; ++c)
        for (int xx = 0; xx < nx; ++xx) cy[((size_t)c * ny + ny - 1) * nx + xx] = NAN;
      sofima_integration_config cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.dt = fire ? 0.001 : 0.05; cfg.gamma = fire ? 0.0 : 0.5; cfg.k = 0.1;
      cfg.num_iters = 40; cfg.fire = fire; cfg.f_alpha = 0.99; cfg.f_inc = 1.1; cfg.f_dec = 0.5;
      cfg.alpha = 0.1; cfg.n_min = 5; cfg.dt_max = 100; cfg.start_cap = cfg.final_cap = 1e6;
      cfg.cap_scale = 1.1; cfg.cap_upscale_every = 100;
      float dt = (float)cfg.dt, alpha = 0.1f, cap = 1e6f, v_max = 0.f;
      int32_t n_pos = 0;
      double e_kin = 0.0;
      for (int chunk = 0; chunk < 2; ++chunk)
        tile_mesh_chunk_emu(x.data(), v.data(), a.data(), cx.data(), cy.data(), ncomp, nz, ny, nx,
                            &cfg, &dt, &alpha, &cap, &n_pos, &e_kin, &v_max);
      printf("fire=%d %dx%d: dt=%g n_pos=%d e_kin=%g v_max=%g x0=%g\n", fire, ny, nx, dt, n_pos,
             e_kin, v_max, x[0]);
    }
  return 0;
}
#endif
