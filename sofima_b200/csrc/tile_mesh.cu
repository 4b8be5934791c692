// Rigid tile-grid relaxation on the device (reference stitch_rigid.py:330-545): the
// one-node-per-tile mesh of stitch_rigid.optimize_coarse_mesh.  The system is tiny (tiles of
// a section: tens to a few thousand nodes) and strictly sequential in time, so a whole chunk
// of `num_iters` integration steps runs inside ONE thread block: no launch per step, the
// FIRE reduction is a block reduction in a fixed order.  The arithmetic lives in
// tile_mesh_core.cuh (shared with the host-compiled check of the tests).  Compiled with
// -fmad=false.
#include "common.cuh"
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

static int check_shape(sofima_ctx* ctx, const sofima_mesh_shape* shape, Shape* s) {
  if (!shape) return fail(ctx, SOFIMA_EINVAL, "shape is NULL");
  if (shape->ncomp != 2 && shape->ncomp != 3)
    return fail(ctx, SOFIMA_EINVAL, "tile mesh: ncomp must be 2 or 3");
  const long long nz = (long long)shape->nb * shape->nz;
  if (nz < 1 || shape->ny < 1 || shape->nx < 1 || nz * shape->ny * shape->nx > (1ll << 24))
    return fail(ctx, SOFIMA_EINVAL, "tile mesh: bad or too large shape");
  s->ncomp = shape->ncomp;
  s->nz = (int)nz;
  s->ny = (int)shape->ny;
  s->nx = (int)shape->nx;
  return SOFIMA_OK;
}

}  // namespace tilemesh
}  // namespace sofima

extern "C" {

int sofima_tile_mesh_force(sofima_ctx* ctx, const float* x, const float* cx, const float* cy,
                           const sofima_mesh_shape* shape, float* out) {
  using namespace sofima;
  if (!ctx) return fail(nullptr, SOFIMA_EINVAL, "ctx is NULL");
  if (!x || !cx || !cy || !out) return fail(ctx, SOFIMA_EINVAL, "NULL array argument");
  tilemesh::Shape s;
  int rc = tilemesh::check_shape(ctx, shape, &s);
  if (rc) return rc;
  DeviceGuard guard(ctx->device);
  const long long n = s.nodes() * s.ncomp;
  const unsigned grid = (unsigned)ceil_div<long long>(n, tilemesh::kThreads);
  LaunchTimer timer(ctx, "tile_mesh_force");
  tilemesh::tile_force_kernel<<<grid, tilemesh::kThreads, 0, ctx->stream>>>(x, cx, cy, s, out);
  SOFIMA_CHECK_LAUNCH(ctx);
  return SOFIMA_OK;
}

int sofima_tile_mesh_chunk(sofima_ctx* ctx, float* x, float* v, float* a, const float* cx,
                           const float* cy, const sofima_mesh_shape* shape,
                           const sofima_integration_config* cfg, float* dt, float* alpha,
                           float* cap, int32_t* n_pos, double* e_kin, float* v_max) {
  using namespace sofima;
  if (!ctx) return fail(nullptr, SOFIMA_EINVAL, "ctx is NULL");
  if (!x || !v || !a || !cx || !cy || !cfg || !dt || !alpha || !cap || !n_pos || !e_kin ||
      !v_max)
    return fail(ctx, SOFIMA_EINVAL, "NULL argument");
  tilemesh::Shape s;
  int rc = tilemesh::check_shape(ctx, shape, &s);
  if (rc) return rc;
  if (cfg->remove_drift)
    return fail(ctx, SOFIMA_EUNSUPPORTED, "tile mesh: remove_drift is not built");
  if (cfg->num_iters < 0) return fail(ctx, SOFIMA_EINVAL, "num_iters < 0");
  DeviceGuard guard(ctx->device);

  const tilemesh::Chunk k = tilemesh::make_chunk(*cfg);
  tilemesh::State st0;
  st0.dt = *dt;
  st0.alpha = *alpha;
  st0.cap = *cap;
  st0.gate = 1.0f;
  st0.n_pos = 0;  // n_pos restarts with every velocity_verlet call (mesh.py:513)

  void* dres = nullptr;
  if ((rc = scratch(ctx, "tilemesh.result", sizeof(tilemesh::Result), &dres))) return rc;
  {
    LaunchTimer timer(ctx, "tile_mesh_chunk");
    tilemesh::tile_chunk_kernel<<<1, tilemesh::kThreads, 0, ctx->stream>>>(
        x, v, a, cx, cy, s, k, st0, static_cast<tilemesh::Result*>(dres));
    SOFIMA_CHECK_LAUNCH(ctx);
  }
  tilemesh::Result hres;
  SOFIMA_CUDA(ctx, cudaMemcpyAsync(&hres, dres, sizeof(hres), cudaMemcpyDeviceToHost,
                                   ctx->stream));
  SOFIMA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (k.fire) {
    *dt = hres.st.dt;
    *alpha = hres.st.alpha;
    *cap = hres.st.cap;
    *n_pos = hres.st.n_pos;
  } else {
    *n_pos = -1;
  }
  *e_kin = hres.e_kin;
  *v_max = hres.v_max;
  return SOFIMA_OK;
}

}  // extern "C"
