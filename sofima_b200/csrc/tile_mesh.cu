// Rigid tile-grid relaxation on the device (reference stitch_rigid.py:330-545): the
// one-node-per-tile mesh of stitch_rigid.optimize_coarse_mesh.  The system is tiny (tiles of
// a section: tens to a few thousand nodes) and strictly sequential in time, so a whole chunk
// of `num_iters` integration steps runs inside ONE thread block: no launch per step, the
// FIRE reduction is a block reduction in a fixed order.  The arithmetic lives in
// tile_mesh_core.cuh (shared with the host-compiled check of the tests).  Compiled with
// -fmad=false.
#include "common.cuh"
#include "tile_mesh_kernels.cuh"

namespace sofima {
namespace tilemesh {

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
  // One thread block walks all nodes in every step: a tile grid is tens to a few thousand
  // nodes (one per image tile).  Anything far beyond that would keep a single SM busy for
  // minutes without a way to interrupt it, so it is refused instead.
  constexpr long long kMaxTileNodes = 1 << 16;
  if (s.nodes() > kMaxTileNodes)
    return fail(ctx, SOFIMA_EUNSUPPORTED, "tile mesh: %lld nodes exceed the single-block "
                "solver's limit of %lld", (long long)s.nodes(), kMaxTileNodes);
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
