/*
 * sofima_b200 -- C ABI of the B200 (sm_100a) backend for SOFIMA's two hot paths.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / C++ types.
 * Every entry point names the reference interface it replaces (file:line under
 * google-research/sofima @ efd7fd6).  All array pointers are DEVICE pointers on
 * the context's device unless a parameter says "host"; all work is enqueued on
 * the context's stream.  Return value: 0 = OK, otherwise a SOFIMA_E* code and
 * sofima_last_error() holds a message.
 *
 * There is no CPU fallback anywhere behind this interface.
 */
#ifndef SOFIMA_B200_H_
#define SOFIMA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SOFIMA_B200_ABI_VERSION 1

enum {
  SOFIMA_OK = 0,
  SOFIMA_EINVAL = 1,       /* bad argument (maps to ValueError / AssertionError) */
  SOFIMA_EUNSUPPORTED = 2, /* valid in the reference, not built yet (NotImplementedError) */
  SOFIMA_ECUDA = 3,        /* CUDA runtime error (RuntimeError) */
  SOFIMA_ENOMEM = 4
};

typedef struct sofima_ctx sofima_ctx;

/* Context = device + stream + cached scratch memory.  `stream` is a cudaStream_t
 * (NULL = legacy default stream).  One context per host thread. */
int sofima_ctx_create(int device, void* stream, sofima_ctx** out);
int sofima_ctx_destroy(sofima_ctx* ctx);
int sofima_ctx_set_stream(sofima_ctx* ctx, void* stream);
const char* sofima_last_error(const sofima_ctx* ctx); /* ctx may be NULL */
int sofima_abi_version(void);
/* Per-kernel timing (bench only): when on, every launch is bracketed by CUDA events
 * on the context's stream; the report is a JSON object {"kernel": {"ms", "n"}}
 * written to a host buffer, and clears the records. */
int sofima_ctx_set_timing(sofima_ctx* ctx, int on);
int sofima_ctx_timing_report(sofima_ctx* ctx, char* buf, int64_t buf_len);
/* Frees every cached scratch buffer larger than `keep_bytes` (the flow path keeps its
 * row-spectra cache and spectra scratch, several GB after a large call, for the next call
 * of the same geometry).  Synchronises the stream.  No reference counterpart: JAX frees
 * its temporaries when the jitted call returns (flow_field.py:683). */
int sofima_ctx_trim(sofima_ctx* ctx, int64_t keep_bytes);
/* Number of kernel launches issued through `ctx` so far (bench "gpu_launches"). */
int64_t sofima_ctx_launch_count(const sofima_ctx* ctx);

/* ------------------------------------------------------------------------- *
 *  Mesh relaxation  (reference: mesh.py)
 * ------------------------------------------------------------------------- */

/* Mirror of mesh.IntegrationConfig (mesh.py:282-338).  Doubles, because the
 * reference folds Python-float constants in double before rounding to fp32. */
typedef struct {
  double dt, gamma, k0, k;
  double stride[3]; /* xy[z]; stride[2] ignored for 2-d */
  int32_t num_iters;
  int32_t fire;
  double f_alpha, f_inc, f_dec, alpha;
  int32_t n_min;
  double dt_max;
  double start_cap, final_cap, cap_scale;
  int32_t cap_upscale_every;
  int32_t prefer_orig_order;
  int32_t remove_drift;
} sofima_integration_config;

enum {
  SOFIMA_FORCE_INPLANE = 0, /* mesh.inplane_force   (mesh.py:42-169)  */
  SOFIMA_FORCE_MESH3D = 1   /* mesh.elastic_mesh_3d (mesh.py:192-279) */
};

/* Mesh arrays are fp32, component-major: [ncomp][nb][nz][ny][nx] contiguous, with
 * ncomp = 2 (in-plane; the reference's [2, z, y, x] has nb = z, nz = 1) or 3
 * ([3, batch.., z, y, x] has nb = prod(batch)). */
typedef struct {
  int32_t ncomp;
  int64_t nb, nz, ny, nx;
  /* Number of batch dimensions folded into nb (3-d meshes only; 0 for [3, z, y, x]).
   * It only matters for remove_drift: the reference averages over axes (1, 2, 3)
   * literally (mesh.py:496-497), which for a [3, tiles, z, y, x] mesh is one mean per
   * x column over (tiles, z, y). */
  int32_t batch_rank;
} sofima_mesh_shape;

/* Replaces mesh.inplane_force / mesh.elastic_mesh_3d called on their own
 * (mesh.py:42, :192).  out = internal spring force field of x. */
int sofima_mesh_force(sofima_ctx* ctx, int force_kind, const float* x,
                      const sofima_mesh_shape* shape, double k,
                      const double* stride, int prefer_orig_order, float* out);

/* Same with the `links` argument of elastic_mesh_3d (mesh.py:197): links_xyz is a
 * HOST array [nlinks][3] of xyz directions in {-1,0,1}; NULL = MESH_LINK_DIRECTIONS. */
int sofima_mesh_force_links(sofima_ctx* ctx, int force_kind, const float* x,
                            const sofima_mesh_shape* shape, double k,
                            const double* stride, int prefer_orig_order,
                            const int32_t* links_xyz, int nlinks, float* out);

/* Replaces ONE call of mesh.velocity_verlet (mesh.py:371-521) plus the two
 * reductions relax_mesh does on its result (mesh.py:584-586).
 *   x, v       in/out, updated in place
 *   a          out (acceleration at the final x), may not be NULL
 *   prev       may be NULL (no inter-section springs)
 *   dt, alpha, cap   host pointers, in: fire_dt / fire_alpha / force_cap,
 *                    out: their values after the chunk (FIRE only)
 *   n_pos      host, out (FIRE only; -1 otherwise)
 *   e_kin      host, out: sum |v|^2      v_max: host, out: max |v|
 * Blocks until the chunk has finished (the reference blocks here too,
 * mesh.py:585). */
int sofima_mesh_chunk(sofima_ctx* ctx, int force_kind, float* x, float* v,
                      float* a, const float* prev,
                      const sofima_mesh_shape* shape,
                      const sofima_integration_config* cfg, float* dt,
                      float* alpha, float* cap, int32_t* n_pos, double* e_kin,
                      float* v_max);

/* Rigid tile-grid step between the coarse offsets and the fine flow of the stitching
 * notebooks: every tile is one node of a [ncomp][nb * nz][ny][nx] mesh (ncomp = 2: XY,
 * 3: XYZ offsets), cx / cy are the measured offsets to the +x / +y neighbour in the same
 * layout (NaN = no neighbour or no estimate).
 * sofima_tile_mesh_force replaces stitch_rigid.elastic_tile_mesh (stitch_rigid.py:330-391)
 * and elastic_tile_mesh_3d (:394-473) called on their own. */
int sofima_tile_mesh_force(sofima_ctx* ctx, const float* x, const float* cx,
                           const float* cy, const sofima_mesh_shape* shape, float* out);

/* Replaces ONE mesh.velocity_verlet call of stitch_rigid.optimize_coarse_mesh
 * (stitch_rigid.py:476-545: mesh_force = elastic_tile_mesh[_3d], prev = None) plus the two
 * reductions of relax_mesh (mesh.py:584-586); arguments as sofima_mesh_chunk.  The whole
 * chunk runs in one thread block.  cfg->k0, k, stride are unused (as in the reference);
 * cfg->remove_drift is not supported. */
int sofima_tile_mesh_chunk(sofima_ctx* ctx, float* x, float* v, float* a, const float* cx,
                           const float* cy, const sofima_mesh_shape* shape,
                           const sofima_integration_config* cfg, float* dt, float* alpha,
                           float* cap, int32_t* n_pos, double* e_kin, float* v_max);

/* Device-side solver state; also the layout of the read-back block. */
typedef struct {
  float dt, alpha, cap, gate; /* FIRE scalars after the last step; gate = (power >= 0) */
  int32_t n_pos;
  uint32_t ticket;            /* internal */
  float mean_x[3], mean_v[3]; /* remove_drift means of the last step */
  double power;               /* vdot(a, v) of the last step (mesh.py:455) */
  double e_kin;               /* sum |v|^2  (mesh.py:585) */
  float v_max;                /* max |v|    (mesh.py:586) */
  int32_t pad;
} sofima_mesh_state;

/* ---- Elastic tile stitching: the per-step target mesh ("prev_fn") -------------
 * Replaces stitch_elastic.compute_target_mesh vmapped over all tiles
 * (stitch_elastic.py:624-676, _update_mesh :573-620, _apply_flow :456-570,
 * map_utils.compose_maps_fast map_utils.py:616-734 with mode='constant'), i.e. the
 * `prev_fn` closure of notebooks/em_stitching.ipynb:545-549.  All pointers are
 * device pointers. */
typedef struct {
  const float* fx;      /* [ndim, ntiles, (fz,) fy, fx] flows between horizontal neighbours */
  const float* fy;      /* [ndim, ntiles, (fz,) fy, fx] flows between vertical neighbours */
  const int32_t* nbors; /* [ntiles, 4, 8 (2-d) or 11 (3-d)] NeighborInfo table
                           (stitch_elastic.py:43-72) */
  int32_t ndim;         /* 2: tile meshes [2, n, y, x]; 3: [3, n, z, y, x] */
  int64_t fx_shape[3];  /* zyx extent of one fx entry (2-d: [0] = 1) */
  int64_t fy_shape[3];
  double stride[3];     /* zyx stride of the mesh / flow grids in pixels (2-d: [0] unused) */
} sofima_stitch_target;

/* out[ndim, ntiles, (nz,) ny, nx] = target mesh of every tile for the tile meshes x of
 * the same shape (shape->nb = ntiles; 2-d: shape->nz = 1). */
int sofima_stitch_target_mesh(sofima_ctx* ctx, const float* x,
                              const sofima_mesh_shape* shape,
                              const sofima_stitch_target* target, float* out);

/* sofima_mesh_chunk for mesh.relax_mesh(x, None, config, prev_fn=...) with the
 * stitching prev_fn: `prev` is re-evaluated on the device from the advanced
 * positions inside every step (mesh.py:429-430).  target->ndim must match the force
 * kind (2: SOFIMA_FORCE_INPLANE, 3: SOFIMA_FORCE_MESH3D). */
int sofima_mesh_chunk_stitch(sofima_ctx* ctx, int force_kind, float* x, float* v,
                             float* a, const sofima_stitch_target* target,
                             const sofima_mesh_shape* shape,
                             const sofima_integration_config* cfg, float* dt,
                             float* alpha, float* cap, int32_t* n_pos,
                             double* e_kin, float* v_max);

/* Replaces map_utils.compose_maps_fast (map_utils.py:616-734): out = map2(map1(p))
 * on the grid of map1, relative format, bilinear (jax map_coordinates order 1).
 *   dim 2: map1 [2, z, y1, x1], map2 [2, z, y2, x2]; dim 3: [3, z, y, x] each
 *   shape1/shape2: zyx extents (3 values); start1/2, stride1/2: the last `dim` axes, zyx
 *   constant_mode: 0 = mode 'nearest', 1 = mode 'constant' with cval NaN */
int sofima_compose_maps(sofima_ctx* ctx, int dim, const float* map1,
                        const int64_t* shape1, const int64_t* start1,
                        const double* stride1, const float* map2,
                        const int64_t* shape2, const int64_t* start2,
                        const double* stride2, int constant_mode, float* out);

/* Same, without the final synchronisation: a sofima_mesh_state is copied to the
 * pinned host block `results_pinned` when the stream reaches that point.  Used to
 * queue many chunks back to back. */
int sofima_mesh_chunk_async(sofima_ctx* ctx, int force_kind, float* x, float* v,
                            float* a, const float* prev,
                            const sofima_mesh_shape* shape,
                            const sofima_integration_config* cfg, float dt,
                            float alpha, float cap,
                            sofima_mesh_state* results_pinned);

/* ---- One mesh sharded by rows over the GPUs of a node (BASELINE config 3) ----
 * No counterpart in the reference (a section is always solved on one device,
 * processor/mesh.py:462); numerically it is mesh.velocity_verlet on the whole mesh.
 * Every rank (one process per GPU) owns a slab of rows [2][nb][ny_local][nx];
 * slabs of ranks that have a lower neighbour must be a multiple of 32 rows.  The
 * step kernels read the neighbours' boundary rows and exchange the FIRE partial
 * sums through peer-mapped memory (CUDA IPC over NVLink), flag-synchronised on the
 * device: no host round trip and no collective call inside a chunk.
 *   create -> export (128-byte blob per rank) -> all-gather the blobs on the host
 *   -> connect -> set_state -> [host barrier] chunk ... -> get_state -> destroy
 * `chunk` returns the rank-LOCAL e_kin / v_max; the caller all-reduces them. */
typedef struct sofima_mesh_shard sofima_mesh_shard;
#define SOFIMA_SHARD_BLOB_BYTES 128
int sofima_shard_create(sofima_ctx* ctx, int rank, int nranks,
                        const sofima_mesh_shape* local_shape, sofima_mesh_shard** out);
int sofima_shard_export(sofima_mesh_shard* shard, void* blob128);
int sofima_shard_connect(sofima_mesh_shard* shard, const void* blobs /* nranks * 128 B */);
int sofima_shard_set_state(sofima_mesh_shard* shard, const float* x, const float* v_or_null,
                           const float* prev_or_null);
int sofima_shard_chunk(sofima_mesh_shard* shard, const sofima_integration_config* cfg,
                       float dt, float alpha, float cap, int64_t global_nodes,
                       sofima_mesh_state* result);
int sofima_shard_get_state(sofima_mesh_shard* shard, float* x, float* v, float* a);
int sofima_shard_destroy(sofima_mesh_shard* shard);

/* ------------------------------------------------------------------------- *
 *  Patch flow  (reference: flow_field.py)
 * ------------------------------------------------------------------------- */

enum { SOFIMA_U8 = 0, SOFIMA_F32 = 1, SOFIMA_U16 = 2, SOFIMA_U32 = 3 /* warp only */,
       SOFIMA_I16 = 4 /* warp_subvolume only */ };

typedef struct {
  int32_t ndim;            /* 2 or 3 */
  int32_t img_dtype;       /* SOFIMA_U8 | SOFIMA_F32 (both images) */
  int64_t pre_shape[3];    /* [[z,] y, x], leading entries first */
  int64_t post_shape[3];
  int64_t pre_mask_shape[3];  /* masks may be larger than the images */
  int64_t post_mask_shape[3];
  int32_t pre_patch[3];
  int32_t post_patch[3];
  int32_t has_mean;        /* 0: per-patch (masked) mean, 1: use `mean` */
  float mean;
  int32_t min_distance;    /* peak_min_distance, scalar (flow_field.py:235) */
  float threshold_rel;     /* 0.5 in the reference (flow_field.py:394) */
  int32_t peak_radius[3];
} sofima_xcorr_params;

/* Replaces flow_field.batched_xcorr_peaks (flow_field.py:385-441) for one batch:
 * gathers `batch` patch pairs at pre_starts / post_starts ([batch, ndim] int32,
 * [[z,] y, x] order), cross-correlates them (masked_xcorr, flow_field.py:36-156),
 * finds the two highest peaks with the reference's batch-coupled rule
 * (flow_field.py:259-268) and writes out_peaks[batch, ndim+2] =
 * (x, y[, z], sharpness, ratio).  Masks are uint8 (0/1) or NULL. */
int sofima_xcorr_peaks(sofima_ctx* ctx, const sofima_xcorr_params* p,
                       const void* pre_img, const void* post_img,
                       const uint8_t* pre_mask, const uint8_t* post_mask,
                       const int32_t* pre_starts, const int32_t* post_starts,
                       int64_t batch, float* out_peaks);

/* Optional accelerator for a series of sofima_xcorr_peaks / _images calls on the SAME
 * image pair (one flow_field call, flow_field.py:610-697).  Patches of a regular flow
 * grid overlap: the forward row transform of the raw pixels is computed once per image
 * row and distinct patch x-start and shared by all patches; the per-patch mean
 * (flow_field.py:340-353) and the flip of the post patch (:78-79) are applied, by
 * linearity of the transform, when the column stage loads the spectra.  Unmasked 2-d
 * patches on the two-pass transform lengths only.
 *   pre_xstarts / post_xstarts: HOST arrays of the distinct (clamped) x starts.
 * Later calls use the cache when images, shapes, dtype and patch widths match; a patch
 * whose x start is not in the set yields NaN.  The image contents must not change
 * while the cache is valid.  p == NULL drops the cache. */
int sofima_xcorr_rowcache(sofima_ctx* ctx, const sofima_xcorr_params* p,
                          const void* pre_img, const void* post_img,
                          const int32_t* pre_xstarts, int32_t n_pre,
                          const int32_t* post_xstarts, int32_t n_post);

/* Test hook: the raw correlation images of one batch, [batch, prod(pre+post-1)]
 * fp32 (what _batched_xcorr returns, flow_field.py:278-371). */
int sofima_xcorr_images(sofima_ctx* ctx, const sofima_xcorr_params* p,
                        const void* pre_img, const void* post_img,
                        const uint8_t* pre_mask, const uint8_t* post_mask,
                        const int32_t* pre_starts, const int32_t* post_starts,
                        int64_t batch, float* out_xcorr);

/* Replaces flow_field._batched_peaks (flow_field.py:205-275) on caller-supplied
 * correlation images img[batch, [z,] y, x] (the LICONN notebook calls it
 * directly).  center_offset in [[z,] y, x] order. */
int sofima_batched_peaks(sofima_ctx* ctx, int ndim, const float* img,
                         const int64_t* img_shape, int64_t batch,
                         const int32_t* center_offset, int min_distance,
                         float threshold_rel, const int32_t* peak_radius,
                         float* out_peaks);

/* ============================================================================
 *  Image warping  (reference: warp.py)
 * ========================================================================== */

/* Replaces the per-voxel work of warp.ndimage_warp (warp.py:189-335): every output
 * voxel interpolates the coordinate map at (index - offset) / stride (warp.py:300-306)
 * and samples the image there (warp.py:309), both exactly as
 * scipy.ndimage.map_coordinates(order, mode='constant', cval=0) in float64.
 *   dim 2 or 3; shapes / offset / stride hold `dim` values in [z]yx order
 *   image: device, img_dtype SOFIMA_U8 | _U16 | _U32 | _F32, image_shape
 *   src_map: device float64 [dim, *map_shape], the map in ABSOLUTE source voxel units
 *            (map_utils.to_absolute + box offset + out_scale, warp.py:245-260; the host
 *            side prepares it exactly as the reference does)
 *   order: 0 (nearest) or 1 (linear); out: device, same dtype as image, out_shape */
int sofima_warp_image(sofima_ctx* ctx, int dim, const void* image, int img_dtype,
                      const int64_t* image_shape, const double* src_map,
                      const int64_t* map_shape, const double* offset,
                      const double* stride, int order, void* out,
                      const int64_t* out_shape);

/* Replaces the per-pixel work of warp.warp_subvolume (warp.py:58-186): for every output
 * pixel of every section, densify the xy coordinate map (scipy RegularGridInterpolator,
 * linear with extrapolation, float64 -> float32; warp.py:144-153), quantise it like
 * cv2.convertMaps to CV_16SC2 (warp.py:155-160) and sample all channels like cv2.remap with
 * a zero constant border (warp.py:162-165).  Results are identical to scipy + OpenCV.
 *   image: device [n][nz][ih][iw] (image_shape = those four), img_dtype SOFIMA_U8 | _U16 |
 *          _I16 | _F32, or _U32 for any 4-byte integer with interpolation 0 (label ids)
 *   abs_map: device float64 [2][nz][my][mx], x then y coordinate of every map node in
 *            pixels of `image` (map_utils.to_absolute + box offsets, warp.py:123-126, done by
 *            the caller in the map's own dtype; map_is_f64 tells whether that dtype was
 *            float64 -- scipy evaluates the two cases in a different order)
 *   grid_y [my], grid_x [mx]: device float64, node positions in output pixels
 *            (warp.py:130-134), strictly ascending
 *   skip: device [nz] or NULL; sections with skip[z] != 0 are left zero (warp.py:117-119)
 *   interpolation: 0 nearest, 1 linear, 2 cubic, 3 lanczos (warp.py:33-40)
 *   out: device [n][nz][oh][ow], same dtype as image */
int sofima_warp_subvolume(sofima_ctx* ctx, const void* image, int img_dtype,
                          const int64_t* image_shape, const double* abs_map, int map_is_f64,
                          const double* grid_y, const double* grid_x, int64_t my, int64_t mx,
                          const uint8_t* skip, int interpolation, void* out, int64_t oh,
                          int64_t ow);

/* ------------------------------------------------------------------------- *
 *  Flow-field post-filters (reference: flow_utils.py, map_utils.py)
 * ------------------------------------------------------------------------- */

/* flow_utils.clean_flow (flow_utils.py:37-78).  flow: [nc][z][y][x] fp32 with nc = dim
 * (vectors) or dim + 2 (vectors, peak sharpness, peak ratio); out: [dim][z][y][x], rejected
 * vectors NaN.  max_magnitude / max_deviation <= 0 disable the respective test. */
int sofima_clean_flow(sofima_ctx* ctx, const float* flow, int nc, int dim, const int64_t* zyx,
                      float min_peak_ratio, float min_peak_sharpness, float max_magnitude,
                      float max_deviation, float* out);

/* flow_utils.reconcile_flows (flow_utils.py:81-135).  out: [nc][z][y][x] (nc = 2 or 3), on
 * entry the most preferred estimate, filtered in place; others: `nothers` further estimates
 * in order of decreasing preference (host array of device pointers). */
int sofima_reconcile_flows(sofima_ctx* ctx, float* out, const float* const* others, int nothers,
                           int nc, const int64_t* zyx, float max_gradient, float max_deviation,
                           int min_patch_size, float min_delta_z);

/* map_utils.mask_irregular (map_utils.py:737-786).  coord_map: [2][ny][nx] relative map,
 * masked nodes set to NaN in place; out_mask: [ny][nx] bytes (1 = masked). */
int sofima_mask_irregular(sofima_ctx* ctx, float* coord_map, int64_t ny, int64_t nx,
                          const double* stride_xy, double frac, double max_frac,
                          int dilation_iters, uint8_t* out_mask);

#ifdef __cplusplus
}
#endif
#endif /* SOFIMA_B200_H_ */
