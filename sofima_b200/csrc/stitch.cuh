// Stitching target mesh on the device: the `prev_fn` of elastic tile stitching.
//
// Replaces stitch_elastic.compute_target_mesh vmapped over all tiles
// (reference stitch_elastic.py:624-676 -> _update_mesh :573-620 -> _apply_flow
// :456-570 -> map_utils.compose_maps_fast map_utils.py:616-734, mode 'constant',
// cval NaN), which the reference re-evaluates inside EVERY integration step
// (mesh.py:429-430).
//
// One thread per mesh node of every tile.  For each of the tile's (up to) four
// neighbours, in table order: if the node lies in the paste rectangle of that
// neighbour's flow field, the neighbour's mesh is sampled bilinearly at the
// flow-displaced position -- jax.scipy.ndimage.map_coordinates(order=1): corner
// order (y0,x0) (y0,x1) (y1,x0) (y1,x1), weight product wy * wx, contributions
// added left to right, NaN for any out-of-range corner even at zero weight -- and a
// non-NaN result replaces the node's target; the last non-NaN update wins
// (dynamic_update_slice of where(isnan(update), previous, update)).  All in the
// reference's fp32 association order (this header is compiled into mesh.cu with
// -fmad=false).
//
// SRC selects where the node positions come from:
//   0  component-major x [2][tiles][my][mx]  (sofima_stitch_target_mesh)
//   1  packed working set, positions as stored (chunk start, mesh.py:501)
//   2  packed working set advanced by the pending step x + dt v + dt^2/2 a, with the
//      lazily applied FIRE gate / drift removal -- bit-identical to phase A of
//      mesh2d_kernel, so the step kernel that follows sees prev_fn(x_new)
//      (mesh.py:439, :429-430).
#pragma once

struct StitchParams {
  const float* fx;   // [2][nt][fx_ny][fx_nx]
  const float* fy;   // [2][nt][fy_ny][fy_nx]
  const int* nbors;  // [nt][4][8]  NeighborInfo (stitch_elastic.py:43-72)
  int nt, my, mx;
  int fx_ny, fx_nx, fy_ny, fy_nx;
  float stride_y, stride_x;
};

template <int SRC>
__global__ void __launch_bounds__(kThreads)
stitch_target2d_kernel(const Params p, const StitchParams q, int fire, float* out,
                       float2* outp) {
  if (SRC == 2) {  // inside the step loop: may be launched early behind the FIRE reduce kernel
    grid_dep_wait();
    grid_dep_launch();
  }
  const int t = blockIdx.y;
  const int node = blockIdx.x * kThreads + threadIdx.x;
  const int my = q.my, mx = q.mx;
  if (node >= my * mx) return;
  const int py = node / mx, px = node - py * mx;
  const float qnan = __int_as_float(0x7fc00000);

  float dt = 0.f, hdt2 = 0.f, gate = 1.f, mx0 = 0.f, mx1 = 0.f, mv0 = 0.f, mv1 = 0.f;
  bool lazy = false, drift = false;
  if (SRC == 2) {
    if (fire) {
      const State S = *p.state;
      dt = S.dt;
      gate = S.gate;
      hdt2 = 0.5f * (dt * dt);
      lazy = true;
      drift = p.drift != 0;
      if (drift) { mx0 = S.mean_x[0]; mx1 = S.mean_x[1]; mv0 = S.mean_v[0]; mv1 = S.mean_v[1]; }
    } else {
      dt = p.c_dt;
      hdt2 = p.c_hdt2;
    }
  }
  const long long tile_nodes = (long long)my * mx;
  auto position = [&](int tile, int iy, int ix) -> float2 {
    const long long o = tile * tile_nodes + (long long)iy * mx + ix;
    if (SRC == 0) return make_float2(__ldg(p.xi + o), __ldg(p.xi + o + p.comp_stride));
    const float4 s = __ldg(p.xvi + o);
    if (SRC == 1) return make_float2(s.x, s.y);
    const float2 aa = __ldg(p.pai + o);
    float x0 = s.x, x1 = s.y, v0 = s.z, v1 = s.w;
    if (lazy) {
      v0 = v0 * gate;
      v1 = v1 * gate;
      if (drift) { x0 = x0 - mx0; x1 = x1 - mx1; v0 = v0 - mv0; v1 = v1 - mv1; }
    }
    return make_float2(x0 + (dt * v0 + hdt2 * aa.x), x1 + (dt * v1 + hdt2 * aa.y));
  };

  // extended paste buffer of the reference (stitch_elastic.py:658-661)
  const int ext_y = my + max(q.fy_ny, q.fx_ny), ext_x = mx + max(q.fy_nx, q.fx_nx);
  float c0 = qnan, c1 = qnan;
  for (int k = 0; k < 4; ++k) {
    const int* nd = q.nbors + ((long long)t * 4 + k) * 8;
    const int nbor = nd[0];
    if (nbor == -1) continue;                       // stitch_elastic.py:604
    const int fidx = nd[1];
    const int mult = (nbor == fidx) ? 1 : -1;       // :598
    const bool horiz = nd[7] == 0;                  // :607
    const int d = horiz ? 0 : 1;
    const float* F = horiz ? q.fx : q.fy;
    const int fny = horiz ? q.fx_ny : q.fy_ny, fnx = horiz ? q.fx_nx : q.fy_nx;
    const int flow_overlap = nd[4], flow_ortho = nd[3], off_ortho = nd[2];
    const int par_size = horiz ? mx : my;           // nbor_mesh.shape[-dim - 1]
    const int ortho_size = horiz ? my : mx;         // nbor_mesh.shape[dim - 2]
    // paste position in the target (:536-547), clamped like dynamic_update_slice
    const int tg_par = (mult == 1) ? 0 : par_size - flow_overlap;
    const int tg_ortho = ((mult == 1 && off_ortho < 0) || (mult == -1 && off_ortho > 0))
                             ? ortho_size - flow_ortho : 0;
    int tgy = tg_par * d + (1 - d) * tg_ortho;
    int tgx = tg_par * (1 - d) + d * tg_ortho;
    tgy = min(max(tgy, 0), ext_y - fny);
    tgx = min(max(tgx, 0), ext_x - fnx);
    const int iy = py - tgy, ix = px - tgx;
    if (iy < 0 || iy >= fny || ix < 0 || ix >= fnx) continue;
    // source window in the neighbour's mesh (:483-497)
    const int st_par = (mult == 1) ? par_size - flow_overlap : 0;
    const int st_ortho = ((mult == 1 && off_ortho > 0) || (mult == -1 && off_ortho < 0))
                             ? ortho_size - flow_ortho : 0;
    const int sy = st_ortho * (1 - d) + d * st_par;
    const int sx = st_ortho * d + (1 - d) * st_par;
    // compose_maps_fast(flow, start, stride, nbor_mesh, 0, stride, 'constant')
    const int oy = min(sy, 0), ox = min(sx, 0);     // origin (map_utils.py:653)
    const long long fo = ((long long)fidx * fny + iy) * fnx + ix;
    const long long fcs = (long long)q.nt * fny * fnx;
    const float f0 = (float)mult * __ldg(F + fo);   // :500-502
    const float f1 = (float)mult * __ldg(F + fo + fcs);
    const float ref1x = (float)(ix + (sx - ox)) * q.stride_x;
    const float ref1y = (float)(iy + (sy - oy)) * q.stride_y;
    const float qx = (ref1x + f0) / q.stride_x;     // map_utils.py:674-675
    const float qy = (ref1y + f1) / q.stride_y;
    const float ly = floorf(qy), lx = floorf(qx);
    // NaN / out-of-int-range coordinates: the weights or every corner are invalid,
    // the update is NaN in both components and the previous value is kept.
    if (!(fabsf(ly) < 1.0e9f) || !(fabsf(lx) < 1.0e9f)) continue;
    const float uwy = qy - ly, uwx = qx - lx;
    const float lwy = 1.0f - uwy, lwx = 1.0f - uwx;
    const int iy0 = (int)ly, ix0 = (int)lx;
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int cy = iy0 + (c >> 1), cx = ix0 + (c & 1);
      const float w = ((c >> 1) ? uwy : lwy) * ((c & 1) ? uwx : lwx);
      float v0 = qnan, v1 = qnan;
      if (cy >= 0 && cy < my && cx >= 0 && cx < mx) {
        const float2 pos = position(nbor, cy, cx);
        v0 = pos.x + (float)(cx - ox) * q.stride_x;  // map2[0] + ref2[-1]
        v1 = pos.y + (float)(cy - oy) * q.stride_y;  // map2[1] + ref2[-2]
      }
      const float t0 = w * v0, t1 = w * v1;
      a0 = (c == 0) ? t0 : a0 + t0;
      a1 = (c == 0) ? t1 : a1 + t1;
    }
    float u0 = a0 - ref1x, u1 = a1 - ref1y;         // map_utils.py:686, :696
    u0 = u0 + (float)(mult * nd[5]);                 // stitch_elastic.py:524-527
    u1 = u1 + (float)(mult * nd[6]);
    if (u0 == u0) c0 = u0;                           // :566-569
    if (u1 == u1) c1 = u1;
  }
  const long long o = t * tile_nodes + node;
  if (out) {
    out[o] = c0;
    out[o + (long long)q.nt * tile_nodes] = c1;
  } else {
    outp[o] = make_float2(c0, c1);
  }
}

// ---------------------------------------------------------------------------------
// map_utils.compose_maps_fast (reference map_utils.py:616-734): out = map2(map1(p))
// on the grid of map1, relative format.  One thread per node of map1.  2-d maps are
// [2, z, y, x] with the sections composed pairwise (:669-697), 3-d maps [3, z, y, x]
// (:698-732).  mode 'nearest' clamps the corner indices, 'constant' yields NaN for a
// query touching any out-of-range corner (cval = NaN).
// ---------------------------------------------------------------------------------
struct ComposeParams {
  const float* map1;
  const float* map2;
  float* out;
  int n1[3], n2[3];      // zyx extents
  int s1[3], s2[3];      // zyx grid offsets: start - origin (2-d: [0] unused)
  float st1[3], st2[3];  // zyx strides (2-d: [0] unused)
  int constant_mode;
};

// floor(c) as int32 like XLA's saturating convert; the upper corner index wraps.
__device__ __forceinline__ void linear_nodes(float c, int size, bool constant_mode, int (&idx)[2],
                                             float (&w)[2], bool (&ok)[2]) {
  const float lower = floorf(c);
  w[1] = c - lower;
  w[0] = 1.0f - w[1];
  const int i0 = __float2int_rd(c);  // saturates, NaN -> 0 (the weights are NaN then)
  const int i1 = (int)((unsigned int)i0 + 1u);
  ok[0] = !constant_mode || (i0 >= 0 && i0 < size);
  ok[1] = !constant_mode || (i1 >= 0 && i1 < size);
  idx[0] = min(max(i0, 0), size - 1);
  idx[1] = min(max(i1, 0), size - 1);
}

template <int DIM>
__global__ void __launch_bounds__(kThreads) compose_maps_kernel(const ComposeParams q) {
  // grid = (x blocks, y, z): no index divisions
  const long long n = (long long)q.n1[0] * q.n1[1] * q.n1[2];
  const int x = blockIdx.x * kThreads + threadIdx.x;
  const int y = blockIdx.y, z = blockIdx.z;
  if (x >= q.n1[2]) return;
  const long long i = ((long long)z * q.n1[1] + y) * q.n1[2] + x;
  const long long cs2 = (long long)q.n2[0] * q.n2[1] * q.n2[2];
  const float qnan = __int_as_float(0x7fc00000);
  const bool cm = q.constant_mode != 0;

  float ref1[3], coord[3];
  const int pos[3] = {z, y, x};
#pragma unroll
  for (int a = 3 - DIM; a < 3; ++a) {
    ref1[a] = (float)(pos[a] + q.s1[a]) * q.st1[a];
    // component order is xyz, axis order zyx: component (2 - a) moves along axis a
    coord[a] = (ref1[a] + q.map1[(2 - a) * n + i]) / q.st2[a];
  }
  int idx[3][2];
  float w[3][2];
  bool ok[3][2];
#pragma unroll
  for (int a = 3 - DIM; a < 3; ++a) linear_nodes(coord[a], q.n2[a], cm, idx[a], w[a], ok[a]);

  float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < (1 << DIM); ++c) {
    const int bz = DIM == 3 ? (c >> 2) & 1 : 0, by = (c >> 1) & 1, bx = c & 1;
    const int cz = DIM == 3 ? idx[0][bz] : z, cy = idx[1][by], cx = idx[2][bx];
    bool valid = ok[1][by] && ok[2][bx];
    float wt = w[1][by] * w[2][bx];
    if (DIM == 3) {
      valid = valid && ok[0][bz];
      wt = (w[0][bz] * w[1][by]) * w[2][bx];
    }
    const long long o = ((long long)cz * q.n2[1] + cy) * q.n2[2] + cx;
    const int cpos[3] = {cz, cy, cx};
#pragma unroll
    for (int a = 3 - DIM; a < 3; ++a) {
      float v = qnan;
      if (valid) v = q.map2[(2 - a) * cs2 + o] + (float)(cpos[a] + q.s2[a]) * q.st2[a];
      const float t = wt * v;
      acc[a] = (c == 0) ? t : acc[a] + t;
    }
  }
#pragma unroll
  for (int a = 3 - DIM; a < 3; ++a) q.out[(2 - a) * n + i] = acc[a] - ref1[a];
}

// ---------------------------------------------------------------------------------
// 3-d tile meshes (LICONN in-plane stitching, x [3, tiles, z, y, x], neighbour rows of
// 11 entries): same construction with a z window (stitch_elastic.py:505-516, :549-556)
// and trilinear sampling (8 corners, z slowest; weights (wz * wy) * wx).
//   SRC 0: component-major positions as stored; SRC 2: advanced by the pending step
//   (the arithmetic of mesh3d_kernel's `advance`).
// ---------------------------------------------------------------------------------
struct StitchParams3 {
  const float* fx;   // [3][nt][fx_n zyx]
  const float* fy;
  const int* nbors;  // [nt][4][11]
  int nt;
  int m[3];          // mesh zyx
  int fxn[3], fyn[3];
  float stride[3];   // zyx
};

template <int SRC>
__global__ void __launch_bounds__(kThreads)
stitch_target3d_kernel(const Params p, const StitchParams3 q, int fire, float* out) {
  if (SRC == 2) {  // inside the step loop: may be launched early behind the FIRE reduce kernel
    grid_dep_wait();
    grid_dep_launch();
  }
  const int t = blockIdx.y;
  const long long tile_nodes = (long long)q.m[0] * q.m[1] * q.m[2];
  const long long node = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (node >= tile_nodes) return;
  const int px = (int)(node % q.m[2]);
  const int py = (int)((node / q.m[2]) % q.m[1]);
  const int pz = (int)(node / ((long long)q.m[2] * q.m[1]));
  const int ppos[3] = {pz, py, px};
  const float qnan = __int_as_float(0x7fc00000);
  const long long cs = p.comp_stride;

  float dt = 0.f, hdt2 = 0.f, gate = 1.f;
  float mx[3] = {0.f, 0.f, 0.f}, mv[3] = {0.f, 0.f, 0.f};
  bool lazy = false, drift = false;
  if (SRC == 2) {
    if (fire) {
      const State S = *p.state;
      dt = S.dt;
      gate = S.gate;
      hdt2 = 0.5f * (dt * dt);
      lazy = true;
      drift = p.drift == 1;
      if (drift)
        for (int c = 0; c < 3; ++c) { mx[c] = S.mean_x[c]; mv[c] = S.mean_v[c]; }
    } else {
      dt = p.c_dt;
      hdt2 = p.c_hdt2;
    }
  }
  const bool coldrift = lazy && p.drift == 2;  // per-column means, see Params::col_mean
  auto position = [&](long long o, int gx, int c) -> float {
    float xc = __ldg(p.xi + o + c * cs);
    if (SRC == 2) {
      float vc = __ldg(p.vi + o + c * cs);
      const float ac = __ldg(p.ai + o + c * cs);
      if (lazy) {
        vc = vc * gate;
        if (coldrift) {
          xc = xc - __ldcg(&p.col_mean[c * q.m[2] + gx]);
          vc = vc - __ldcg(&p.col_mean[(3 + c) * q.m[2] + gx]);
        } else if (drift) {
          xc = xc - mx[c];
          vc = vc - mv[c];
        }
      }
      xc = xc + (dt * vc + hdt2 * ac);
    }
    return xc;
  };

  float cur[3] = {qnan, qnan, qnan};  // xyz components
  for (int k = 0; k < 4; ++k) {
    const int* nd = q.nbors + ((long long)t * 4 + k) * 11;
    const int nbor = nd[0];
    if (nbor == -1) continue;
    const int fidx = nd[1];
    const int mult = (nbor == fidx) ? 1 : -1;
    const bool horiz = nd[7] == 0;
    const int d = horiz ? 0 : 1;
    const float* F = horiz ? q.fx : q.fy;
    const int* fn = horiz ? q.fxn : q.fyn;
    const int flow_overlap = nd[4], flow_ortho = nd[3], off_ortho = nd[2];
    const int off_z = nd[8], flow_z = nd[9];
    const int par_size = horiz ? q.m[2] : q.m[1];
    const int ortho_size = horiz ? q.m[1] : q.m[2];
    const int tg_par = (mult == 1) ? 0 : par_size - flow_overlap;
    const int tg_ortho = ((mult == 1 && off_ortho < 0) || (mult == -1 && off_ortho > 0))
                             ? ortho_size - flow_ortho : 0;
    const int tg_z = ((mult == 1 && off_z < 0) || (mult == -1 && off_z > 0)) ? q.m[0] - flow_z : 0;
    int tg[3] = {tg_z, tg_par * d + (1 - d) * tg_ortho, tg_par * (1 - d) + d * tg_ortho};
    int idx[3];
    bool inside = true;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int ext = q.m[a] + max(q.fyn[a], q.fxn[a]);
      tg[a] = min(max(tg[a], 0), ext - fn[a]);
      idx[a] = ppos[a] - tg[a];
      inside = inside && idx[a] >= 0 && idx[a] < fn[a];
    }
    if (!inside) continue;
    const int st_par = (mult == 1) ? par_size - flow_overlap : 0;
    const int st_ortho = ((mult == 1 && off_ortho > 0) || (mult == -1 && off_ortho < 0))
                             ? ortho_size - flow_ortho : 0;
    const int st_z = ((mult == 1 && off_z > 0) || (mult == -1 && off_z < 0)) ? q.m[0] - flow_z : 0;
    const int st[3] = {st_z, st_ortho * (1 - d) + d * st_par, st_ortho * d + (1 - d) * st_par};
    const long long fvol = (long long)fn[0] * fn[1] * fn[2];
    const long long fo = (long long)fidx * fvol + ((long long)idx[0] * fn[1] + idx[1]) * fn[2] + idx[2];
    const long long fcs = (long long)q.nt * fvol;
    int org[3];
    float ref1[3], coord[3];
    bool bad = false;
    int i0[3];
    float w[3][2];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      org[a] = min(st[a], 0);
      ref1[a] = (float)(idx[a] + (st[a] - org[a])) * q.stride[a];
      const float f = (float)mult * __ldg(F + fo + (2 - a) * fcs);  // component 2 - a
      coord[a] = (ref1[a] + f) / q.stride[a];
      const float lower = floorf(coord[a]);
      if (!(fabsf(lower) < 1.0e9f)) bad = true;
      w[a][1] = coord[a] - lower;
      w[a][0] = 1.0f - w[a][1];
      i0[a] = (int)lower;
    }
    if (bad) continue;  // NaN / far out of range: the update is NaN in all components
    float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int cz = i0[0] + ((c >> 2) & 1), cy = i0[1] + ((c >> 1) & 1), cx = i0[2] + (c & 1);
      const float wt = (w[0][(c >> 2) & 1] * w[1][(c >> 1) & 1]) * w[2][c & 1];
      const bool valid = cz >= 0 && cz < q.m[0] && cy >= 0 && cy < q.m[1] && cx >= 0 && cx < q.m[2];
      const int cpos[3] = {cz, cy, cx};
      const long long o = (long long)nbor * tile_nodes + ((long long)cz * q.m[1] + cy) * q.m[2] + cx;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        float v = qnan;
        if (valid) v = position(o, cx, 2 - a) + (float)(cpos[a] - org[a]) * q.stride[a];
        const float term = wt * v;
        acc[a] = (c == 0) ? term : acc[a] + term;
      }
    }
    const int fine[3] = {nd[10], nd[6], nd[5]};  // zyx
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float u = acc[a] - ref1[a];
      u = u + (float)(mult * fine[a]);
      if (u == u) cur[2 - a] = u;
    }
  }
  const long long o = (long long)t * tile_nodes + node;
#pragma unroll
  for (int c = 0; c < 3; ++c) out[o + c * (long long)q.nt * tile_nodes] = cur[c];
}
