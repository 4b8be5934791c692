// Rigid tile-grid relaxation (reference stitch_rigid.py:330-545): every tile is ONE node of
// a small mesh [ncomp][nz][ny][nx] (ncomp = 2: XY offsets, 3: XYZ), the springs want the
// measured coarse offsets cx (to the +x neighbour) and cy (to the +y neighbour), and the
// mesh is relaxed by mesh.relax_mesh with this force instead of the Hookean one
// (stitch_rigid.optimize_coarse_mesh, :476-545).
//
// This header holds the per-node arithmetic and the scalar FIRE update as
// __host__ __device__ functions: tile_mesh.cu runs them in one thread block, and
// tests/host/tile_mesh_host.cpp compiles the SAME functions for the host, where the result
// is compared bit for bit with the golden vectors of the reference's own run
// (tests/golden/coarse_golden.npz).  Compile without FMA contraction (-fmad=false /
// -ffp-contract=off): every operation below is one IEEE fp32 operation in the order of
// mesh.py:436-499.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/sofima_b200.h"

#ifndef __CUDACC__
#ifndef __host__
#define __host__
#endif
#ifndef __device__
#define __device__
#endif
#endif

namespace sofima {
namespace tilemesh {

struct Shape {
  int ncomp, nz, ny, nx;
  __host__ __device__ long long nodes() const { return (long long)nz * ny * nx; }
};

// Scalars of one chunk (mesh.py:371-521), rounded on the host exactly where the
// reference rounds them.
struct Chunk {
  int fire, num_iters, n_min, cap_upscale_every;
  // FIRE: fp32 scalars of mesh.py:448-499
  float gamma, f_inc, f_dec, f_alpha, alpha0, dt_ceiling, cap_scale, final_cap;
  // plain velocity Verlet: Python-float constants folded in double, rounded once
  float c_dt, c_hdt2, c_hdt, fact0, fact1;
};

struct State {
  float dt, alpha, cap, gate;
  int n_pos;
};

// Scalars of a chunk from the IntegrationConfig mirror, rounded where the reference rounds
// them: FIRE works on fp32 scalars throughout (mesh.py:448-499, JAX x64 off); the plain
// integrator folds Python floats in double and rounds once (mesh.py:439-445).
inline Chunk make_chunk(const sofima_integration_config& cfg) {
  Chunk k;
  memset(&k, 0, sizeof(k));
  k.fire = cfg.fire != 0;
  k.num_iters = cfg.num_iters;
  k.n_min = cfg.n_min;
  k.cap_upscale_every = cfg.cap_upscale_every > 0 ? cfg.cap_upscale_every : 1;
  k.gamma = (float)cfg.gamma;
  k.f_inc = (float)cfg.f_inc;
  k.f_dec = (float)cfg.f_dec;
  k.f_alpha = (float)cfg.f_alpha;
  k.alpha0 = (float)cfg.alpha;
  k.dt_ceiling = (float)(cfg.dt_max * cfg.dt);
  k.cap_scale = (float)cfg.cap_scale;
  k.final_cap = (float)cfg.final_cap;
  k.c_dt = (float)cfg.dt;
  k.c_hdt2 = (float)(0.5 * (cfg.dt * cfg.dt));
  k.c_hdt = (float)(0.5 * cfg.dt);
  k.fact0 = (float)(1.0 / (1.0 + 0.5 * cfg.dt * cfg.gamma));
  k.fact1 = (float)(1.0 - 0.5 * cfg.dt * cfg.gamma);
  return k;
}

// jnp.nan_to_num defaults: NaN -> 0, +-inf -> +-FLT_MAX.
__host__ __device__ inline float nan_to_num(float f) {
  if (f != f) return 0.f;
  if (f > 3.4028234663852886e38f) return 3.4028234663852886e38f;
  if (f < -3.4028234663852886e38f) return -3.4028234663852886e38f;
  return f;
}

// One spring term of stitch_rigid.py:355-390 at node (z, y, x), component c, along the x
// (axis == 0) or y (axis == 1) neighbour direction, with target array t (cx or cy):
//   f = nan_to_num(x[c, next] - x[c, this] - t[c, this]);  this += f;  next -= f
// The reference adds zero-padded copies of f to the whole force array, i.e. per node
//   acc = (acc + f(this -> next)) - f(prev -> this),   missing neighbours contribute 0.
__host__ __device__ inline float spring_term(const float* xs, const float* t, const Shape& s,
                                             int c, int z, int y, int x, int axis, float acc) {
  const long long plane = (long long)s.ny * s.nx;
  const long long base = ((long long)c * s.nz + z) * plane;
  const long long i = base + (long long)y * s.nx + x;
  const long long step = axis == 0 ? 1 : s.nx;
  const int pos = axis == 0 ? x : y;
  const int len = axis == 0 ? s.nx : s.ny;
  float out_f = 0.f, in_f = 0.f;
  if (pos + 1 < len) out_f = nan_to_num((xs[i + step] - xs[i]) - t[i]);
  if (pos > 0) in_f = nan_to_num((xs[i] - xs[i - step]) - t[i - step]);
  acc = acc + out_f;
  acc = acc - in_f;
  return acc;
}

// stitch_rigid.elastic_tile_mesh (:330-391) / elastic_tile_mesh_3d (:394-473): force on
// component c of node (z, y, x).  Order of the terms per component as in the reference:
//   c = 0: (x neighbour, cx[0]) then (y neighbour, cy[0])
//   c = 1: (y neighbour, cy[1]) then (x neighbour, cx[1])
//   c = 2: (x neighbour, cx[2]) then (y neighbour, cy[2])
__host__ __device__ inline float tile_force(const float* xs, const float* cx, const float* cy,
                                            const Shape& s, int c, int z, int y, int x) {
  float acc = 0.f;
  if (c == 1) {
    acc = spring_term(xs, cy, s, c, z, y, x, 1, acc);
    acc = spring_term(xs, cx, s, c, z, y, x, 0, acc);
  } else {
    acc = spring_term(xs, cx, s, c, z, y, x, 0, acc);
    acc = spring_term(xs, cy, s, c, z, y, x, 1, acc);
  }
  return acc;
}

// mesh.py:439: x += dt v + dt^2 / 2 a for every component of node n.
__host__ __device__ inline void advance_node(float* xs, const float* v, const float* a,
                                             const Shape& s, long long n, const Chunk& k,
                                             const State& st) {
  const long long m = s.nodes();
  float dt, hdt2;
  if (k.fire) {
    dt = st.dt;
    hdt2 = 0.5f * (dt * dt);
  } else {
    dt = k.c_dt;
    hdt2 = k.c_hdt2;
  }
  for (int c = 0; c < s.ncomp; ++c) {
    const long long i = c * m + n;
    const float t1 = dt * v[i];
    const float t2 = hdt2 * a[i];
    xs[i] = xs[i] + (t1 + t2);
  }
}

// mesh.py:440-456 for node n: new force, velocity update, FIRE mixing.  Returns the node's
// contribution to power = <a, v> (exact products, fp64).
__host__ __device__ inline double kick_node(const float* xs, float* v, float* a, const float* cx,
                                            const float* cy, const Shape& s, long long n,
                                            const Chunk& k, const State& st) {
  const long long m = s.nodes();
  const long long plane = (long long)s.ny * s.nx;
  const int z = (int)(n / plane);
  const int y = (int)((n - z * plane) / s.nx);
  const int x = (int)(n - z * plane - (long long)y * s.nx);
  float an[3], vn[3];
  for (int c = 0; c < s.ncomp; ++c) {
    const long long i = c * m + n;
    const float a_prev = a[i];
    const float a_new = tile_force(xs, cx, cy, s, c, z, y, x);
    float vv;
    if (k.fire) {
      const float hdt = 0.5f * st.dt;
      const float hdtg = hdt * k.gamma;
      const float fact0 = 1.0f / (1.0f + hdtg);
      const float fact1 = 1.0f - hdtg;
      vv = fact0 * (v[i] * fact1 + hdt * (a_prev + a_new));
    } else {
      vv = k.fact0 * (v[i] * k.fact1 + k.c_hdt * (a_prev + a_new));
    }
    a[i] = a_new;
    an[c] = a_new;
    vn[c] = vv;
  }
  double power = 0.0;
  if (k.fire) {
    float asq = an[0] * an[0], vsq = vn[0] * vn[0];
    for (int c = 1; c < s.ncomp; ++c) {
      asq = asq + an[c] * an[c];
      vsq = vsq + vn[c] * vn[c];
    }
    const float a_norm = sqrtf(asq) + 1e-6f;
    const float v_norm = sqrtf(vsq);
    for (int c = 0; c < s.ncomp; ++c) {
      power += (double)an[c] * (double)vn[c];
      vn[c] = vn[c] + st.alpha * (an[c] / a_norm * v_norm - vn[c]);
    }
  }
  for (int c = 0; c < s.ncomp; ++c) v[c * m + n] = vn[c];
  return power;
}

// mesh.py:459-492: n_pos, dt, alpha, cap and the velocity gate from the sign of power.
__host__ __device__ inline void fire_update(State* st, const Chunk& k, double power) {
  const bool pos = power >= 0.0;
  st->n_pos = pos ? st->n_pos + 1 : 0;
  if (pos) {
    if (st->n_pos > k.n_min) {
      const float grown = st->dt * k.f_inc;
      st->dt = grown < k.dt_ceiling ? grown : k.dt_ceiling;
      st->alpha = st->alpha * k.f_alpha;
    }
    if (st->n_pos > 0 && st->n_pos % k.cap_upscale_every == 0) st->cap = k.cap_scale * st->cap;
  } else {
    st->dt = st->dt * k.f_dec;
    st->alpha = k.alpha0;
  }
  st->cap = st->cap < k.final_cap ? st->cap : k.final_cap;
  st->gate = pos ? 1.0f : 0.0f;
}

// |v| of node n as relax_mesh measures it (mesh.py:584-586).
__host__ __device__ inline float speed_node(const float* v, const Shape& s, long long n) {
  const long long m = s.nodes();
  float sq = v[n] * v[n];
  for (int c = 1; c < s.ncomp; ++c) sq = sq + v[c * m + n] * v[c * m + n];
  return sqrtf(sq);
}

}  // namespace tilemesh
}  // namespace sofima
