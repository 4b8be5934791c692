/*
 * Plain-C restatement of the reference 2-d spring-mesh integrator (TEST ORACLE and
 * CPU baseline, never shipped): mesh.inplane_force (reference mesh.py:42-169) +
 * mesh.velocity_verlet (mesh.py:371-521), fp32 in the reference's association
 * order.  Must be compiled with -ffp-contract=off (no FMA contraction) and without
 * -ffast-math; see oracle/Makefile.  It is validated bit-for-bit against
 * oracle/mesh_oracle.py (tests/test_oracle_c.py), which in turn is pinned against
 * the reference's own source.
 *
 * Global reductions (power = vdot(a, v), drift means, e_kin) accumulate in double,
 * like the NumPy oracle.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  double dt, gamma, k0, k;
  double stride[3];
  int32_t num_iters;
  int32_t fire;
  double f_alpha, f_inc, f_dec, alpha;
  int32_t n_min;
  double dt_max;
  double start_cap, final_cap, cap_scale;
  int32_t cap_upscale_every;
  int32_t prefer_orig_order;
  int32_t remove_drift;
} oracle_config; /* same layout as sofima_integration_config */

static inline float zero_nonfinite(float f) { return isfinite(f) ? f : 0.0f; }
static inline float sgn(float x) {
  return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : (x == 0.0f ? 0.0f : x));
}
static inline float nan_to_num(float d) {
  if (isnan(d)) return 0.0f;
  if (isinf(d)) return d > 0 ? 3.402823466e+38f : -3.402823466e+38f;
  return d;
}

typedef struct {
  int dx, dy;
  float l0x, l0y, l0, neg_k;
} link_t;

static inline void link_force(const float* x0, const float* x1, long to, long from,
                              const link_t* L, int poo, float* f0, float* f1) {
  const float d0 = (x0[to] - x0[from]) + L->l0x;
  const float d1 = (x1[to] - x1[from]) + L->l0y;
  const float sq = d0 * d0 + d1 * d1;
  const float len = sqrtf(sq);
  const float q = L->l0 / len;
  float t0 = q, t1 = q;
  if (poo) {
    if (L->dx != 0) t0 = ((float)L->dx * sgn(d0)) * q;
    if (L->dy != 0) t1 = ((float)L->dy * sgn(d1)) * q;
  }
  *f0 = zero_nonfinite((L->neg_k * (1.0f - t0)) * d0);
  *f1 = zero_nonfinite((L->neg_k * (1.0f - t1)) * d1);
}

/* a = inplane_force(x) + clip(-k0 * nan_to_num(x - prev), -cap, cap) on one plane
 * set [2][nz][ny][nx]; prev may be NULL. */
static void total_force(const float* x, const float* prev, float* a, long nz, long ny,
                        long nx, const link_t* links, int poo, float neg_k0, float cap) {
  const long plane = ny * nx, n = nz * plane;
#pragma omp parallel for collapse(2) schedule(static)
  for (long z = 0; z < nz; ++z) {
    for (long y = 0; y < ny; ++y) {
      const float* x0 = x + z * plane;
      const float* x1 = x + n + z * plane;
      for (long xx = 0; xx < nx; ++xx) {
        const long i = y * nx + xx;
        float p[4][2], m[4][2];
        for (int l = 0; l < 4; ++l) {
          const link_t* L = &links[l];
          p[l][0] = p[l][1] = m[l][0] = m[l][1] = 0.0f;
          /* link ending here: from = this - dir */
          const long fy = y - L->dy, fx = xx - L->dx;
          if (fy >= 0 && fy < ny && fx >= 0 && fx < nx)
            link_force(x0, x1, i, fy * nx + fx, L, poo, &p[l][0], &p[l][1]);
          /* link starting here: to = this + dir */
          const long ty = y + L->dy, tx = xx + L->dx;
          if (ty >= 0 && ty < ny && tx >= 0 && tx < nx)
            link_force(x0, x1, ty * nx + tx, i, L, poo, &m[l][0], &m[l][1]);
        }
        for (int c = 0; c < 2; ++c) {
          /* mesh.py:169: f1p + f2p + f3p + f4p - f1n - f2n - f3n - f4n */
          float s = p[0][c] + p[1][c];
          s = s + p[2][c];
          s = s + p[3][c];
          s = s - m[0][c];
          s = s - m[1][c];
          s = s - m[2][c];
          s = s - m[3][c];
          const long gi = c * n + z * plane + i;
          if (prev) {
            const float d = nan_to_num(x[gi] - prev[gi]);
            float pull = neg_k0 * d;
            pull = fminf(fmaxf(pull, -cap), cap);
            s = s + pull;
          }
          a[gi] = s;
        }
      }
    }
  }
}

/* One velocity_verlet call (mesh.py:371-521) on x, v [2][nz][ny][nx], in place.
 * a: output/scratch of the same size.  FIRE scalars in/out.  Returns 0. */
int oracle_mesh_chunk(float* x, float* v, float* a, const float* prev, long nz, long ny,
                      long nx, const oracle_config* cfg, float* dt_io, float* alpha_io,
                      float* cap_io, int32_t* n_pos_out, double* e_kin, float* v_max) {
  const long n = nz * ny * nx, n2 = 2 * n;
  link_t links[4];
  const float sx = (float)cfg->stride[0], sy = (float)cfg->stride[1];
  const float l0d = (float)sqrt(cfg->stride[0] * cfg->stride[0] + cfg->stride[1] * cfg->stride[1]);
  const float k_ax = (float)cfg->k, k_diag = k_ax / sqrtf(2.0f);
  const int dirs[4][2] = {{1, 0}, {0, 1}, {1, 1}, {-1, 1}};
  for (int l = 0; l < 4; ++l) {
    links[l].dx = dirs[l][0];
    links[l].dy = dirs[l][1];
    links[l].l0x = (float)dirs[l][0] * sx;
    links[l].l0y = (float)dirs[l][1] * sy;
    links[l].l0 = l == 0 ? sx : (l == 1 ? sy : l0d);
    links[l].neg_k = -(l < 2 ? k_ax : k_diag);
  }
  const int poo = cfg->prefer_orig_order != 0;
  const float neg_k0 = -(float)cfg->k0;
  float* a_prev = (float*)malloc(sizeof(float) * (size_t)(n2 > 0 ? n2 : 1));
  if (!a_prev) return 4;

  float dt = *dt_io, alpha = *alpha_io, cap = *cap_io;
  int n_pos = 0;
  total_force(x, prev, a, nz, ny, nx, links, poo, neg_k0, cap);

  if (!cfg->fire) {
    const double d = cfg->dt, g = cfg->gamma;
    const float c_dt = (float)d, c_hdt2 = (float)(0.5 * d * d);
    const float fact0 = (float)(1.0 / (1.0 + 0.5 * d * g)), fact1 = (float)(1.0 - 0.5 * d * g);
    const float c_hdt = (float)(0.5 * d);
    for (int it = 0; it < cfg->num_iters; ++it) {
#pragma omp parallel for schedule(static)
      for (long i = 0; i < n2; ++i) {
        x[i] = x[i] + (c_dt * v[i] + c_hdt2 * a[i]);
        a_prev[i] = a[i];
      }
      total_force(x, prev, a, nz, ny, nx, links, poo, neg_k0, cap);
#pragma omp parallel for schedule(static)
      for (long i = 0; i < n2; ++i) v[i] = fact0 * (v[i] * fact1 + c_hdt * (a_prev[i] + a[i]));
    }
    *n_pos_out = -1;
  } else {
    const float dt_ceiling = (float)(cfg->dt_max * cfg->dt);
    const float gamma = (float)cfg->gamma;
    for (int it = 0; it < cfg->num_iters; ++it) {
      const float hdt2 = 0.5f * (dt * dt);
#pragma omp parallel for schedule(static)
      for (long i = 0; i < n2; ++i) {
        x[i] = x[i] + (dt * v[i] + hdt2 * a[i]);
        a_prev[i] = a[i];
      }
      total_force(x, prev, a, nz, ny, nx, links, poo, neg_k0, cap);
      const float hdt = 0.5f * dt, hdtg = hdt * gamma;
      const float fact0 = 1.0f / (1.0f + hdtg), fact1 = 1.0f - hdtg;
      double power = 0.0, sx0 = 0.0, sx1 = 0.0, sv0 = 0.0, sv1 = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : power)
      for (long i = 0; i < n; ++i) {
        float v0 = fact0 * (v[i] * fact1 + hdt * (a_prev[i] + a[i]));
        float v1 = fact0 * (v[i + n] * fact1 + hdt * (a_prev[i + n] + a[i + n]));
        const float a0 = a[i], a1 = a[i + n];
        const float a_norm = sqrtf(a0 * a0 + a1 * a1) + 1e-6f;
        const float v_norm = sqrtf(v0 * v0 + v1 * v1);
        power += (double)a0 * (double)v0 + (double)a1 * (double)v1;
        v0 = v0 + alpha * (a0 / a_norm * v_norm - v0);
        v1 = v1 + alpha * (a1 / a_norm * v_norm - v1);
        v[i] = v0;
        v[i + n] = v1;
      }
      const int pos = power >= 0.0;
      n_pos = pos ? n_pos + 1 : 0;
      const float alpha_used = alpha;
      (void)alpha_used;
      if (pos) {
        if (n_pos > cfg->n_min) {
          dt = fminf(dt * (float)cfg->f_inc, dt_ceiling);
          alpha = alpha * (float)cfg->f_alpha;
        }
        if (n_pos > 0 && (n_pos % cfg->cap_upscale_every) == 0) cap = (float)cfg->cap_scale * cap;
      } else {
        dt = dt * (float)cfg->f_dec;
        alpha = (float)cfg->alpha;
      }
      cap = fminf(cap, (float)cfg->final_cap);
      if (!pos) {
#pragma omp parallel for schedule(static)
        for (long i = 0; i < n2; ++i) v[i] = v[i] * 0.0f;
      }
      if (cfg->remove_drift) {
#pragma omp parallel for schedule(static) reduction(+ : sx0, sx1, sv0, sv1)
        for (long i = 0; i < n; ++i) {
          sx0 += x[i]; sx1 += x[i + n]; sv0 += v[i]; sv1 += v[i + n];
        }
        const float mx0 = (float)(sx0 / (double)n), mx1 = (float)(sx1 / (double)n);
        const float mv0 = (float)(sv0 / (double)n), mv1 = (float)(sv1 / (double)n);
#pragma omp parallel for schedule(static)
        for (long i = 0; i < n; ++i) {
          x[i] -= mx0; x[i + n] -= mx1; v[i] -= mv0; v[i + n] -= mv1;
        }
      }
    }
    *n_pos_out = n_pos;
  }
  double ek = 0.0;
  float vm = 0.0f;
#pragma omp parallel for schedule(static) reduction(+ : ek) reduction(max : vm)
  for (long i = 0; i < n; ++i) {
    const float mag = sqrtf(v[i] * v[i] + v[i + n] * v[i + n]);
    ek += (double)(mag * mag);
    vm = fmaxf(vm, mag);
  }
  *e_kin = ek;
  *v_max = vm;
  *dt_io = dt;
  *alpha_io = alpha;
  *cap_io = cap;
  free(a_prev);
  return 0;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
