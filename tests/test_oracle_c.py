"""oracle/mesh_oracle.c (fast CPU restatement) == oracle/mesh_oracle.py, bit for bit."""

import ast

import numpy as np
import pytest

from oracle import mesh_oracle as mo
from oracle import mesh_oracle_c as mc
from sofima_b200.mesh import IntegrationConfig


@pytest.mark.parametrize('tag', ['fire_poo', 'fire_cap', 'fire_drift', 'damped',
                                 'fire_nan_x'])
def test_c_oracle_matches_golden(mesh_golden, tag):
  g = mesh_golden
  cfg = IntegrationConfig(**ast.literal_eval(str(g[f'{tag}_cfg'])))
  x = g[f'{tag}_x0']
  v = np.zeros_like(x)
  prev = g[f'{tag}_prev'] if f'{tag}_prev' in g.files else None
  dt, alpha, cap = cfg.dt, cfg.alpha, cfg.start_cap
  for i in range(g[f'{tag}_xs'].shape[0]):
    x, v, _, dt_n, alpha_n, n_pos, cap_n, _, _ = mc.velocity_verlet(
        x, v, prev, cfg, cap, dt, alpha)
    if cfg.fire:
      dt, alpha, cap = dt_n, alpha_n, cap_n
      np.testing.assert_allclose([dt, alpha, n_pos, cap], g[f'{tag}_scalars'][i],
                                 rtol=1e-6)
    tol = 5e-5 if cfg.remove_drift else 0.0
    np.testing.assert_allclose(x, g[f'{tag}_xs'][i], rtol=0, atol=tol)
    np.testing.assert_allclose(v, g[f'{tag}_vs'][i], rtol=0, atol=tol)


def test_c_oracle_matches_numpy_oracle_on_random_mesh():
  rng = np.random.default_rng(3)
  shape = (2, 2, 37, 53)
  prev = (rng.standard_normal(shape) * 5).astype(np.float32)
  prev[rng.random(shape) < 0.02] = np.nan
  cfg = IntegrationConfig(dt=0.001, gamma=0.0, k0=0.1, k=0.1, stride=(40.0, 40.0),
                          num_iters=80, max_iters=240, stop_v_max=0.0, fire=True,
                          dt_max=1000.0, prefer_orig_order=True)
  want, ek_w, t_w = mo.relax_mesh(np.zeros(shape, np.float32), prev, cfg)
  got, ek_g, t_g = mc.relax_mesh(np.zeros(shape, np.float32), prev, cfg)
  assert t_w == t_g
  np.testing.assert_array_equal(got, want)
  np.testing.assert_allclose(ek_g, ek_w, rtol=1e-12)
