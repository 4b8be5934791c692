"""Oracle comparison AT the benchmarked sizes (SURVEY 8 d): what bench.py times is checked
here value for value, not only through size-independent properties.

* BASELINE config 3 geometry: [2, 1, 2048, 2048] mesh, fixed `prev` with 1 % NaN nodes, FIRE,
  1000 steps (one chunk) then a second chunk of 200 -- final x within 1e-5 abs of
  oracle/mesh_oracle.c, identical NaN pattern, and the per-chunk (dt, alpha, n_pos, cap,
  e_kin) trace.  The C oracle does 2048^2 x 1000 steps in well under a minute on the box's
  host cores.
* BASELINE config 2 tile: one 4096^2 uint8 tile pair, patch 160 / step 40, batch 1024
  (9801 patch pairs) against oracle/flow_oracle.py: integer channels equal, NaN pattern
  equal, statistics within 2e-3 rel.
"""

import numpy as np
import pytest
import scipy.ndimage as ndi

pytestmark = pytest.mark.gpu


def _need_gpu():
  import torch
  if not torch.cuda.is_available():
    pytest.skip('needs a CUDA device')


def config3_inputs(n=2048):
  """smooth random displacement field, max-abs 8 px, 1 % NaN."""
  rng = np.random.default_rng(2)
  prev = ndi.gaussian_filter(rng.standard_normal((2, 1, n, n)), (0, 0, 16, 16))
  prev = (prev * (8.0 / np.abs(prev).max())).astype(np.float32)
  prev[:, rng.random((1, n, n)) < 0.01] = np.nan
  return np.zeros((2, 1, n, n), np.float32), prev


def test_config3_mesh_vs_c_oracle_full_size():
  _need_gpu()
  from oracle import mesh_oracle_c as mc
  from sofima_b200 import mesh
  x0, prev = config3_inputs()
  cfg = mesh.IntegrationConfig(dt=0.001, gamma=0.0, k0=0.1, k=0.1, stride=(40.0, 40.0),
                               num_iters=1000, max_iters=10000, stop_v_max=0.0, fire=True,
                               dt_max=1000.0, prefer_orig_order=True)
  # chunk 1: 1000 steps; chunk 2: 200 steps continuing from the carried (dt, alpha, cap, v)
  cfg2 = mesh.IntegrationConfig(**{**cfg.to_dict(), 'num_iters': 200})
  xg, vg = x0, np.zeros_like(x0)
  xo, vo = x0, np.zeros_like(x0)
  dt_g = dt_o = cfg.dt
  al_g = al_o = cfg.alpha
  cap_g = cap_o = cfg.start_cap
  for c in (cfg, cfg2):
    xg, vg, ag, dt_g, al_g, np_g, cap_g = mesh.velocity_verlet(xg, vg, prev, c, cap_g, dt_g, al_g)
    xo, vo, ao, dt_o, al_o, np_o, cap_o, ek_o, vmax_o = mc.velocity_verlet(
        xo, vo, prev, c, cap_o, dt_o, al_o)
    # the scalar trace of the chunk (SURVEY 8 d)
    assert np_g == np_o
    np.testing.assert_allclose([dt_g, al_g, cap_g], [dt_o, al_o, cap_o], rtol=1e-6)
    np.testing.assert_array_equal(np.isnan(xg), np.isnan(xo))
    ok = ~np.isnan(xo)
    assert np.abs(xg[ok] - xo[ok]).max() <= 1e-5       # north_star: 1e-5 abs on positions
    assert np.abs(vg[ok] - vo[ok]).max() <= 1e-5
    ek_g = float((vg[ok].astype(np.float64) ** 2).sum())
    np.testing.assert_allclose(ek_g, ek_o, rtol=1e-6)
  assert np.nanmax(np.abs(xg)) > 0.5                   # the mesh did move


def test_config3_relax_mesh_trace_vs_c_oracle():
  """relax_mesh itself (host loop + chunks) at 2048^2: t, e_kin history and x."""
  _need_gpu()
  from oracle import mesh_oracle_c as mc
  from sofima_b200 import mesh
  x0, prev = config3_inputs()
  cfg = mesh.IntegrationConfig(dt=0.001, gamma=0.0, k0=0.1, k=0.1, stride=(40.0, 40.0),
                               num_iters=250, max_iters=750, stop_v_max=0.0, fire=True,
                               dt_max=1000.0, prefer_orig_order=True)
  got, ek_g, t_g = mesh.relax_mesh(x0, prev, cfg)
  want, ek_o, t_o = mc.relax_mesh(x0, prev, cfg)
  assert t_g == t_o == 750 and len(ek_g) == len(ek_o) == 3
  np.testing.assert_allclose(ek_g, ek_o, rtol=1e-5)
  np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
  assert np.nanmax(np.abs(got - want)) <= 1e-5


def flow_tile_pair(n=4096, seed=9):
  rng = np.random.default_rng(seed)
  base = ndi.gaussian_filter(rng.standard_normal((n + 64, n + 64)).astype(np.float32), 2.0)
  base = ((base - base.min()) / (base.max() - base.min()) * 255).astype(np.uint8)
  pre = np.ascontiguousarray(base[32:32 + n, 32:32 + n])
  post = base[32 + 4:32 + 4 + n, 32 - 7:32 - 7 + n].astype(np.float32)
  post = np.clip(post + rng.normal(0, 6, post.shape), 0, 255).astype(np.uint8)
  return pre, np.ascontiguousarray(post)


def test_config2_tile_pair_vs_flow_oracle_full_size():
  _need_gpu()
  from oracle import flow_oracle as fo
  from sofima_b200 import flow_field as ff
  pre, post = flow_tile_pair()
  got = ff.JAXMaskedXCorrWithStatsCalculator().flow_field(pre, post, 160, 40, batch_size=1024)
  want = fo.MaskedXCorrWithStatsCalculator().flow_field(pre, post, 160, 40, batch_size=1024)
  assert got.shape == want.shape == (4, 99, 99)
  np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
  np.testing.assert_array_equal(got[:2], want[:2])      # integer flow vectors: equal
  # Statistics channels.  sharpness = peak / min(11 x 11 window): where the window minimum
  # is close to zero the quotient amplifies the fp32 noise of the FFT outputs without
  # bound, so it is compared as min / peak = 1 / sharpness with an ABSOLUTE tolerance
  # (1e-5 of the peak height; the two FFT implementations differ by ~1e-6 of it), and as a
  # relative value (2e-3) wherever the window minimum is at least 1 % of the peak.
  ok = ~np.isnan(want[2])
  inv_g, inv_w = 1.0 / got[2][ok].astype(np.float64), 1.0 / want[2][ok].astype(np.float64)
  np.testing.assert_allclose(inv_g, inv_w, rtol=0, atol=1e-5)
  well = ok & (np.abs(want[2]) < 100)
  np.testing.assert_allclose(got[2][well], want[2][well], rtol=2e-3)
  assert well.sum() > 0.5 * ok.sum()
  np.testing.assert_allclose(got[3][ok], want[3][ok], rtol=2e-3, atol=1e-6)
  assert (got[0] == -7).mean() > 0.99 and (got[1] == 4).mean() > 0.99
