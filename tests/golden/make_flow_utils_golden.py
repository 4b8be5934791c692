"""Golden vectors for sofima_b200.flow_utils from the reference's own flow_utils.py
(pure NumPy / SciPy, importable without JAX):

  python tests/golden/make_flow_utils_golden.py      # needs /root/reference
"""
import importlib.util
import os

import numpy as np
import scipy.ndimage as ndi

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
  spec = importlib.util.spec_from_file_location('ref_flow_utils', '/root/reference/flow_utils.py')
  ref = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(ref)
  rng = np.random.default_rng(12)
  out = {}
  f4 = np.zeros((4, 3, 30, 34), np.float32)
  f4[:2] = ndi.gaussian_filter(rng.standard_normal((2, 3, 30, 34)), (0, 0, 2, 2)) * 20
  f4[:2] += (rng.random((2, 3, 30, 34)) < 0.03) * rng.standard_normal((2, 3, 30, 34)) * 30
  f4[2] = rng.random((3, 30, 34)) * 3 - 0.5
  f4[3] = np.where(rng.random((3, 30, 34)) < 0.3, 0.0, rng.random((3, 30, 34)) * 3)
  f4[:, 1, 4:9, 5:12] = np.nan
  out['clean_in'] = f4
  out['clean_out'] = ref.clean_flow(f4.copy(), min_peak_ratio=1.4, min_peak_sharpness=0.6,
                                    max_magnitude=25.0, max_deviation=4.0)
  out['clean_out_nodev'] = ref.clean_flow(f4.copy(), 1.4, 0.6, 0, 0)
  out['clean2_out'] = ref.clean_flow(f4[:2].copy(), 0, 0, 25.0, 4.0)
  f5 = np.concatenate([f4[:2, :2, :12, :14],
                       rng.standard_normal((1, 2, 12, 14)).astype(np.float32),
                       f4[2:, :2, :12, :14]])
  out['clean3d_in'] = f5
  out['clean3d_out'] = ref.clean_flow(f5.copy(), 1.4, 0.6, 25.0, 4.0, dim=3)
  a = out['clean_out'].copy()
  b = ref.clean_flow((f4 * np.float32(1.05)).copy(), 1.2, 0.3, 30.0, 6.0)
  c = np.zeros_like(a)
  c[:] = 1.5
  out['rec_a'], out['rec_b'], out['rec_c'] = a, b, c
  out['rec_out'] = ref.reconcile_flows([a.copy(), b.copy(), c.copy()], max_gradient=5.0,
                                       max_deviation=3.0, min_patch_size=12)
  out['rec_out_nofilter'] = ref.reconcile_flows([a.copy(), b.copy()], 0, 0, 0)
  a3 = np.concatenate([a, np.full((1,) + a.shape[1:], 1, np.float32)])
  a3[2][np.isnan(a3[0])] = np.nan
  b3 = np.concatenate([b, rng.integers(1, 4, (1,) + b.shape[1:]).astype(np.float32)])
  out['rec3_a'], out['rec3_b'] = a3, b3
  out['rec3_out'] = ref.reconcile_flows([a3.copy(), b3.copy()], 5.0, 3.0, 12, min_delta_z=2)
  np.savez_compressed(os.path.join(HERE, 'flow_utils_golden.npz'), **out)


if __name__ == '__main__':
  main()
