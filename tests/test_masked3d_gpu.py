"""Masked 3-d correlation (csrc/flow3d_masked.cuh, reference flow_field.py:91-155 with
dim = 3) in a child process.  Passed on a B200 at the end of round 1 (GPUTEST_r01: xpassed), so
the path is on by default.  The oracle side is pinned on the reference's own NumPy branch
(tests/test_oracle_flow.py::test_reference_numpy_branch_golden)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden', 'xcorr_numpy_golden.npz')

CHILD = r'''
import sys
import numpy as np
sys.path.insert(0, sys.argv[1])
from sofima_b200 import flow_field as ff
g = np.load(sys.argv[2])
out = {}
out['xc'] = ff.masked_xcorr(g['xn_3dm_prev'], g['xn_3dm_curr'], g['xn_3dm_pm'], g['xn_3dm_cm'],
                            dim=3)
# unequal sizes, one-sided mask
rng = np.random.default_rng(8)
prev = (rng.standard_normal((10, 12, 9)) * 7).astype(np.float32)
curr = (rng.standard_normal((5, 6, 7)) * 7).astype(np.float32)
pm = rng.random(prev.shape) > 0.7
out['prev'], out['curr'], out['pm'] = prev, curr, pm
out['xc_one'] = ff.masked_xcorr(prev, curr, pm, None, dim=3)
# the whole flow_field call with masks that take part in the correlation
vol = rng.standard_normal((30, 34, 36)).astype(np.float32)
import scipy.ndimage as ndi
vol = ndi.gaussian_filter(vol, 1.2)
pre, post = vol[2:26, 3:29, 4:32].copy(), vol[3:27, 1:27, 6:34].copy()
m = np.zeros(pre.shape, bool); m[:6, :8, :10] = True
out['pre'], out['post'], out['m'] = pre, post, m
out['flow'] = ff.JAXMaskedXCorrWithStatsCalculator().flow_field(
    pre, post, (12, 14, 16), (4, 4, 4), pre_mask=m, post_mask=m, batch_size=8)
np.savez(sys.argv[3], **out)
'''


def test_masked_3d_correlation(tmp_path):
  import torch
  if not torch.cuda.is_available():
    pytest.skip('needs a CUDA device')
  from oracle import flow_oracle as fo
  res = tmp_path / 'out.npz'
  env = dict(os.environ)
  proc = subprocess.run([sys.executable, '-c', CHILD, ROOT, GOLDEN, str(res)], env=env,
                        capture_output=True, text=True, timeout=300)
  assert proc.returncode == 0, proc.stderr[-2000:]
  out, g = np.load(res), np.load(GOLDEN)
  np.testing.assert_allclose(out['xc'], g['xn_3dm_masked'], rtol=0, atol=2e-4)
  want = fo.masked_xcorr(out['prev'], out['curr'], out['pm'], None, dim=3)
  np.testing.assert_allclose(out['xc_one'], want, rtol=0, atol=2e-4)
  want = fo.MaskedXCorrWithStatsCalculator().flow_field(
      out['pre'], out['post'], (12, 14, 16), (4, 4, 4), pre_mask=out['m'], post_mask=out['m'],
      batch_size=8)
  got = out['flow']
  np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
  np.testing.assert_array_equal(got[:3], want[:3])
