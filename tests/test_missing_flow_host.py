"""Host logic of processor.flow.EstimateMissingFlow / EstimateFlow on CPU: the reference's own
processor tests (processor/flow_test.py:59-170) with the CUDA calculator replaced by the oracle
calculator (oracle/flow_oracle.py).  The GPU twins are in tests/test_missing_flow_gpu.py."""

import numpy as np
import pytest

from oracle import flow_oracle as fo
from sofima_b200 import compat
from sofima_b200.processor import flow as pflow


class MockVolume:
  """[C, Z, Y, X] array with the few members the processor uses (processor/flow_test.py:24-42)."""

  def __init__(self, data):
    self._data = data
    self.shape = data.shape
    self.dtype = data.dtype

  def clip_box_to_volume(self, box):
    vol_box = compat.BoundingBox(start=(0, 0, 0), size=self.volume_size)
    return box.intersection(vol_box)

  @property
  def asarray(self):
    return self._data

  @property
  def volume_size(self):
    return (self._data.shape[3], self._data.shape[2], self._data.shape[1])

  def __getitem__(self, key):
    return self._data[key]


@pytest.fixture()
def oracle_calculator(monkeypatch):
  class Calc(fo.MaskedXCorrWithStatsCalculator):
    non_spatial_flow_channels = 2

  monkeypatch.setattr(pflow.flow_field, 'JAXMaskedXCorrWithStatsCalculator', Calc)
  return Calc


def _config(**kw):
  base = dict(patch_size=16, stride=16, delta_z=1, max_delta_z=2, max_attempts=1,
              mask_configs=None, mask_only_for_patch_selection=False,
              selection_mask_configs=None, min_peak_sharpness=0.0, min_peak_ratio=0.0,
              max_magnitude=0, batch_size=10, image_volinfo='dummy_path',
              image_cache_bytes=0, mask_cache_bytes=0, search_radius=16)
  base.update(kw)
  return pflow.EstimateMissingFlow.Config(**base)


def _processor(config, vol):
  class Proc(pflow.EstimateMissingFlow):
    def _open_volume(self, path):
      return vol
  return Proc(config)


def test_process(oracle_calculator):  # processor/flow_test.py:59-124
  rng = np.random.default_rng(0)
  vol_data = rng.random((1, 10, 128, 128)).astype(np.float32)
  dx, dy = 2, 3
  prev_slice = vol_data[0, 3]
  shifted = np.zeros_like(prev_slice)
  shifted[dy:, dx:] = prev_slice[:-dy, :-dx]
  shifted[:dy, :] = rng.random((dy, 128))
  shifted[:, :dx] = rng.random((128, dx))
  vol_data[0, 5] = shifted
  proc = _processor(_config(), MockVolume(vol_data))
  box = compat.BoundingBox(start=(2, 2, 5), size=(2, 2, 1))
  out = proc.process(compat.Subvolume(np.full((2, 1, 2, 2), np.nan, dtype=np.float32), box))
  assert out.data.shape == (3, 1, 2, 2)
  assert not np.any(np.isnan(out.data)), 'Result contains NaNs'
  np.testing.assert_allclose(out.data[2, ...], 2, err_msg='delta_z incorrect')
  np.testing.assert_allclose(out.data[0, 0, 0, 0], -dx, atol=0.5)
  np.testing.assert_allclose(out.data[1, 0, 0, 0], -dy, atol=0.5)


def test_process_clipped_context(oracle_calculator):  # processor/flow_test.py:126-170
  rng = np.random.default_rng(1)
  vol_data = rng.random((1, 10, 128, 128)).astype(np.float32)
  proc = _processor(_config(max_delta_z=5), MockVolume(vol_data))
  box = compat.BoundingBox(start=(2, 2, 1), size=(2, 2, 1))
  out = proc.process(compat.Subvolume(np.full((2, 1, 2, 2), np.nan, dtype=np.float32), box))
  assert out.data.shape == (3, 1, 2, 2)
  assert np.all(np.isnan(out.data[0, ...]))
  assert np.all(np.isnan(out.data[1, ...]))
  assert out.data[2, 0, 0, 0] == 1
