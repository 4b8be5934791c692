"""bench.py helpers that do not need a GPU."""
import json

import bench


def test_flow_kernel_rooflines():
  k = {'flow_cols': 3.0, 'flow_rows_inv': 2.0, 'flow_rowspec': 0.5, 'flow_peak2': 0.3}
  r = bench.flow_kernel_rooflines(k, 9801, 4096, 160, 40, 6650.0)
  assert set(r) == {'flow_cols', 'flow_rows_inv', 'flow_rowspec'}
  # one pair: 319 x 161 complex64 product spectra in, 319^2 fp32 image out
  assert r['flow_rows_inv']['algorithmic_bytes_per_step'] == 9801 * (319 * 161 * 8 + 319 * 319 * 4)
  assert r['flow_cols']['algorithmic_bytes_per_step'] == 9801 * (2 * 160 * 161 * 8 + 319 * 161 * 8)
  for v in r.values():
    assert 0 < v['frac'] < 1 and v['unit'] == 'GB/s'
  json.dumps(r)
  assert bench.flow_kernel_rooflines({}, 9801, 4096, 160, 40, 6650.0) == {}


def test_peaks_fallback(tmp_path, monkeypatch):
  monkeypatch.setattr(bench, 'ROOT', str(tmp_path))
  assert bench._peaks()['source'].startswith('fallback')
  (tmp_path / 'MEASURED_PEAKS.json').write_text('{"hbm_gbs": 6500.0}')   # incomplete file
  assert bench._peaks()['source'].startswith('fallback')
  (tmp_path / 'MEASURED_PEAKS.json').write_text(
      '{"hbm_gbs": 6500.0, "bf16_tflops_sustained": 1400.0}')
  p = bench._peaks()
  assert p['hbm_gbs'] == 6500.0 and p['tflops'] == 1400.0 and p['source'].startswith('measured')
