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


def test_committed_bench_lines_follow_the_contract():
  """The last bench lines measured on a B200 (profiles/) carry every key the driver's
  contract names -- a guard against dropping one while editing bench.py."""
  import os
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  line = json.load(open(os.path.join(root, 'profiles', 'bench_r2_default.json')))
  for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step',
              'higher_is_better', 'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'clocks',
              'e2e', 'gpu_launches', 'roofline', 'cpu_baseline'):
    assert key in line, key
  assert line['metric'] == line['unit'] == 'patch-pairs/s' and line['vs_baseline'] is None
  assert 'workload' in line['config'] and line['gpu_launches'] > 0
  assert set(line['e2e']) >= {'value', 'unit', 'h2d_bytes_per_step', 'd2h_bytes_per_step'}
  assert line['e2e']['h2d_bytes_per_step'] > 0 and line['e2e']['value'] < line['value']
  assert set(line['roofline']) >= {'bound', 'achieved', 'peak', 'unit', 'frac', 'traffic'}
  assert set(line['cpu_baseline']) >= {'value', 'unit', 'cores', 'kind', 'sample'}
  assert set(line['clocks']) >= {'sm_mhz', 'sm_max_mhz', 'reasons'}
  mesh = line['mesh']
  assert mesh['roofline']['bound'] == 'hbm' and 0 < mesh['roofline']['frac'] < 1
  assert mesh['cpu_baseline']['kind'] == 'port'
  assert line['roofline']['bound'] == 'hbm' and 0 < line['roofline']['frac'] < 1
  assert line['roofline']['traffic'] > line['roofline']['algorithmic_io_bytes_per_step']
  assert line['sustained']['seconds'] >= 1.99  # (bench.py now aims at 2.1 s)
  ref = json.load(open(os.path.join(root, 'profiles', 'bench_r2_reference_arm.json')))
  assert ref['impl'] == 'reference' and ref['metric'] == line['metric']
  assert ref['e2e']['h2d_bytes_per_step'] == 0 and ref['cpu_baseline']['value'] == ref['value']
  # multi-GPU lines carry the in-run parity record of the sharded mesh
  for n in (2, 4, 8):
    multi = json.load(open(os.path.join(root, 'profiles', f'bench_r2_n{n}_gpus.json')))
    assert multi['n_gpus'] == n and multi['mesh']['parity']['bit_identical'] is True
    assert multi['mesh']['parity']['max_abs_err'] == 0.0
