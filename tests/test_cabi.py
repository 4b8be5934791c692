"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and
exports every symbol that include/sofima_b200.h declares; the ctypes mirrors match
the C structs; the product fails loudly (no CPU fallback) and never imports the
oracle."""

import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'sofima_b200.h')


def _declared_functions():
  text = open(HEADER).read()
  text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
  return sorted(set(re.findall(r'\b(sofima_[a-z0-9_]+)\s*\(', text)))


@pytest.fixture(scope='module')
def native():
  from sofima_b200 import _native
  _native.build()
  return _native


def test_library_exports_every_declared_symbol(native):
  declared = _declared_functions()
  assert len(declared) >= 12
  handle = ctypes.CDLL(native.LIB_PATH)
  missing = [name for name in declared if not hasattr(handle, name)]
  assert not missing, f'not exported: {missing}'
  assert set(declared) == set(native.EXPORTED_SYMBOLS)
  assert native.lib().sofima_abi_version() == 1


def test_ctypes_structs_match_header(native, tmp_path):
  src = tmp_path / 'sizes.c'
  src.write_text(
      '#include <stdio.h>\n#include <stddef.h>\n#include "sofima_b200.h"\n'
      'int main(void) { printf("%zu %zu %zu %zu %zu %zu\\n", '
      'sizeof(sofima_integration_config), sizeof(sofima_mesh_shape), '
      'sizeof(sofima_mesh_state), sizeof(sofima_xcorr_params), '
      'offsetof(sofima_integration_config, cap_upscale_every), '
      'offsetof(sofima_xcorr_params, peak_radius)); return 0; }\n')
  exe = tmp_path / 'sizes'
  subprocess.run(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)],
                 check=True)
  got = [int(v) for v in subprocess.run([str(exe)], capture_output=True,
                                        check=True).stdout.split()]
  want = [ctypes.sizeof(native.IntegrationConfigPod), ctypes.sizeof(native.MeshShape),
          ctypes.sizeof(native.MeshState), ctypes.sizeof(native.XcorrParams),
          native.IntegrationConfigPod.cap_upscale_every.offset,
          native.XcorrParams.peak_radius.offset]
  assert got == want


def test_library_is_sm100a_cuda(native):
  out = subprocess.run(['cuobjdump', '-lelf', native.LIB_PATH], capture_output=True,
                       text=True).stdout
  assert 'sm_100a' in out


def test_no_cpu_fallback_and_no_oracle_in_product(native):
  import torch
  pkg = os.path.join(ROOT, 'sofima_b200')
  for dirpath, _, files in os.walk(pkg):
    for f in files:
      if f.endswith(('.py', '.cu', '.cuh', '.h')):
        assert 'oracle' not in open(os.path.join(dirpath, f)).read().replace(
            'oracle/mesh_oracle.py for the line-by-line restatement', ''), f
  if torch.cuda.is_available():
    pytest.skip('GPU present: the loud-failure path is exercised on CPU boxes only')
  from sofima_b200 import flow_field, mesh
  cfg = mesh.IntegrationConfig(dt=0.01, gamma=0.0, k0=0.1, k=0.1, stride=(10, 10),
                               num_iters=1, max_iters=1, stop_v_max=0.1)
  x = np.zeros((2, 1, 4, 4))
  with pytest.raises(native.NativeError):
    mesh.relax_mesh(x, x, cfg)
  with pytest.raises(native.NativeError):
    mesh.inplane_force(x, 0.1, (10, 10))
  img = np.zeros((64, 64), np.uint8)
  with pytest.raises(native.NativeError):
    flow_field.JAXMaskedXCorrWithStatsCalculator().flow_field(img, img, 32, 16)


def test_integration_config_api():
  from sofima_b200 import mesh
  cfg = mesh.IntegrationConfig(dt=0.001, gamma=0.0, k0=0.01, k=0.1, stride=[40, 40],
                               num_iters=1000, max_iters=100000, stop_v_max=0.005,
                               dt_max=1000, start_cap=0.01, final_cap=10,
                               prefer_orig_order=True)
  assert cfg.stride == (40, 40) and hash(cfg) == hash(cfg)  # frozen + hashable
  assert cfg.fire and cfg.f_alpha == 0.99 and cfg.cap_upscale_every == 100
  again = mesh.IntegrationConfig.from_dict(cfg.to_dict())
  assert again == cfg
  assert mesh.IntegrationConfig.from_json(cfg.to_json()) == cfg
  with pytest.raises(Exception):
    cfg.dt = 1.0
  assert len(mesh.MESH_LINK_DIRECTIONS) == 13
