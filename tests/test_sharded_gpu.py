"""Row-sharded mesh on >= 2 GPUs (SURVEY 8 e): spawns one rank per visible GPU with
torch.distributed.run and requires the gathered slabs to equal the single-GPU solve BIT FOR
BIT (tests/multi/run_sharded_mesh.py does the work and exits non-zero otherwise).  Skipped on
a one-GPU box; bench.py repeats the same comparison inside every N > 1 run and reports it as
`mesh.parity`."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
  import torch
  return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize('ny,nx,iters', [(512, 384, 200), (2048, 2048, 100)])
def test_sharded_mesh_equals_single_gpu(ny, nx, iters):
  n = _ngpus()
  if n < 2:
    pytest.skip('needs at least two GPUs')
  n = min(n, 8)
  if ny // 32 < n:
    pytest.skip('mesh too small for this many ranks')
  cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={n}',
         '--master-addr', '127.0.0.1', '--master-port', '29517',
         os.path.join(ROOT, 'tests', 'multi', 'run_sharded_mesh.py'), str(ny), str(nx), str(iters)]
  proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
  assert proc.returncode == 0, (proc.stdout[-1500:], proc.stderr[-1500:])
  line = [l for l in proc.stdout.splitlines() if l.startswith('{')][-1]
  rec = json.loads(line)
  assert rec['world'] == n and rec['ok'] and rec['max_abs_err'] == 0.0
