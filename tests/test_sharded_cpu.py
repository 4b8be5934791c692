"""N > 1 path on CPU: world_size-2/3 gloo runs of the row-sharded mesh protocol
(emulated with the oracle arithmetic) must reproduce the single-domain oracle."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import mesh_oracle as mo
from sofima_b200.mesh import IntegrationConfig
from sofima_b200.mesh_sharded import partition_rows


def test_partition_rows():
  assert partition_rows(2048, 8) == [(256 * r, 256 * (r + 1)) for r in range(8)]
  assert partition_rows(205, 3) == [(0, 96), (96, 160), (160, 205)]
  assert partition_rows(100, 1) == [(0, 100)]
  for ny, n in ((2048, 4), (1000, 7), (97, 3), (64, 2)):
    rows = partition_rows(ny, n)
    assert rows[0][0] == 0 and rows[-1][1] == ny
    for (a, b), (c, d) in zip(rows, rows[1:]):
      assert b == c and (b - a) % 32 == 0 and b > a
    assert rows[-1][1] > rows[-1][0]
  with pytest.raises(ValueError):
    partition_rows(40, 3)


def _free_port():
  with socket.socket() as s:
    s.bind(('127.0.0.1', 0))
    return s.getsockname()[1]


def _worker(rank, world, port, shape, seed, out_dir):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    from tests.sharded_reference import sharded_relax
    x0, prev, cfg = _problem(shape, seed)
    x, e_kin, t, (y0, y1) = sharded_relax(x0, prev, cfg)
    np.savez(os.path.join(out_dir, f'rank{rank}.npz'), x=x, e_kin=e_kin, t=t, y0=y0, y1=y1)
  finally:
    dist.destroy_process_group()


def _problem(shape, seed):
  rng = np.random.default_rng(seed)
  prev = (rng.standard_normal(shape) * 4).astype(np.float32)
  prev[rng.random(shape) < 0.02] = np.nan
  cfg = IntegrationConfig(dt=0.001, gamma=0.0, k0=0.1, k=0.1, stride=(40.0, 40.0),
                          num_iters=40, max_iters=120, stop_v_max=0.0, fire=True,
                          dt_max=1000.0, prefer_orig_order=True)
  return np.zeros(shape, np.float32), prev, cfg


@pytest.mark.parametrize('world,shape', [(2, (2, 2, 70, 23)), (3, (2, 1, 100, 17))])
def test_sharded_protocol_matches_single_domain(tmp_path, world, shape):
  port = _free_port()
  mp.spawn(_worker, args=(world, port, shape, 11, str(tmp_path)), nprocs=world, join=True)
  x0, prev, cfg = _problem(shape, 11)
  want, ek_w, t_w = mo.relax_mesh(x0, prev, cfg)
  got = np.empty_like(want)
  for r in range(world):
    d = np.load(tmp_path / f'rank{r}.npz')
    got[:, :, int(d['y0']):int(d['y1'])] = d['x']
    assert int(d['t']) == t_w
    np.testing.assert_allclose(d['e_kin'], ek_w, rtol=1e-12)
  np.testing.assert_array_equal(got, want)  # one halo row of (x, v, a) is sufficient
