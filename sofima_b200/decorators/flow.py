"""Flow decorators on the B200 backend (reference decorators/flow.py).

The reference wraps TensorStore `virtual_chunked` views (`OptimFlow` :131,
`MeshRelaxFlowFilter` :108).  TensorStore and gin are not available in this image,
so this module provides the chunk functions those decorators apply (`optim_flow`,
`mesh_relax_flow` reach the hot path; `clean_flow`, `reconcile_flow` are the host
filters of `CleanFlowFilter` :47 / `ReconcileFlowFilter` :358) with the reference's
argument conventions; the
TensorStore wrappers raise a clear ImportError when constructed without it.
"""

from __future__ import annotations

from typing import Sequence

import numpy as np

from .. import flow_field
from .. import flow_utils
from .. import mesh


def clean_flow(flow: np.ndarray, **filter_args) -> np.ndarray:
  """Chunk function of `CleanFlowFilter` (decorators/flow.py:38-43): [dim + 2, ...]
  flow with singleton non-spatial axes -> [dim, ...] cleaned vectors."""
  final_shape = list(flow.shape)
  final_shape[0] -= 2
  return flow_utils.clean_flow(flow.squeeze(), dim=flow.shape[0] - 2,
                               **filter_args).reshape(final_shape)


def reconcile_flow(flow: np.ndarray, **filter_args) -> np.ndarray:
  """Chunk function of `ReconcileFlowFilter` (decorators/flow.py:349-354)."""
  return flow_utils.reconcile_flows([flow.squeeze()], **filter_args).reshape(flow.shape)


def mesh_relax_flow(flow: np.ndarray, **filter_args) -> np.ndarray:
  """Chunk function of `MeshRelaxFlowFilter` (decorators/flow.py:96-105).

  `flow` is a [2 or 3, ...] flow field with singleton non-spatial axes; it becomes
  the `prev` target of a mesh that starts at zero and is relaxed with
  `IntegrationConfig(**filter_args)`.
  """
  cfg = mesh.IntegrationConfig(**filter_args)
  target = flow.squeeze()
  x0 = np.zeros_like(target)
  ncomp = flow.shape[0]
  if ncomp == 2:
    if target.ndim == 3:  # [2, y, x] -> [2, 1, y, x]
      res = mesh.relax_mesh(x0[:, None], target[:, None], cfg)
    else:
      res = mesh.relax_mesh(x0, target, cfg)
  elif ncomp == 3:
    res = mesh.relax_mesh(x0, target, cfg, mesh_force=mesh.elastic_mesh_3d)
  else:
    raise ValueError(f'`num_spatial_dim` must be 2 or 3 but is {ncomp}.')
  return np.asarray(res[0]).reshape(flow.shape)


def optim_flow(pre_image: np.ndarray, post_image: np.ndarray, patch_zyx: Sequence[int],
               step_zyx: Sequence[int], batch_size: int = 1024, pre_mask=None,
               post_mask=None, **flow_args) -> np.ndarray:
  """Chunk function of `OptimFlow` (decorators/flow.py:279-290): float32 images in
  ([z,] y, x) order -> flow field from post to pre."""
  calc = flow_field.JAXMaskedXCorrWithStatsCalculator()
  return calc.flow_field(
      pre_image=np.asarray(pre_image, dtype=np.float32),
      post_image=np.asarray(post_image, dtype=np.float32), pre_mask=pre_mask,
      post_mask=post_mask, patch_size=tuple(patch_zyx), step=tuple(step_zyx),
      batch_size=batch_size, **flow_args)


class _NeedsTensorStore:

  def __init__(self, *args, **kwargs):
    del args, kwargs
    try:
      import tensorstore  # noqa: F401  pylint: disable=unused-import
    except ImportError as e:
      raise ImportError(
          f'{type(self).__name__} wraps TensorStore virtual_chunked views; '
          'tensorstore is not installed. Use mesh_relax_flow / optim_flow on '
          'NumPy chunks instead.') from e
    raise NotImplementedError('TensorStore wrapper not built in this round')


class MeshRelaxFlowFilter(_NeedsTensorStore):
  """decorators/flow.py:108-128."""


class OptimFlow(_NeedsTensorStore):
  """decorators/flow.py:131-355."""


class CleanFlowFilter(_NeedsTensorStore):
  """decorators/flow.py:47-86."""


class ReconcileFlowFilter(_NeedsTensorStore):
  """decorators/flow.py:358-369."""
