"""Helpers for clients of the mesh processors (reference processor/client_utils.py)."""

from __future__ import annotations

import bisect
from typing import Sequence


def get_block_id(z: int, starts: Sequence[int], backward: bool) -> int:
  """Number of the block that section `z` belongs to (processor/client_utils.py:22-27).

  `starts` are the first sections of the blocks in ascending order.  Forward solves count
  the block of the first section as block 1; backward solves put a boundary section into
  the lower block.
  """
  if backward:
    return bisect.bisect_left(starts, z)
  return bisect.bisect_right(starts, z)
