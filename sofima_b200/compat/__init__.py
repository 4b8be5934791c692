"""Minimal stand-ins for the parts of the un-vendored `connectomics` package that
the SOFIMA plugin classes build on (reference setup.cfg:23; the package is not
installable offline).  If the real package is importable it is used instead.

Only what `processor.flow.EstimateFlow` / `processor.mesh.RelaxMesh` touch is
provided: BoundingBox, Subvolume, SubvolumeProcessor (+ SuggestedXyz) and no-op
counters.
"""

from __future__ import annotations

import contextlib
import dataclasses
from typing import Sequence

import numpy as np

try:  # pragma: no cover - not available in this image
  from connectomics.common import bounding_box as _bb
  from connectomics.volume import subvolume as _sv
  from connectomics.volume import subvolume_processor as _sp
  BoundingBox = _bb.BoundingBox
  Subvolume = _sv.Subvolume
  SubvolumeProcessor = _sp.SubvolumeProcessor
  SuggestedXyz = _sp.SuggestedXyz
  HAVE_CONNECTOMICS = True
except ImportError:
  HAVE_CONNECTOMICS = False

  class BoundingBox:
    """Axis-aligned integer box; `start` / `size` / `end` are in XYZ order."""

    def __init__(self, start=None, size=None, end=None):
      if start is not None and size is not None:
        start, size = np.asarray(start), np.asarray(size)
      elif start is not None and end is not None:
        start = np.asarray(start)
        size = np.asarray(end) - start
      elif size is not None and end is not None:
        size = np.asarray(size)
        start = np.asarray(end) - size
      else:
        raise ValueError('two of start / size / end are required')
      self.start = start.astype(int).copy()
      self.size = size.astype(int).copy()

    @property
    def end(self):
      return self.start + self.size

    @property
    def rank(self):
      return len(self.start)

    def scale(self, factors) -> 'BoundingBox':
      """Scales start and end (floor / ceil), as connectomics' BoundingBox.scale."""
      f = np.asarray(factors, dtype=float)
      start = np.floor(self.start * f).astype(int)
      end = np.ceil(self.end * f).astype(int)
      return BoundingBox(start=start, end=end)

    def adjusted_by(self, start=None, end=None) -> 'BoundingBox':
      s = self.start + (0 if start is None else np.asarray(start))
      e = self.end + (0 if end is None else np.asarray(end))
      return BoundingBox(start=s, end=e)

    def intersection(self, other: 'BoundingBox'):
      """Overlap of two boxes, or None if they do not intersect."""
      start = np.maximum(self.start, other.start)
      end = np.minimum(self.end, other.end)
      if np.any(end <= start):
        return None
      return BoundingBox(start=start, end=end)

    def translate(self, offset) -> 'BoundingBox':
      return BoundingBox(start=self.start + np.asarray(offset), size=self.size)

    def to_slice3d(self):
      return tuple(slice(int(s), int(e)) for s, e in
                   zip(self.start[::-1], self.end[::-1]))

    def to_slice4d(self):
      return (slice(None),) + self.to_slice3d()

    def __eq__(self, other):
      return (isinstance(other, BoundingBox) and np.array_equal(self.start, other.start)
              and np.array_equal(self.size, other.size))

    def __repr__(self):
      return f'BoundingBox(start={self.start.tolist()}, size={self.size.tolist()})'

  @dataclasses.dataclass
  class Subvolume:
    """[C, Z, Y, X] data with the XYZ box it covers."""
    data: np.ndarray
    bbox: BoundingBox

    @property
    def shape(self):
      return self.data.shape

  class SuggestedXyz(tuple):
    def __new__(cls, x, y, z):
      return super().__new__(cls, (x, y, z))

  class SubvolumeProcessor:
    """Geometry contract of connectomics' SubvolumeProcessor used by SOFIMA."""

    crop_at_borders = True
    output_num = 1
    ignores_input_data = False

    @property
    def namespace(self) -> str:
      return type(self).__name__

    def context(self):
      return (0, 0, 0), (0, 0, 0)

    def subvolume_size(self):
      return None

    def output_type(self, input_type):
      return input_type

    def num_channels(self, input_channels):
      return input_channels

    def pixelsize(self, psize):
      return psize

    def overlap(self):
      pre, post = self.context()
      return tuple(int(a) + int(b) for a, b in zip(pre, post))

    def crop_box(self, box: BoundingBox) -> BoundingBox:
      """Removes the context from a box."""
      pre, post = self.context()
      return box.adjusted_by(start=np.asarray(pre), end=-np.asarray(post))

    def crop_box_and_data(self, box: BoundingBox, data: np.ndarray) -> Subvolume:
      pre, post = self.context()
      cropped = self.crop_box(box)
      sel = [slice(None)]
      for ax in (2, 1, 0):  # data is [C, Z, Y, X]
        hi = data.shape[3 - ax] - int(post[ax])
        sel.append(slice(int(pre[ax]), hi))
      return Subvolume(data[tuple(sel)], cropped)

    def expected_output_box(self, box: BoundingBox) -> BoundingBox:
      return self.crop_box(box)

    def process(self, subvol: Subvolume):
      raise NotImplementedError


class _Counter:

  def inc(self, n: int = 1):
    del n


def counter(namespace: str, name: str) -> _Counter:
  del namespace, name
  return _Counter()


@contextlib.contextmanager
def timer_counter(namespace: str, name: str):
  del namespace, name
  yield
