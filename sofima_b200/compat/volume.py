"""A small in-memory stand-in for the TensorStore objects the SOFIMA decorators wrap
(reference decorators/flow.py; `tensorstore` and `connectomics.volume.decorators` are not
installable in this image).

Only the surface the decorators use is provided:

  ArrayStore(array, labels)          a labelled N-d array: .shape, .dtype, .domain (labels,
                                     per-dimension inclusive_min / exclusive_max / size),
                                     indexing with `...`, slices or an IndexDomain, `.read()
                                     .result()`, `np.array(store)`
  VirtualChunked(read_fn, ...)       what `ts.virtual_chunked` gives the reference: a lazy view
                                     whose chunks are produced by `read_fn(domain, array,
                                     read_params)` when they are read
  Decorator / Filter                 the two base classes of connectomics.volume.decorators
                                     with the constructor arguments the reference passes

Objects of the real `tensorstore` package duck-type into the same code paths for reading
(`.shape`, `.domain.labels`, indexing, `np.array`).
"""

from __future__ import annotations

import dataclasses
import itertools
from typing import Any, Callable, Mapping, Optional, Sequence

import numpy as np


@dataclasses.dataclass(frozen=True)
class Dim:
  inclusive_min: int
  exclusive_max: int
  label: str = ''

  @property
  def size(self) -> int:
    return self.exclusive_max - self.inclusive_min


class IndexDomain:
  """Ordered list of labelled dimensions (ts.IndexDomain)."""

  def __init__(self, dims: Sequence[Dim]):
    self._dims = list(dims)

  @property
  def labels(self):
    return tuple(d.label for d in self._dims)

  @property
  def shape(self):
    return tuple(d.size for d in self._dims)

  def __iter__(self):
    return iter(self._dims)

  def __len__(self):
    return len(self._dims)

  def __getitem__(self, key):
    if isinstance(key, str):
      return self._dims[self.labels.index(key)]
    return self._dims[key]


class _Future:

  def __init__(self, value):
    self._value = value

  def result(self):
    return self._value


class _StoreBase:
  """Reading interface shared by ArrayStore and VirtualChunked."""

  domain: IndexDomain
  dtype: np.dtype

  @property
  def shape(self):
    return self.domain.shape

  @property
  def rank(self):
    return len(self.domain)

  def _read(self, domain: IndexDomain) -> np.ndarray:
    raise NotImplementedError

  def _select(self, key) -> IndexDomain:
    """IndexDomain of `store[key]`; label-matched for IndexDomain keys, positional else."""
    if isinstance(key, IndexDomain):
      mine = {d.label: d for d in self.domain}
      if set(key.labels) != set(mine):
        raise ValueError(f'domain labels {key.labels} do not match {self.domain.labels}')
      return IndexDomain([key[l] for l in self.domain.labels])
    if key is Ellipsis:
      return self.domain
    if not isinstance(key, tuple):
      key = (key,)
    if Ellipsis in key:
      i = key.index(Ellipsis)
      key = key[:i] + (slice(None),) * (self.rank - len(key) + 1) + key[i + 1:]
    key = key + (slice(None),) * (self.rank - len(key))
    dims = []
    for d, k in zip(self.domain, key):
      if not isinstance(k, slice) or k.step not in (None, 1):
        raise NotImplementedError('only contiguous slices are supported')
      lo = d.inclusive_min if k.start is None else k.start
      hi = d.exclusive_max if k.stop is None else k.stop
      dims.append(Dim(max(lo, d.inclusive_min), min(hi, d.exclusive_max), d.label))
    return IndexDomain(dims)

  def __getitem__(self, key):
    return _View(self, self._select(key))

  def read(self):
    return _Future(self._read(self.domain))

  def __array__(self, dtype=None, copy=None):
    a = self._read(self.domain)
    return a.astype(dtype) if dtype is not None else a


class _View(_StoreBase):

  def __init__(self, base: _StoreBase, domain: IndexDomain):
    self._base, self.domain, self.dtype = base, domain, base.dtype

  def _read(self, domain):
    return self._base._read(domain)


class ArrayStore(_StoreBase):
  """Labelled in-memory array (what an opened ts.TensorStore is to the decorators)."""

  def __init__(self, array, labels: Optional[Sequence[str]] = None,
               inclusive_min: Optional[Sequence[int]] = None):
    self._a = np.asarray(array)
    labels = list(labels) if labels is not None else [''] * self._a.ndim
    lo = list(inclusive_min) if inclusive_min is not None else [0] * self._a.ndim
    self.domain = IndexDomain([Dim(l, l + n, lab) for l, n, lab in zip(lo, self._a.shape, labels)])
    self.dtype = self._a.dtype

  def _read(self, domain):
    sl = tuple(slice(q.inclusive_min - d.inclusive_min, q.exclusive_max - d.inclusive_min)
               for d, q in zip(self.domain, domain))
    return np.array(self._a[sl])

  def __setitem__(self, key, value):
    dom = self._select(key)
    sl = tuple(slice(q.inclusive_min - d.inclusive_min, q.exclusive_max - d.inclusive_min)
               for d, q in zip(self.domain, dom))
    self._a[sl] = value


class VirtualChunked(_StoreBase):
  """Lazy chunked view: `read_fn(domain, array, read_params)` fills one chunk."""

  def __init__(self, read_fn: Callable, domain: IndexDomain, chunk_shape: Sequence[int], dtype):
    self._fn, self.domain, self._chunk = read_fn, domain, [int(c) for c in chunk_shape]
    self.dtype = np.dtype(dtype)

  def _read(self, domain):
    out = np.empty(domain.shape, self.dtype)
    # every chunk that intersects the request is produced whole, then cropped
    ranges = []
    for d, q, c in zip(self.domain, domain, self._chunk):
      first = (q.inclusive_min - d.inclusive_min) // c
      last = (q.exclusive_max - 1 - d.inclusive_min) // c
      ranges.append(range(first, last + 1))
    for idx in itertools.product(*ranges):
      dims = [Dim(d.inclusive_min + i * c, min(d.inclusive_min + (i + 1) * c, d.exclusive_max),
                  d.label) for d, i, c in zip(self.domain, idx, self._chunk)]
      cdom = IndexDomain(dims)
      buf = np.empty(cdom.shape, self.dtype)
      self._fn(cdom, buf, None)
      src, dst = [], []
      for cd, q in zip(cdom, domain):
        lo, hi = max(cd.inclusive_min, q.inclusive_min), min(cd.exclusive_max, q.exclusive_max)
        src.append(slice(lo - cd.inclusive_min, hi - cd.inclusive_min))
        dst.append(slice(lo - q.inclusive_min, hi - q.inclusive_min))
      out[tuple(dst)] = buf[tuple(src)]
    return out


def open_store(spec) -> _StoreBase:
  """ts.open(spec).result() stand-in: accepts an already opened store (ArrayStore /
  VirtualChunked / TensorStore) or a mapping {'array': ndarray, 'labels': [...]}."""
  if isinstance(spec, Mapping):
    if 'array' in spec:
      return ArrayStore(spec['array'], spec.get('labels'), spec.get('inclusive_min'))
    try:
      import tensorstore as ts  # pragma: no cover - not installed here
    except ImportError as e:
      raise ImportError('a TensorStore JSON spec needs the tensorstore package; pass an opened '
                        "store or {'array': ..., 'labels': ...} instead") from e
    return ts.open(spec).result()  # pragma: no cover
  return spec


class Decorator:
  """connectomics.volume.decorators.Decorator: holds the virtual_chunked context spec."""

  def __init__(self, context_spec: Optional[Mapping[str, Any]] = None):
    self._context_spec = context_spec
    self._context = context_spec

  def decorate(self, input_ts):
    raise NotImplementedError


class Filter(Decorator):
  """connectomics.volume.decorators.Filter: applies `filter_fun(chunk, **filter_args)` to
  every chunk of at least `min_chunksize` elements per dimension (default: the whole
  array is one chunk), same shape out as in."""

  def __init__(self, filter_fun: Callable, context_spec=None,
               min_chunksize: Optional[Sequence[int]] = None, **filter_args):
    super().__init__(context_spec)
    self._filter_fun = filter_fun
    self._min_chunksize = min_chunksize
    self._filter_args = filter_args

  def _chunk_shape(self, input_ts):
    if self._min_chunksize is None:
      return list(input_ts.shape)
    return [min(int(c), int(n)) for c, n in zip(self._min_chunksize, input_ts.shape)]

  def decorate(self, input_ts):
    dom = _domain_of(input_ts)

    def filt_read(domain, array, unused_read_params):
      array[...] = self._filter_fun(np.array(input_ts[domain]), **self._filter_args)

    return VirtualChunked(filt_read, dom, self._chunk_shape(input_ts), input_ts.dtype)


def _domain_of(store) -> IndexDomain:
  """IndexDomain of an ArrayStore / VirtualChunked / duck-typed TensorStore."""
  if isinstance(store.domain, IndexDomain):
    return store.domain
  return IndexDomain([Dim(int(d.inclusive_min), int(d.exclusive_max), d.label)  # pragma: no cover
                      for d in store.domain])
