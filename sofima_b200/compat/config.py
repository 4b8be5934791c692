"""Stand-ins for the configuration helpers of the un-vendored `connectomics` package that the
reference's default / pipeline configs use (processor/defaults/em_2d.py, pipeline/*.py):

  connectomics.common.utils.update_dataclass          -> update_dataclass
  connectomics.volume.subvolume_processor.ProcessingConfig, DefaultConfigType,
      register_default_config, default_config        -> same names
  dataclasses_json.DataClassJsonMixin                 -> JsonMixin (to_dict / from_dict /
                                                         to_json / from_json, nested)
"""

from __future__ import annotations

import dataclasses
import enum
import json
import typing
from typing import Any, Callable, Sequence


def update_dataclass(config, overrides: dict[str, Any] | None):
  """Returns a copy of `config` with `overrides` applied; a dict given for a field that
  holds a dataclass updates that dataclass recursively."""
  if not overrides:
    return config
  changes = {}
  names = {f.name for f in dataclasses.fields(config)}
  for key, value in overrides.items():
    if key not in names:
      raise KeyError(f'{type(config).__name__} has no field {key!r}')
    current = getattr(config, key)
    if dataclasses.is_dataclass(current) and isinstance(value, dict):
      changes[key] = update_dataclass(current, value)
    else:
      changes[key] = value
  return dataclasses.replace(config, **changes)


def _plain(v):
  if dataclasses.is_dataclass(v) and not isinstance(v, type):
    return {f.name: _plain(getattr(v, f.name)) for f in dataclasses.fields(v)}
  if isinstance(v, enum.Enum):
    return v.value
  if isinstance(v, dict):
    return {k: _plain(x) for k, x in v.items()}
  if isinstance(v, (list, tuple, set, frozenset)):
    return [_plain(x) for x in v]
  return v


class JsonMixin:
  """The four methods of dataclasses_json.DataClassJsonMixin the reference uses."""

  def to_dict(self) -> dict[str, Any]:
    return _plain(self)

  def to_json(self, **kw) -> str:
    return json.dumps(self.to_dict(), **kw)

  @classmethod
  def from_dict(cls, kvs: dict[str, Any]):
    try:
      hints = typing.get_type_hints(cls)
    except Exception:  # pylint: disable=broad-except
      hints = {}
    kwargs = {}
    for f in dataclasses.fields(cls):
      if f.name not in kvs:
        continue
      v = kvs[f.name]
      t = hints.get(f.name)
      if isinstance(v, dict) and isinstance(t, type) and dataclasses.is_dataclass(t):
        v = t.from_dict(v) if hasattr(t, 'from_dict') else t(**v)
      kwargs[f.name] = v
    return cls(**kwargs)

  @classmethod
  def from_json(cls, s: str):
    return cls.from_dict(json.loads(s))


@dataclasses.dataclass(eq=True)
class ProcessingConfig(JsonMixin):
  """How a volume is cut into subvolumes for a processor (XYZ sizes)."""
  overlap: Sequence[int] | None = None
  subvolume_size: Sequence[int] | None = None
  batch_size: int | None = None


class DefaultConfigType(enum.Enum):
  EM_2D = 'em_2d'
  EM = 'em'
  LM = 'lm'


_DEFAULTS: dict[tuple[DefaultConfigType, type], Callable] = {}


def register_default_config(config_type: DefaultConfigType, config_class: type,
                            fn: Callable[[dict[str, Any] | None], Any]):
  _DEFAULTS[(config_type, config_class)] = fn


def default_config(config_class: type, config_type: DefaultConfigType,
                   overrides: dict[str, Any] | None = None):
  return _DEFAULTS[(config_type, config_class)](overrides)
