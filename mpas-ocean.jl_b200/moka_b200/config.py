"""YAML configuration, mirror of src/infra/Config.jl (host-side control plane the hot path is driven from).

  yaml_config, GlobalConfig            Config.jl:10-33
  ConfigGet / ConfigAdd / ConfigSet    Config.jl:42-86
  ConfigRead                           Config.jl:96-118  (top-level key `omega`, `streams` split off)
  parse_Datetimes, timestamp pattern   Config.jl:120-224

Julia's `Dates` values map to: DateTime -> datetime.datetime, Time -> datetime.time, and the single-unit
periods Year/Month/Day/Hour/Minute/Second -> `Period(value, unit)` (time_manager.py), which adds to a
datetime with Julia's calendar rules.
"""
from __future__ import annotations

import datetime as _dt
import os
import re

import yaml

from ._lib import MokaError
from .time_manager import Period

# Config.jl:139-149 (extended-mode regex): [[[Y-]M-]D][_]hh:mm:ss
_TIMESTAMP = re.compile(r"^(?:(?:(\d{1,4})-)?(?:(\d\d?)-)?(\d+))?_?(\d\d):(\d\d):(\d\d)$")
_UNITS = ("year", "month", "day", "hour", "minute", "second")


class yaml_config:
    """Config.jl:10-17: a wrapper around one level of the YAML tree."""

    def __init__(self, d: dict | None = None):
        self.dict = {} if d is None else d

    def __repr__(self):
        return f"yaml_config({list(self.dict)})"


class GlobalConfig:
    """Config.jl:22-33: the namelist and streams trees."""

    def __init__(self, namelist: yaml_config | None = None, streams: yaml_config | None = None):
        self.namelist = namelist or yaml_config()
        self.streams = streams or yaml_config()


def ConfigGet(d: yaml_config, s: str):
    """Config.jl:42-56: a nested dict comes back wrapped, a leaf as its value; a missing key raises KeyError."""
    c = d.dict[s]
    return type(d)(c) if isinstance(c, dict) else c


def ConfigAdd(d: yaml_config, s: str, val) -> None:
    """Config.jl:60-67."""
    if s in d.dict:
        raise MokaError(f"ConfigAdd: variable {s} already exists use ConfigSet instead")
    d.dict[s] = val


def ConfigSet(d: yaml_config, s: str, val) -> None:
    """Config.jl:71-86 (a type change only warns in the reference; it is allowed here too)."""
    if s not in d.dict:
        raise MokaError(f"ConfigSet: Could not find variable {s}")
    d.dict[s] = val


def DateTime_from_String(string: str):
    """Config.jl:165-224: 'Y-M-D_h:m:s' with non-zero month and day -> datetime; exactly one non-zero field -> that
    Period; no date part (or a zero day) -> time of day; otherwise the string itself."""
    mat = _TIMESTAMP.match(string)
    if mat is None:
        raise MokaError("could not make sense of timestamp format")
    cap = mat.groups()
    if all(c is not None for c in cap):
        yr, mn, dy, h, m, s = (int(c) for c in cap)
        if mn != 0 and dy != 0:
            return _dt.datetime(yr, mn, dy, h, m, s)
    vals = [0 if c is None else int(c) for c in cap]
    if sum(v != 0 for v in vals) == 1:
        i = next(k for k, v in enumerate(vals) if v != 0)
        return Period(vals[i], _UNITS[i])
    h, m, s = vals[3:]
    if all(c is None for c in cap[:3]):
        return _dt.time(h, m, s)
    if cap[0] is None and cap[1] is None and vals[2] == 0:
        return _dt.time(h, m, s)
    return string


def parse_Datetimes(d: dict) -> dict:
    """Config.jl:120-137: in place, recursive; only strings matching the timestamp pattern are converted."""
    for key, value in d.items():
        if isinstance(value, dict):
            parse_Datetimes(value)
        elif isinstance(value, str) and _TIMESTAMP.match(value):
            d[key] = DateTime_from_String(value)
    return d


class _Loader(yaml.SafeLoader):
    """YAML 1.1 would turn 0000-00-00_00:05:00-like scalars into sexagesimal ints and 1.e25 into a string;
    YAML.jl yields strings / floats there.  Keep timestamps as strings, read 1.e25 as a float."""


_Loader.yaml_implicit_resolvers = {k: [(t, r) for t, r in v if t not in ("tag:yaml.org,2002:int", "tag:yaml.org,2002:float",
                                                                         "tag:yaml.org,2002:timestamp")]
                                   for k, v in yaml.SafeLoader.yaml_implicit_resolvers.items()}
_Loader.add_implicit_resolver("tag:yaml.org,2002:int", re.compile(r"^[-+]?(?:0|[1-9][0-9_]*)$"), list("-+0123456789"))
_Loader.add_implicit_resolver(
    "tag:yaml.org,2002:float",
    re.compile(r"^[-+]?(?:[0-9][0-9_]*\.[0-9_]*(?:[eE][-+]?[0-9]+)?|\.[0-9_]+(?:[eE][-+]?[0-9]+)?|[0-9][0-9_]*[eE][-+]?[0-9]+"
               r"|\.(?:inf|Inf|INF)|\.(?:nan|NaN|NAN))$"), list("-+0123456789."))


def ConfigRead(filepath: str) -> GlobalConfig:
    """Config.jl:96-118."""
    if not os.path.isfile(filepath):
        raise MokaError("YAML configuration file does not exist")
    with open(filepath) as f:
        config = yaml.load(f, Loader=_Loader)
    streams = config["omega"].pop("streams")
    namelist = config.pop("omega")
    return GlobalConfig(yaml_config(parse_Datetimes(namelist)), yaml_config(parse_Datetimes(streams)))
