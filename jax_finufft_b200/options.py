"""Tuning options: the ``opts`` argument of ``nufft1/2/3``.

Host-side mirror of the reference's option structs (``src/jax_finufft/options.py``): the same
class names, field names, defaults and resolution rule, so that user code written against
jax-finufft (``Opts(gpu_method=..., modeord=...)``, ``NestedOpts(type1=..., backward=...)``) keeps
working.  The implementation is table-driven: one table lists every field with its default, its
validation rule and whether it crosses the custom-call boundary.  Only seven GPU fields do
(options.py:105-119, lowering.py:157-174); the CPU-only fields are accepted and ignored -- there is
no CPU path in this backend.
"""

import dataclasses
from enum import IntEnum
from types import SimpleNamespace

__all__ = ["Opts", "NestedOpts", "unpack_opts", "DebugLevel", "GpuDebugLevel", "GpuMethod",
           "SpreadSort", "SpreadThread", "FftwFlags"]

# ------------------------------------------------------------------------------------------ enums
# (options.py:9-43).  FftwFlags: upstream reads these integers from its CPU extension module
# (FFTW's public flag values); there is no FFTW here, the enum only keeps `Opts(fftw=...)` call
# sites working.  GpuMethod additionally names the reference library's method 3.
DebugLevel = IntEnum("DebugLevel", {"Silent": 0, "Verbose": 1, "Noisy": 2})
GpuDebugLevel = IntEnum("GpuDebugLevel", {"Silent": 0, "Verbose": 1})
SpreadSort = IntEnum("SpreadSort", {"NoSort": 0, "Sort": 1, "Heuristic": 2})
SpreadThread = IntEnum("SpreadThread", {"Auto": 0, "Sequential": 1, "Parallel": 2})
GpuMethod = IntEnum("GpuMethod", {"Auto": 0, "Driven": 1, "Shared": 2, "OutputDriven": 3})
FftwFlags = IntEnum("FftwFlags", {"Estimate": 1 << 6, "Measure": 0, "Patient": 1 << 5, "Exhaustive": 1 << 3,
                                  "WisdomOnly": 1 << 21})
for _e in (DebugLevel, GpuDebugLevel, SpreadSort, SpreadThread, GpuMethod, FftwFlags):
    _e.__module__ = __name__

# ----------------------------------------------------------------------------------------- fields
# name, default, rule, name of the attribute it becomes on the custom call (None: stays on the host)
_FLAG = "flag"          # bool (or 0 / 1)
_ANY = None             # not validated: ignored by this backend
_FIELDS = (
    # CPU FINUFFT fields of the reference struct (options.py:50-63)
    ("modeord", False, _FLAG, "modeord"),
    ("debug", DebugLevel.Silent, _ANY, None),
    ("spread_debug", DebugLevel.Silent, _ANY, None),
    ("showwarn", False, _ANY, None),
    ("nthreads", 0, _ANY, None),
    ("fftw", FftwFlags.Estimate, _ANY, None),
    ("spread_sort", SpreadSort.Heuristic, _ANY, None),
    ("spread_kerevalmeth", True, _ANY, None),
    ("spread_kerpad", True, _ANY, None),
    ("upsampfac", 0.0, _ANY, None),
    ("spread_thread", SpreadThread.Auto, _ANY, None),
    ("maxbatchsize", 0, _ANY, None),
    ("spread_nthr_atomic", -1, _ANY, None),
    ("spread_max_sp_size", 0, _ANY, None),
    # GPU fields (options.py:65-78); defaults of V/src/cuda/cufinufft.cu:133-152
    ("gpu_upsampfac", 2.0, _ANY, "upsampfac"),
    ("gpu_method", 0, (0, 1, 2, 3), "gpu_method"),
    ("gpu_sort", True, _FLAG, "gpu_sort"),
    ("gpu_binsizex", 0, _ANY, None),
    ("gpu_binsizey", 0, _ANY, None),
    ("gpu_binsizez", 0, _ANY, None),
    ("gpu_obinsizex", 0, _ANY, None),
    ("gpu_obinsizey", 0, _ANY, None),
    ("gpu_obinsizez", 0, _ANY, None),
    ("gpu_maxsubprobsize", 1024, _ANY, None),
    ("gpu_kerevalmeth", True, _FLAG, "gpu_kerevalmeth"),
    ("gpu_spreadinterponly", False, _FLAG, None),
    ("gpu_maxbatchsize", 0, "nonneg", "gpu_maxbatchsize"),
    ("gpu_debug", GpuDebugLevel.Silent, (0, 1), "debug"),
)


def _check(self):
    for name, _, rule, _ in _FIELDS:
        v = getattr(self, name)
        if rule == _FLAG:
            if not isinstance(v, (bool, int)) or int(v) not in (0, 1):
                raise ValueError(f"Opts.{name} must be a bool, got {v!r}")
        elif rule == "nonneg":
            if int(v) < 0:
                raise ValueError(f"Opts.{name} must be >= 0")
        elif isinstance(rule, tuple):
            if int(v) not in rule:
                raise ValueError(f"Opts.{name} must be one of {rule}, got {v!r}")


def _to_cufinufft_opts(self):
    """The attributes that cross the custom-call boundary, as plain numbers (what the reference's
    ``Opts.to_cufinufft_opts`` hands to its lowering)."""
    out = SimpleNamespace()
    for name, default, _, attr in _FIELDS:
        if attr is not None:
            v = getattr(self, name)
            setattr(out, attr, float(v) if isinstance(default, float) else int(v))
    return out


Opts = dataclasses.make_dataclass(
    "Opts", [(n, type(d) if not isinstance(d, IntEnum) else int, dataclasses.field(default=d)) for n, d, _, _ in _FIELDS],
    frozen=True, namespace={"__post_init__": _check, "to_cufinufft_opts": _to_cufinufft_opts})
Opts.__module__ = __name__
Opts.__doc__ = "Options of one transform (same fields and defaults as the reference's ``Opts``)."

# Per-type and per-direction options (options.py:122-129): `type1/2/3` select by transform type,
# `forward` overrides them, `backward` (an Opts or another NestedOpts) applies to the transforms
# the differentiation rules create.
NestedOpts = dataclasses.make_dataclass(
    "NestedOpts", [(n, object, dataclasses.field(default=None)) for n in ("type1", "type2", "type3", "forward", "backward")],
    frozen=True)
NestedOpts.__module__ = __name__


def unpack_opts(opts, finufft_type, forward):
    """Resolve ``opts`` for one transform (options.py:132-148): a plain ``Opts`` (or None) applies
    to everything; a ``NestedOpts`` is looked up by direction, then by type; if the lookup has no
    answer for a backward transform the NestedOpts itself is returned (the caller resolves it
    again for the transposed type)."""
    if not isinstance(opts, NestedOpts):
        return opts
    if not forward:
        return opts if opts.backward is None else opts.backward
    if opts.forward is not None:
        return opts.forward
    return {1: opts.type1, 2: opts.type2, 3: opts.type3}.get(finufft_type, opts)
