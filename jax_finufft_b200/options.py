"""Tuning options -- host-side mirror of ``src/jax_finufft/options.py`` (reference file:line cited
per item).  Same names, defaults and resolution rules; only the seven GPU fields that the
reference forwards across its FFI boundary (options.py:105-119, lowering.py:157-174) reach the
backend, exactly as upstream.  CPU-only fields are accepted and ignored (there is no CPU path).
"""

from dataclasses import dataclass
from enum import IntEnum
from typing import Optional, Union

__all__ = ["Opts", "NestedOpts", "unpack_opts", "DebugLevel", "GpuDebugLevel", "GpuMethod",
           "SpreadSort", "SpreadThread", "FftwFlags"]


class DebugLevel(IntEnum):  # options.py:9-12
    Silent = 0
    Verbose = 1
    Noisy = 2


class GpuDebugLevel(IntEnum):  # options.py:15-17
    Silent = 0
    Verbose = 1


class FftwFlags(IntEnum):
    # options.py:20-25 reads these ints from the CPU extension (FFTW's public flag values);
    # there is no FFTW here, the enum only keeps `Opts(fftw=...)` call sites working.
    Estimate = 1 << 6
    Measure = 0
    Patient = 1 << 5
    Exhaustive = 1 << 3
    WisdomOnly = 1 << 21


class SpreadSort(IntEnum):  # options.py:28-31
    NoSort = 0
    Sort = 1
    Heuristic = 2


class SpreadThread(IntEnum):  # options.py:34-37
    Auto = 0
    Sequential = 1
    Parallel = 2


class GpuMethod(IntEnum):  # options.py:40-43 (+ the reference library's method 3)
    Auto = 0
    Driven = 1
    Shared = 2
    OutputDriven = 3


@dataclass(frozen=True)
class Opts:  # options.py:46-79
    modeord: bool = False
    debug: int = DebugLevel.Silent
    spread_debug: int = DebugLevel.Silent
    showwarn: bool = False
    nthreads: int = 0
    fftw: int = FftwFlags.Estimate
    spread_sort: int = SpreadSort.Heuristic
    spread_kerevalmeth: bool = True
    spread_kerpad: bool = True
    upsampfac: float = 0.0
    spread_thread: int = SpreadThread.Auto
    maxbatchsize: int = 0
    spread_nthr_atomic: int = -1
    spread_max_sp_size: int = 0

    gpu_upsampfac: float = 2.0
    gpu_method: int = 0
    gpu_sort: bool = True
    gpu_binsizex: int = 0
    gpu_binsizey: int = 0
    gpu_binsizez: int = 0
    gpu_obinsizex: int = 0
    gpu_obinsizey: int = 0
    gpu_obinsizez: int = 0
    gpu_maxsubprobsize: int = 1024
    gpu_kerevalmeth: bool = True
    gpu_spreadinterponly: bool = False
    gpu_maxbatchsize: int = 0
    gpu_debug: int = GpuDebugLevel.Silent

    def __post_init__(self):
        for name in ("modeord", "gpu_sort", "gpu_kerevalmeth", "gpu_spreadinterponly"):
            v = getattr(self, name)
            if not isinstance(v, (bool, int)) or int(v) not in (0, 1):
                raise ValueError(f"Opts.{name} must be a bool, got {v!r}")
        if int(self.gpu_method) not in (0, 1, 2, 3):
            raise ValueError(f"Opts.gpu_method must be 0, 1, 2 or 3, got {self.gpu_method!r}")
        if int(self.gpu_debug) not in (0, 1):
            raise ValueError(f"Opts.gpu_debug must be 0 or 1, got {self.gpu_debug!r}")
        if int(self.gpu_maxbatchsize) < 0:
            raise ValueError("Opts.gpu_maxbatchsize must be >= 0")

    def to_cufinufft_opts(self):
        """The seven attributes that cross the FFI boundary (options.py:105-119)."""

        class NativeOpts:
            pass

        opts = NativeOpts()
        opts.modeord = int(self.modeord)
        opts.upsampfac = float(self.gpu_upsampfac)
        opts.gpu_method = int(self.gpu_method)
        opts.gpu_sort = int(self.gpu_sort)
        opts.gpu_kerevalmeth = int(self.gpu_kerevalmeth)
        opts.gpu_maxbatchsize = int(self.gpu_maxbatchsize)
        opts.debug = int(self.gpu_debug)
        return opts


@dataclass(frozen=True)
class NestedOpts:  # options.py:122-129
    type1: Optional[Opts] = None
    type2: Optional[Opts] = None
    type3: Optional[Opts] = None

    forward: Optional[Opts] = None
    backward: Optional[Union[Opts, "NestedOpts"]] = None


def unpack_opts(opts, finufft_type, forward):  # options.py:132-148
    if opts is None or isinstance(opts, Opts):
        return opts

    if forward:
        if opts.forward is not None:
            return opts.forward
        elif finufft_type == 1:
            return opts.type1
        elif finufft_type == 2:
            return opts.type2
        elif finufft_type == 3:
            return opts.type3
    elif opts.backward is not None:
        return opts.backward

    return opts
