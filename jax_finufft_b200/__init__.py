"""jax_finufft_b200 -- B200-native (sm_100a) NUFFT backend behind jax-finufft's Python API.

``nufft1 / nufft2 / nufft3`` mirror ``jax_finufft.nufft1/2/3`` (same names, argument meaning,
``Opts`` / ``NestedOpts`` tuning structs, vmap stacking and JVP/VJP rules) on CUDA torch tensors;
every transform runs in ``libb200nufft.so`` (hand-written CUDA for sm_100a + cuFFT).  The C ABI
is declared in ``include/b200nufft.h``; ``INTEGRATION.md`` shows how the reference binds it.
"""

from .options import NestedOpts, Opts, unpack_opts  # noqa: F401
from .ops import nufft1, nufft2, nufft3  # noqa: F401



def clear_cache():
    """Drop the library's cached plans and hand its device memory back to the driver.  The plan
    cache lives in a private CUDA memory pool that ``torch.cuda.empty_cache()`` cannot see."""
    from . import _lib
    _lib.lib().b2n_cache_clear()


def set_cache_limit(nbytes):
    """Bound the device memory parked plans may hold (default: a quarter of the device; the entry
    count is capped at 8 as well).  ``None`` restores the default.  Returns the previous limit."""
    from . import _lib
    return int(_lib.lib().b2n_set_cache_limit(-1 if nbytes is None else int(nbytes)))


def cache_bytes():
    """(reserved, used) bytes of the library's device memory pool on the current device."""
    import ctypes as C
    from . import _lib
    r, u = C.c_ulonglong(), C.c_ulonglong()
    _lib.lib().b2n_cache_bytes(C.byref(r), C.byref(u))
    return int(r.value), int(u.value)


__all__ = ["nufft1", "nufft2", "nufft3", "Opts", "NestedOpts", "clear_cache", "set_cache_limit", "cache_bytes"]
__version__ = "0.1.0"
