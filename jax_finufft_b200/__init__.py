"""jax_finufft_b200 -- B200-native (sm_100a) NUFFT backend behind jax-finufft's Python API.

``nufft1 / nufft2 / nufft3`` mirror ``jax_finufft.nufft1/2/3`` (same names, argument meaning,
``Opts`` / ``NestedOpts`` tuning structs, vmap stacking and JVP/VJP rules) on CUDA torch tensors;
every transform runs in ``libb200nufft.so`` (hand-written CUDA for sm_100a + cuFFT).  The C ABI
is declared in ``include/b200nufft.h``; ``INTEGRATION.md`` shows how the reference binds it.
"""

from .options import NestedOpts, Opts, unpack_opts  # noqa: F401
from .ops import nufft1, nufft2, nufft3  # noqa: F401

__all__ = ["nufft1", "nufft2", "nufft3", "Opts", "NestedOpts"]
__version__ = "0.1.0"
