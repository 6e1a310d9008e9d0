"""The C-ABI library loads and exports every symbol include/b200nufft.h declares (no GPU)."""
import ctypes
import os
import re

from jax_finufft_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = set()
    for h in ("b200nufft.h", "b200nufft_dist.h"):
        p = os.path.join(ROOT, "include", h)
        if os.path.exists(p):
            src = re.sub(r"/\*.*?\*/", "", open(p).read(), flags=re.S)
            names |= set(re.findall(r"\b(b2n_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_header_symbols_exported():
    L = ctypes.CDLL(_lib.LIB_PATH)
    decl = _declared()
    assert len(decl) >= 18
    missing = [n for n in decl if not hasattr(L, n)]
    assert not missing, missing
    assert set(_lib.EXPORTED) <= set(decl)


def test_default_opts_match_reference_defaults():
    # V/src/cuda/cufinufft.cu:133-152: method 0 (auto), sort 1, kerevalmeth 1, upsampfac 0 (auto), modeord 0
    o = _lib.default_opts()
    assert (o.gpu_method, o.gpu_sort, o.gpu_kerevalmeth, o.upsampfac, o.modeord, o.gpu_maxbatchsize) == (0, 1, 1, 0.0, 0, 0)


def test_error_codes_before_any_device_work():
    # integer codes pinned by V/test/cuda/cufinufft_error_handling.cu:24-95 / test_makeplan.c
    L = _lib.lib()
    h = ctypes.c_void_p()
    o = _lib.default_opts()
    nm = (ctypes.c_int64 * 3)(10, 10, 10)
    mk = lambda typ, dim, ntr, eps=1e-6, modes=nm: L.b2n_makeplan(typ, dim, modes, 1, ntr, eps, 0, ctypes.byref(h), ctypes.byref(o))
    assert mk(1, 0, 1) == 12 and mk(1, 4, 1) == 12          # dim
    assert mk(0, 2, 1) == 10 and mk(4, 2, 1) == 10          # type
    assert mk(1, 2, 0) == 9                                 # ntransf
    bad = (ctypes.c_int64 * 3)(10, -1, 10)
    assert mk(1, 2, 1, modes=bad) == 14                     # negative modes
    big = (ctypes.c_int64 * 3)(1 << 31, 4, 4)
    assert mk(1, 2, 1, modes=big) == 14                     # oversize modes
    huge = (ctypes.c_int64 * 3)(1 << 20, 1 << 20, 1)
    assert mk(1, 2, 1, modes=huge) == 14                    # product overflows int32
    o.upsampfac = 0.9
    o.gpu_kerevalmeth = 0
    assert mk(1, 2, 1) == 7                                 # sigma <= 1, direct evaluation
    o.gpu_kerevalmeth = 1
    assert mk(1, 2, 1) == 8                                 # Horner needs sigma in {2, 1.25}
    assert L.b2n_destroy(None) == 16                        # destroy(NULL)
    assert L.b2n_setpts(None, 0, None, None, None, 0, None, None, None) == 16
