"""CPU oracle for the B200 NUFFT hot path -- TEST INFRASTRUCTURE ONLY.

ctypes front-end of ``oracle/nufft_oracle.c`` (a float64 restatement of the reference
cuFINUFFT algorithm, every function citing the reference file:line it follows) plus, on a GPU
box, ``oracle/ref_cufinufft.py`` (ctypes binding of the UNMODIFIED reference library built into
``oracle/_ref``).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package; ``jax_finufft_b200``
never does.
"""

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)


class OrcInfo(C.Structure):
    _fields_ = [("ns", C.c_int), ("beta", C.c_double), ("nf", C.c_long * 3)]


def build():
    """Compile the C oracle in-tree (gcc, seconds)."""
    subprocess.run(["make", "-f", os.path.join(_HERE, "Makefile")], check=True, cwd=os.path.dirname(_HERE))


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libnufft_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.orc_es_kernel.restype = C.c_double
        _LIB.orc_es_kernel.argtypes = [C.c_double, C.c_int, C.c_double]
        _LIB.orc_fold_rescale.restype = C.c_double
        _LIB.orc_fold_rescale.argtypes = [C.c_double, C.c_long, C.c_int]
        _LIB.orc_next235beven.restype = C.c_long
        _LIB.orc_next235beven.argtypes = [C.c_long, C.c_long]
        _LIB.orc_set_nf_type12.restype = C.c_long
        _LIB.orc_set_nf_type12.argtypes = [C.c_long, C.c_double, C.c_int]
        _LIB.orc_setup_spreader.restype = C.c_int
        _LIB.orc_setup_spreader.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int, c_dp, C.POINTER(C.c_int)]
    return _LIB


def _d(a):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _cz(a):
    return np.ascontiguousarray(a, dtype=np.complex128)


def setup_spreader(eps, upsampfac=2.0, kerevalmeth=1, is_float=False):
    """-> (ns, beta, ier).  Reference: V/src/cuda/spreadinterp.cpp:48-90."""
    beta = C.c_double()
    ier = C.c_int()
    ns = lib().orc_setup_spreader(eps, upsampfac, kerevalmeth, int(is_float), C.byref(beta), C.byref(ier))
    return ns, beta.value, ier.value


def next235beven(n, b=1):
    return lib().orc_next235beven(n, b)


def set_nf_type12(ms, upsampfac, ns):
    return lib().orc_set_nf_type12(ms, upsampfac, ns)


def es_kernel(x, ns, beta):
    return lib().orc_es_kernel(float(x), ns, beta)


def fold_rescale(x, nf, prec=0):
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    return np.array([lib().orc_fold_rescale(float(v), nf, prec) for v in x])


def gaussquad(n):
    x = np.zeros(n)
    w = np.zeros(n)
    lib().orc_gaussquad(C.c_int(n), _p(x), _p(w))
    return x, w


def fseries(nf, ns, beta):
    out = np.zeros(nf // 2 + 1)
    lib().orc_fseries(C.c_long(nf), C.c_int(ns), C.c_double(beta), _p(out))
    return out


def nuft(ns, beta, k):
    k = _d(k)
    out = np.zeros_like(k)
    lib().orc_nuft(C.c_int(ns), C.c_double(beta), C.c_long(k.size), _p(k), _p(out))
    return out


def _pts(pts, dim):
    pts = [_d(p) for p in pts]
    while len(pts) < 3:
        pts.append(None)
    return pts


def binsort(pts, nf, binsize, prec=0):
    """-> (binid[M], hist[nbins]) with x-fastest bin numbering (3d/spreadinterp3d.cuh:28-56)."""
    dim = len(pts)
    x, y, z = _pts(pts, dim)
    nf = list(nf) + [1] * (3 - dim)
    bs = list(binsize) + [1] * (3 - dim)
    M = x.size
    nb = [(nf[d] + bs[d] - 1) // bs[d] if d < dim else 1 for d in range(3)]
    binid = np.zeros(M, dtype=np.int32)
    hist = np.zeros(nb[0] * nb[1] * nb[2], dtype=np.int32)
    lib().orc_binsort(C.c_int(dim), C.c_long(M), _p(x), _p(y), _p(z), C.c_long(nf[0]), C.c_long(nf[1]),
                      C.c_long(nf[2]), C.c_int(bs[0]), C.c_int(bs[1]), C.c_int(bs[2]), C.c_int(prec),
                      _p(binid), _p(hist))
    return binid, hist


def binsort_anchor(pts, nf, binsize, ns, prec=0):
    """-> (binid[M], hist[nbins], uz[M]): anchor-cell bins of the 3-D sliding-window kernels."""
    x, y, z = _pts(pts, 3)
    M = x.size
    nb = [(nf[d] + binsize[d] - 1) // binsize[d] for d in range(3)]
    binid = np.zeros(M, dtype=np.int32)
    uz = np.zeros(M, dtype=np.int32)
    hist = np.zeros(nb[0] * nb[1] * nb[2], dtype=np.int32)
    lib().orc_binsort_anchor(C.c_long(M), _p(x), _p(y), _p(z), C.c_long(nf[0]), C.c_long(nf[1]), C.c_long(nf[2]),
                             C.c_int(binsize[0]), C.c_int(binsize[1]), C.c_int(binsize[2]), C.c_int(ns),
                             C.c_int(prec), _p(binid), _p(hist), _p(uz))
    return binid, hist, uz


def spread(pts, c, nf, ns, beta, prec=0):
    """Type-1 gridding of strengths c at pts onto a zeroed fine grid of shape nf[::-1]."""
    dim = len(pts)
    x, y, z = _pts(pts, dim)
    nf3 = list(nf) + [1] * (3 - dim)
    c = _cz(c)
    fw = np.zeros(tuple(int(n) for n in nf[::-1]), dtype=np.complex128)
    lib().orc_spread(C.c_int(dim), C.c_long(x.size), _p(x), _p(y), _p(z), _p(c), C.c_long(nf3[0]),
                     C.c_long(nf3[1]), C.c_long(nf3[2]), C.c_int(ns), C.c_double(beta), C.c_int(prec), _p(fw))
    return fw


def interp(pts, fw, ns, beta, prec=0):
    """Type-2 gather from fine grid fw (shape nf[::-1]) at pts."""
    dim = len(pts)
    x, y, z = _pts(pts, dim)
    fw = _cz(fw)
    nf = list(fw.shape[::-1]) + [1] * (3 - dim)
    c = np.zeros(x.size, dtype=np.complex128)
    lib().orc_interp(C.c_int(dim), C.c_long(x.size), _p(x), _p(y), _p(z), _p(c), C.c_long(nf[0]),
                     C.c_long(nf[1]), C.c_long(nf[2]), C.c_int(ns), C.c_double(beta), C.c_int(prec), _p(fw))
    return c


def _stack(a, inner_ndim):
    a = np.asarray(a)
    single = a.ndim == inner_ndim
    return (a[None] if single else a), single


def nufft1(n_modes, c, *pts, iflag=1, eps=1e-6, upsampfac=2.0, modeord=0, kerevalmeth=1, prec=0,
           return_info=False):
    """Type 1.  n_modes/pts in the *backend* (x-fastest) order: n_modes=(ms,mt,mu), pts=(x,y,z).
    Output shape (ntransf?, mu, mt, ms) i.e. C-order with ms fastest."""
    dim = len(pts)
    x, y, z = _pts(pts, dim)
    c, single = _stack(c, 1)
    c = _cz(c)
    nm = (C.c_long * 3)(*(list(n_modes) + [1] * (3 - dim)))
    out = np.zeros((c.shape[0],) + tuple(int(n) for n in n_modes[::-1]), dtype=np.complex128)
    info = OrcInfo()
    ier = lib().orc_nufft12(C.c_int(1), C.c_int(dim), C.c_long(x.size), _p(x), _p(y), _p(z), _p(c), C.c_int(iflag),
                            C.c_double(eps), nm, _p(out), C.c_int(c.shape[0]), C.c_double(upsampfac),
                            C.c_int(modeord), C.c_int(kerevalmeth), C.c_int(prec), C.byref(info))
    if ier > 1:
        raise RuntimeError(f"oracle nufft1 failed with code {ier}")
    out = out[0] if single else out
    return (out, info) if return_info else out


def nufft2(f, *pts, iflag=-1, eps=1e-6, upsampfac=2.0, modeord=0, kerevalmeth=1, prec=0, return_info=False):
    """Type 2.  f has shape (ntransf?, mu, mt, ms) (ms fastest); pts=(x,y,z)."""
    dim = len(pts)
    x, y, z = _pts(pts, dim)
    f, single = _stack(f, dim)
    f = _cz(f)
    n_modes = list(f.shape[1:][::-1])
    nm = (C.c_long * 3)(*(n_modes + [1] * (3 - dim)))
    out = np.zeros((f.shape[0], x.size), dtype=np.complex128)
    info = OrcInfo()
    ier = lib().orc_nufft12(C.c_int(2), C.c_int(dim), C.c_long(x.size), _p(x), _p(y), _p(z), _p(out), C.c_int(iflag),
                            C.c_double(eps), nm, _p(f), C.c_int(f.shape[0]), C.c_double(upsampfac),
                            C.c_int(modeord), C.c_int(kerevalmeth), C.c_int(prec), C.byref(info))
    if ier > 1:
        raise RuntimeError(f"oracle nufft2 failed with code {ier}")
    out = out[0] if single else out
    return (out, info) if return_info else out


def nufft3(c, pts, tgt, iflag=-1, eps=1e-6, upsampfac=2.0, kerevalmeth=1, prec=0, return_info=False):
    """Type 3.  pts=(x,y,z) sources, tgt=(s,t,u) targets."""
    dim = len(pts)
    x, y, z = _pts(pts, dim)
    s, t, u = _pts(tgt, dim)
    c, single = _stack(c, 1)
    c = _cz(c)
    out = np.zeros((c.shape[0], s.size), dtype=np.complex128)
    info = OrcInfo()
    ier = lib().orc_nufft3(C.c_int(dim), C.c_long(x.size), _p(x), _p(y), _p(z), _p(c), C.c_int(iflag),
                           C.c_double(eps), C.c_long(s.size), _p(s), _p(t), _p(u), _p(out), C.c_int(c.shape[0]),
                           C.c_double(upsampfac), C.c_int(kerevalmeth), C.c_int(prec), C.byref(info))
    if ier > 1:
        raise RuntimeError(f"oracle nufft3 failed with code {ier}")
    out = out[0] if single else out
    return (out, info) if return_info else out


def dirft1(n_modes, c, *pts, iflag=1, modeord=0):
    dim = len(pts)
    x, y, z = _pts(pts, dim)
    c = _cz(c)
    nm = list(n_modes) + [1] * (3 - dim)
    out = np.zeros(tuple(int(n) for n in n_modes[::-1]), dtype=np.complex128)
    lib().orc_dirft1(C.c_int(dim), C.c_long(x.size), _p(x), _p(y), _p(z), _p(c), C.c_int(iflag), C.c_long(nm[0]),
                     C.c_long(nm[1]), C.c_long(nm[2]), _p(out), C.c_int(modeord))
    return out


def dirft2(f, *pts, iflag=-1, modeord=0):
    dim = len(pts)
    x, y, z = _pts(pts, dim)
    f = _cz(f)
    nm = list(f.shape[::-1]) + [1] * (3 - dim)
    out = np.zeros(x.size, dtype=np.complex128)
    lib().orc_dirft2(C.c_int(dim), C.c_long(x.size), _p(x), _p(y), _p(z), _p(out), C.c_int(iflag), C.c_long(nm[0]),
                     C.c_long(nm[1]), C.c_long(nm[2]), _p(f), C.c_int(modeord))
    return out


def dirft3(c, pts, tgt, iflag=-1):
    dim = len(pts)
    x, y, z = _pts(pts, dim)
    s, t, u = _pts(tgt, dim)
    c = _cz(c)
    out = np.zeros(s.size, dtype=np.complex128)
    lib().orc_dirft3(C.c_int(dim), C.c_long(x.size), _p(x), _p(y), _p(z), _p(c), C.c_int(iflag), C.c_long(s.size),
                     _p(s), _p(t), _p(u), _p(out))
    return out


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    """OpenMP threads of the C restatement (launchers such as torchrun export OMP_NUM_THREADS=1)."""
    lib().orc_set_num_threads(C.c_int(int(n)))


def relerr(a, b):
    """relative l2 error ||a-b||/||b|| in float64 (V/test/utils/norms.hpp:15-37)."""
    a = np.asarray(a, dtype=np.complex128).ravel()
    b = np.asarray(b, dtype=np.complex128).ravel()
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))
