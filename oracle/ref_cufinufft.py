"""ctypes binding of the UNMODIFIED reference cuFINUFFT built into ``oracle/_ref`` by
``oracle/Makefile.ref`` (sources stay under /root/reference; only the .so travels).

TEST INFRASTRUCTURE ONLY (checker + "reference GPU kernels on the same B200" timing).  C API:
vendor/finufft/include/cufinufft.h:19-39; options struct: vendor/finufft/include/cufinufft_opts.h.
"""

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(_HERE, "_ref", "libcufinufft_ref.so")
_lib = None


class CufinufftOpts(C.Structure):
    _fields_ = [
        ("upsampfac", C.c_double),
        ("gpu_method", C.c_int),
        ("gpu_sort", C.c_int),
        ("gpu_binsizex", C.c_int),
        ("gpu_binsizey", C.c_int),
        ("gpu_binsizez", C.c_int),
        ("gpu_obinsizex", C.c_int),
        ("gpu_obinsizey", C.c_int),
        ("gpu_obinsizez", C.c_int),
        ("gpu_maxsubprobsize", C.c_int),
        ("gpu_kerevalmeth", C.c_int),
        ("gpu_spreadinterponly", C.c_int),
        ("gpu_maxbatchsize", C.c_int),
        ("gpu_device_id", C.c_int),
        ("gpu_stream", C.c_void_p),
        ("modeord", C.c_int),
        ("gpu_np", C.c_int),
        ("debug", C.c_int),
    ]


def available():
    return os.path.exists(PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(PATH)
        vp, i64 = C.c_void_p, C.c_int64
        L.cufinufft_default_opts.argtypes = [C.POINTER(CufinufftOpts)]
        L.cufinufft_default_opts.restype = None
        L.cufinufftf_makeplan.argtypes = [C.c_int, C.c_int, C.POINTER(i64), C.c_int, C.c_int, C.c_float,
                                          C.POINTER(vp), C.POINTER(CufinufftOpts)]
        L.cufinufft_makeplan.argtypes = [C.c_int, C.c_int, C.POINTER(i64), C.c_int, C.c_int, C.c_double,
                                         C.POINTER(vp), C.POINTER(CufinufftOpts)]
        for n in ("cufinufftf_setpts", "cufinufft_setpts"):
            getattr(L, n).argtypes = [vp, i64, vp, vp, vp, C.c_int, vp, vp, vp]
        for n in ("cufinufftf_execute", "cufinufft_execute"):
            getattr(L, n).argtypes = [vp, vp, vp]
        for n in ("cufinufftf_destroy", "cufinufft_destroy"):
            getattr(L, n).argtypes = [vp]
        _lib = L
    return _lib


class RefPlan:
    """Same surface as jax_finufft_b200.plan.Plan, over the reference library.

    Mirrors how jax-finufft drives it (lib/kernels.cc.cu:25-92): upsampfac defaults to 2.0
    (src/jax_finufft/options.py:66), 3-D double forces gpu_method=1 (lib/cufinufft_wrapper.cc:26-36).
    """

    def __init__(self, nufft_type, n_modes_or_dim, n_trans=1, eps=1e-6, isign=None, dtype="complex64",
                 jax_defaults=True, **opts):
        L = lib()
        self._L = L
        self.type = int(nufft_type)
        if self.type == 3:
            self.dim = int(n_modes_or_dim)
            n_modes = (1, 1, 1)
        else:
            n_modes = tuple(int(n) for n in n_modes_or_dim)
            self.dim = len(n_modes)
        self.n_modes = n_modes
        self.cdtype = {"complex64": torch.complex64, "complex128": torch.complex128}[str(dtype).replace("torch.", "")]
        self.rdtype = torch.float32 if self.cdtype == torch.complex64 else torch.float64
        self.double = self.cdtype == torch.complex128
        self.n_trans = int(n_trans)
        if isign is None:
            isign = 1 if self.type == 1 else -1
        o = CufinufftOpts()
        L.cufinufft_default_opts(C.byref(o))
        if jax_defaults:
            o.upsampfac = 2.0
            if self.double and self.dim > 2:
                o.gpu_method = 1
        for k, v in opts.items():
            if not hasattr(o, k):
                raise TypeError(f"unknown option {k}")
            setattr(o, k, v)
        o.gpu_stream = torch.cuda.current_stream().cuda_stream
        o.gpu_device_id = torch.cuda.current_device()
        nm = (C.c_int64 * 3)(*(list(n_modes) + [1] * (3 - len(n_modes))))
        h = C.c_void_p()
        if self.double:
            ier = L.cufinufft_makeplan(self.type, self.dim, nm, int(isign), self.n_trans, float(eps), C.byref(h), C.byref(o))
        else:
            ier = L.cufinufftf_makeplan(self.type, self.dim, nm, int(isign), self.n_trans, float(eps), C.byref(h), C.byref(o))
        if ier > 1:
            raise RuntimeError(f"reference makeplan failed with code {ier}")
        self._h = h
        self.M = self.N = 0

    def setpts(self, x, y=None, z=None, s=None, t=None, u=None):
        pts = [p.to(self.rdtype).contiguous() for p in (x, y, z) if p is not None]
        tg = [p.to(self.rdtype).contiguous() for p in (s, t, u) if p is not None]
        self._keep = (pts, tg)
        self.M = pts[0].numel()
        self.N = tg[0].numel() if tg else 0
        pp = [C.c_void_p(p.data_ptr()) for p in pts] + [None] * (3 - len(pts))
        tp = [C.c_void_p(p.data_ptr()) for p in tg] + [None] * (3 - len(tg))
        f = self._L.cufinufft_setpts if self.double else self._L.cufinufftf_setpts
        ier = f(self._h, self.M, *pp, self.N, *tp)
        if ier != 0:
            raise RuntimeError(f"reference setpts failed with code {ier}")
        return self

    def execute(self, data, out=None):
        data = data.to(self.cdtype).contiguous()
        dev = data.device
        if self.type == 1:
            out = torch.empty((self.n_trans,) + tuple(self.n_modes[::-1]), dtype=self.cdtype, device=dev) if out is None else out
            c, fk = data, out
        elif self.type == 2:
            out = torch.empty((self.n_trans, self.M), dtype=self.cdtype, device=dev) if out is None else out
            c, fk = out, data
        else:
            out = torch.empty((self.n_trans, self.N), dtype=self.cdtype, device=dev) if out is None else out
            c, fk = data, out
        f = self._L.cufinufft_execute if self.double else self._L.cufinufftf_execute
        ier = f(self._h, C.c_void_p(c.data_ptr()), C.c_void_p(fk.data_ptr()))
        if ier != 0:
            raise RuntimeError(f"reference execute failed with code {ier}")
        return out

    def destroy(self):
        if self._h is not None and self._h.value:
            torch.cuda.synchronize()
            (self._L.cufinufft_destroy if self.double else self._L.cufinufftf_destroy)(self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
