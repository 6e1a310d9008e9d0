"""Plan-level interface (makeplan / setpts / execute / destroy) over the C ABI.

Mirrors the guru interface the reference's host driver uses internally
(lib/cufinufft_wrapper.h: makeplan/setpts/execute/destroy; V/include/cufinufft.h:19-39).  The
public ``nufft1/2/3`` go through ``b2n_run`` instead; this class exists for stage-level tests,
spread/interp-only use (``gpu_spreadinterponly``) and benchmarking setpts/execute separately as
the reference's ``cuperftest`` does (V/perftest/cuda/cuperftest.cu:183-303).
"""

import ctypes as C

import torch

from . import _lib

__all__ = ["Plan"]


class Plan:
    def __init__(self, nufft_type, n_modes_or_dim, n_trans=1, eps=1e-6, isign=None, dtype="complex64", **opts):
        L = _lib.lib()
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
        self.n_trans = int(n_trans)
        if isign is None:
            isign = 1 if self.type == 1 else -1
        o = _lib.default_opts()
        for k, v in opts.items():
            if not hasattr(o, k):
                raise TypeError(f"unknown option {k}")
            setattr(o, k, v)
        self._opts = o
        nm = (C.c_int64 * 3)(*(list(n_modes) + [1] * (3 - len(n_modes))))
        h = C.c_void_p()
        o.gpu_stream = torch.cuda.current_stream().cuda_stream
        ier = L.b2n_makeplan(self.type, self.dim, nm, int(isign), self.n_trans, float(eps),
                             int(self.cdtype == torch.complex128), C.byref(h), C.byref(o))
        if ier > 1:
            raise RuntimeError(f"b2n_makeplan failed with code {ier}")
        self.warning = ier
        self._h = h
        self._keep = None
        self.M = 0
        self.N = 0

    def info(self):
        inf = _lib.B2nPlanInfo()
        self._L.b2n_plan_info_get(self._h, C.byref(inf))
        return inf

    def setpts(self, x, y=None, z=None, s=None, t=None, u=None):
        pts = [p for p in (x, y, z) if p is not None]
        tg = [p for p in (s, t, u) if p is not None]
        assert len(pts) == self.dim
        pts = [p.to(self.rdtype).contiguous() for p in pts]
        tg = [p.to(self.rdtype).contiguous() for p in tg]
        self._keep = (pts, tg)
        self.M = pts[0].numel()
        self.N = tg[0].numel() if tg else 0
        pp = [C.c_void_p(p.data_ptr()) for p in pts] + [None] * (3 - len(pts))
        tp = [C.c_void_p(p.data_ptr()) for p in tg] + [None] * (3 - len(tg))
        ier = self._L.b2n_setpts(self._h, self.M, *pp, self.N, *tp)
        if ier != 0:
            raise RuntimeError(f"b2n_setpts failed with code {ier}")
        return self

    def execute(self, data, out=None):
        """type 1/3: data = strengths [n_trans, M] -> modes / targets; type 2: data = modes -> [n_trans, M]."""
        data = data.to(self.cdtype).contiguous()
        dev = data.device
        if self.type == 1:
            shape = (self.n_trans,) + tuple(self.n_modes[::-1])
            out = torch.empty(shape, dtype=self.cdtype, device=dev) if out is None else out
            c, fk = data, out
        elif self.type == 2:
            out = torch.empty((self.n_trans, self.M), dtype=self.cdtype, device=dev) if out is None else out
            c, fk = out, data
        else:
            out = torch.empty((self.n_trans, self.N), dtype=self.cdtype, device=dev) if out is None else out
            c, fk = data, out
        ier = self._L.b2n_execute(self._h, C.c_void_p(c.data_ptr()), C.c_void_p(fk.data_ptr()))
        if ier != 0:
            raise RuntimeError(f"b2n_execute failed with code {ier}")
        return out

    def sort_arrays(self):
        """(idx[M], bin_start[nbins+1]) as torch int32 tensors (copies)."""
        inf = self.info()
        nb = int(inf.nbins[0]) * int(inf.nbins[1]) * int(inf.nbins[2])
        out_idx = torch.empty(self.M, dtype=torch.int32, device="cuda")
        out_bs = torch.empty(nb + 1, dtype=torch.int32, device="cuda")
        ier = self._L.b2n_plan_sort_copy(self._h, C.c_void_p(out_idx.data_ptr()), C.c_void_p(out_bs.data_ptr()))
        if ier:
            raise RuntimeError(f"b2n_plan_sort_copy failed with code {ier}")
        return out_idx, out_bs

    def timings(self):
        t = (C.c_double * 7)()
        self._L.b2n_plan_timings(self._h, t)
        return dict(zip(["sort", "spread", "fft", "deconv_amplify", "interp", "type3_prepost", "memset"], list(t)))

    def destroy(self):
        if self._h is not None and self._h.value:
            torch.cuda.synchronize()
            self._L.b2n_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
