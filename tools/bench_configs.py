#!/usr/bin/env python
"""Times the BASELINE.json configs that are NOT the bench.py headline (C2, C4, C5 + the C5 gradient
pass) on one GPU, through the public API, with the library's per-stage device timers.  One JSON
line per config on stdout.  Usage (GPU box): python tools/bench_configs.py [c2 c4 c5 c5grad]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jax_finufft_b200 as J  # noqa: E402
from jax_finufft_b200.plan import Plan  # noqa: E402

dev = torch.device("cuda:0")


def rnd(shape, g, lo=-1.0, hi=1.0):
    return torch.rand(shape, device=dev, generator=g) * (hi - lo) + lo


def cplx(shape, g):
    return torch.complex(rnd(shape, g), rnd(shape, g))


def timed(step, K=5, W=3):
    for _ in range(W):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K


def stages(p, setpts, execute, K=3):
    for it in range(K + 1):
        setpts()
        execute()
        if it == 0:
            p.timings()
    return {k: round(v / K, 4) for k, v in p.timings().items()}


def c1():
    """1-D type 1, M=1e6 random points, N=1e6 modes, eps=1e-6, complex128 (BASELINE configs[0], the
    reference's CPU-runnable case; here on the GPU in double precision)."""
    g = torch.Generator(device=dev).manual_seed(123)
    M, N = 10 ** 6, 10 ** 6
    x = (torch.rand(M, device=dev, generator=g, dtype=torch.float64) * 2 * np.pi)
    c = torch.complex(torch.randn(M, device=dev, generator=g, dtype=torch.float64),
                      torch.randn(M, device=dev, generator=g, dtype=torch.float64))
    ms = timed(lambda: J.nufft1((N,), c, x, eps=1e-6, iflag=1))
    p = Plan(1, (N,), eps=1e-6, isign=1, dtype="complex128", debug=1)
    st = stages(p, lambda: p.setpts(x), lambda: p.execute(c[None]))
    p.destroy()
    return {"config": "C1 1-D type 1 M=1e6 N=1e6 eps=1e-6 c128", "ms_per_step": ms, "points_per_s": M / ms * 1e3,
            "stages_ms": st}


def c2():
    """2-D type 2, M=1e7, N=2048^2, eps=1e-5, complex64 (BASELINE configs[1])."""
    g = torch.Generator(device=dev).manual_seed(1)
    M, nm = 10 ** 7, (2048, 2048)
    x, y = rnd(M, g, -np.pi, np.pi), rnd(M, g, -np.pi, np.pi)
    f = cplx(nm, g)
    ms = timed(lambda: J.nufft2(f, x, y, eps=1e-5, iflag=-1))
    p = Plan(2, nm[::-1], eps=1e-5, isign=-1, debug=1)
    st = stages(p, lambda: p.setpts(y, x), lambda: p.execute(f[None]))
    p.destroy()
    return {"config": "C2 2-D type 2 M=1e7 N=2048^2 eps=1e-5 c64", "ms_per_step": ms, "points_per_s": M / ms * 1e3,
            "stages_ms": st}


def c4():
    """2-D type 1 stacked, n_transf=64 sharing points, M=1e7, N=1024^2, eps=1e-6 (configs[3], one GPU)."""
    g = torch.Generator(device=dev).manual_seed(3)
    M, nm, nt = 10 ** 7, (1024, 1024), 64
    x, y = rnd(M, g, -np.pi, np.pi), rnd(M, g, -np.pi, np.pi)
    c = cplx((nt, M), g)
    ms = timed(lambda: J.nufft1(nm, c, x, y, eps=1e-6, iflag=1), K=3, W=2)
    p = Plan(1, nm[::-1], n_trans=nt, eps=1e-6, isign=1, debug=1)
    st = stages(p, lambda: p.setpts(y, x), lambda: p.execute(c))
    p.destroy()
    return {"config": "C4 2-D type 1 x64 stacked M=1e7 N=1024^2 eps=1e-6 c64", "ms_per_step": ms,
            "points_per_s": nt * M / ms * 1e3, "stages_ms": st}


def c5(grad=False):
    """3-D type 3, M=1e7 sources -> 1e7 targets in [-64,64)^3, eps=1e-6 (configs[4]); grad=True adds the
    VJP w.r.t. strengths and source points (ops.py:238-273 of the reference)."""
    g = torch.Generator(device=dev).manual_seed(4)
    M = N = 10 ** 7
    x = [rnd(M, g, -np.pi, np.pi) for _ in range(3)]
    s = [rnd(N, g, -64.0, 64.0) for _ in range(3)]
    c = cplx(M, g)
    if not grad:
        ms = timed(lambda: J.nufft3(c, *x, *s, eps=1e-6, iflag=-1), K=3, W=2)
        p = Plan(3, 3, eps=1e-6, isign=-1, debug=1, upsampfac=2.0)  # jax-finufft default (options.py:66)
        st = stages(p, lambda: p.setpts(x[2], x[1], x[0], s[2], s[1], s[0]), lambda: p.execute(c[None]))
        p.destroy()
        return {"config": "C5 3-D type 3 M=1e7 -> N=1e7 (targets in [-64,64)^3) eps=1e-6 c64", "ms_per_step": ms,
                "points_per_s": M / ms * 1e3, "stages_ms": st}
    cr = c.clone().requires_grad_(True)
    xr = [t.clone().requires_grad_(True) for t in x]
    w = cplx(N, g)

    def step():
        out = J.nufft3(cr, *xr, *s, eps=1e-6, iflag=-1)
        loss = (out * w.conj()).real.sum()
        loss.backward()
        cr.grad = None
        for t in xr:
            t.grad = None

    ms = timed(step, K=2, W=1)
    return {"config": "C5 forward + VJP (strengths and source points)", "ms_per_step": ms, "points_per_s": M / ms * 1e3}


if __name__ == "__main__":
    which = sys.argv[1:] or ["c1", "c2", "c4", "c5", "c5grad"]
    for wname in which:
        fn = {"c1": c1, "c2": c2, "c4": c4, "c5": c5, "c5grad": lambda: c5(True)}[wname]
        try:
            r = fn()
        except Exception as e:  # keep going: one config failing must not hide the others
            r = {"config": wname, "error": repr(e)[:300]}
        print(json.dumps(r), flush=True)
        torch.cuda.empty_cache()
