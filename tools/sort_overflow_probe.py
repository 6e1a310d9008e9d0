#!/usr/bin/env python
"""Two-pass sort under overflow: clustered / half-clustered point sets large enough for the bucketed path,
checked against the three-pass pipeline (B2N_SORT_THREE_PASS is read once per process, so the reference
result comes from the oracle-free property: type 1 of the same data must agree between two plans)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from jax_finufft_b200.plan import Plan

M = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4_000_000
nm = (64, 64, 64)
g = torch.Generator(device="cuda").manual_seed(5)
for name, frac in (("uniform", 0.0), ("half", 0.5), ("clustered", 1.0)):
    k = int(M * frac)
    pts = []
    for d in range(3):
        u = (torch.rand(M, device="cuda", generator=g) * 2 - 1) * np.pi
        u[:k] = -np.pi + torch.rand(k, device="cuda", generator=g) * (8 * 2 * np.pi / 128)
        pts.append(u[torch.randperm(M, device="cuda", generator=g)] if d == 0 else u)
    c = torch.complex(torch.rand(M, device="cuda", generator=g), torch.rand(M, device="cuda", generator=g))[None]
    outs = []
    for rep in range(3):
        p = Plan(1, nm, eps=1e-6, isign=1)
        p.setpts(*pts)
        outs.append(p.execute(c).clone())
        idx, bs = p.sort_arrays()
        assert int(bs[-1]) == M and torch.equal(torch.sort(idx.long())[0], torch.arange(M, device="cuda"))
        p.destroy()
    torch.cuda.synchronize()
    e = float(torch.linalg.vector_norm(outs[1] - outs[0]) / torch.linalg.vector_norm(outs[0]))
    print(name, "ok, rerun diff", e)
