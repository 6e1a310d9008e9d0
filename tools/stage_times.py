#!/usr/bin/env python
"""Per-stage device times (CUDA events inside the library, b2n_plan_timings) of the C3 workload
for the library named by B2N_LIB: python tools/stage_times.py [M] [clustered]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from jax_finufft_b200.plan import Plan

M = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10 ** 8
clustered = len(sys.argv) > 2
nm, eps = (256, 256, 256), 1e-6
g = torch.Generator(device="cuda").manual_seed(1)
if clustered:
    pts = [-np.pi + torch.rand(M, device="cuda", generator=g) * (8 * 2 * np.pi / 512) for _ in range(3)]
else:
    pts = [(torch.rand(M, device="cuda", generator=g) * 2 - 1) * np.pi for _ in range(3)]
c = torch.complex(torch.rand(M, device="cuda", generator=g), torch.rand(M, device="cuda", generator=g))[None]
f = torch.complex(torch.rand(nm, device="cuda", generator=g), torch.rand(nm, device="cuda", generator=g))[None]
res = {}
for typ, data in ((1, c), (2, f)):
    p = Plan(typ, nm, eps=eps, isign=1 if typ == 1 else -1, debug=1)
    out = None
    K = 4
    for it in range(K + 1):
        p.setpts(*pts)
        out = p.execute(data, out=out)
        if it == 0:
            p.timings()
    st = {k: round(v / K, 3) for k, v in p.timings().items() if v}
    p.destroy()
    res[f"t{typ}"] = st
print(os.path.basename(os.environ.get("B2N_LIB", "libb200nufft.so")), res)
