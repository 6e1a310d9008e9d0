#!/usr/bin/env python
"""Reference cuFINUFFT (oracle/_ref) vs this library on secondary workloads: double precision,
2-D, stacked.  Plan kept for the reference (setpts + execute), public API for ours.  One JSON line each."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jax_finufft_b200 as J  # noqa: E402
from oracle import ref_cufinufft as ref  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "tools"))
from bench_ref_gpu import timed  # noqa: E402

dev = torch.device("cuda:0")


def run(name, typ, dim, nm, M, eps, dbl, ntr=1):
    g = torch.Generator(device=dev).manual_seed(7)
    rdt = torch.float64 if dbl else torch.float32
    pts = [(torch.rand(M, device=dev, generator=g, dtype=rdt) * 2 - 1) * np.pi for _ in range(dim)]
    shape = ((ntr, M) if ntr > 1 else (M,)) if typ == 1 else (((ntr,) + nm) if ntr > 1 else nm)
    data = torch.complex(torch.rand(shape, device=dev, generator=g, dtype=rdt) * 2 - 1,
                         torch.rand(shape, device=dev, generator=g, dtype=rdt) * 2 - 1)
    isign = 1 if typ == 1 else -1
    p = ref.RefPlan(typ, nm[::-1], n_trans=ntr, eps=eps, isign=isign, dtype="complex128" if dbl else "complex64")
    res = [None]
    d2 = data if ntr > 1 else data[None]

    def kept():
        p.setpts(*pts[::-1])
        res[0] = p.execute(d2, out=res[0])

    out = {"workload": name, "ref_plan_kept_ms": timed(kept)}
    want = res[0].clone()
    p.destroy()
    ours = (lambda: J.nufft1(nm, data, *pts, eps=eps, iflag=isign)) if typ == 1 else (lambda: J.nufft2(data, *pts, eps=eps, iflag=isign))
    out["ours_ms"] = timed(ours, K=5, W=2)
    got = ours().reshape(want.shape)
    out["rel_l2_ours_vs_ref"] = float((got - want).abs().double().pow(2).sum().sqrt() / want.abs().double().pow(2).sum().sqrt())
    out["speedup_vs_ref_plan_kept"] = out["ref_plan_kept_ms"] / out["ours_ms"]
    return out


CASES = [
    ("3d_t1_f64_M1e7_N128_eps1e-9", 1, 3, (128, 128, 128), 10 ** 7, 1e-9, True, 1),
    ("3d_t2_f64_M1e7_N128_eps1e-9", 2, 3, (128, 128, 128), 10 ** 7, 1e-9, True, 1),
    ("2d_t1_f64_M1e7_N1024_eps1e-9", 1, 2, (1024, 1024), 10 ** 7, 1e-9, True, 1),
    ("2d_t2_f64_M1e7_N1024_eps1e-9", 2, 2, (1024, 1024), 10 ** 7, 1e-9, True, 1),
    ("2d_t1_f32_M1e7_N1024_eps1e-6", 1, 2, (1024, 1024), 10 ** 7, 1e-6, False, 1),
    ("2d_t2_f32_M1e7_N2048_eps1e-5 (C2)", 2, 2, (2048, 2048), 10 ** 7, 1e-5, False, 1),
    ("2d_t1_f32_x16_M1e7_N1024_eps1e-6 (C4/4)", 1, 2, (1024, 1024), 10 ** 7, 1e-6, False, 16),
    ("1d_t1_f64_M1e6_N1e6_eps1e-6 (C1)", 1, 1, (10 ** 6,), 10 ** 6, 1e-6, True, 1),
    ("3d_t1_f32_M1e7_N128_eps1e-3", 1, 3, (128, 128, 128), 10 ** 7, 1e-3, False, 1),
]
if __name__ == "__main__":
    if not ref.available():
        print(json.dumps({"unavailable": "oracle/_ref/libcufinufft_ref.so not built"})); sys.exit(0)
    for c in CASES:
        try:
            r = run(*c)
        except Exception as e:
            r = {"workload": c[0], "error": repr(e)[:300]}
        print(json.dumps(r), flush=True)
        torch.cuda.empty_cache()
