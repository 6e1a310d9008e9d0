#!/usr/bin/env python
"""The UNMODIFIED reference cuFINUFFT (oracle/_ref, built by oracle/Makefile.ref) timed on the same
B200 on the headline workloads, driven the way jax-finufft drives it (lib/kernels.cc.cu:25-92:
makeplan + setpts + execute + destroy per call) and, more favourably, with the plan kept
(setpts + execute only, as V/perftest/cuda/cuperftest.cu does).  Next to it: this library on the
same tensors.  Test/measurement infrastructure only.  One JSON line per workload."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jax_finufft_b200 as J  # noqa: E402
from oracle import ref_cufinufft as ref  # noqa: E402

dev = torch.device("cuda:0")


def timed(step, K=3, W=1):
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


def run(typ, M, nm, eps, dist="uniform"):
    g = torch.Generator(device=dev).manual_seed(1)
    pts = []
    for d in range(3):
        u = torch.rand(M, device=dev, generator=g)
        pts.append((u * 2 - 1) * np.pi if dist == "uniform" else -np.pi + u * (8 * 2 * np.pi / (2 * nm[-1])))
    shape = (M,) if typ == 1 else nm
    data = torch.complex(torch.rand(shape, device=dev, generator=g) * 2 - 1, torch.rand(shape, device=dev, generator=g) * 2 - 1)
    isign = 1 if typ == 1 else -1
    out = {}

    def percall():
        p = ref.RefPlan(typ, nm[::-1], eps=eps, isign=isign)
        p.setpts(pts[2], pts[1], pts[0])
        r = p.execute(data[None])
        torch.cuda.synchronize()
        p.destroy()
        return r

    out["ref_percall_ms"] = timed(percall)
    p = ref.RefPlan(typ, nm[::-1], eps=eps, isign=isign)
    res = [None]

    def kept():
        p.setpts(pts[2], pts[1], pts[0])
        res[0] = p.execute(data[None], out=res[0])

    out["ref_plan_kept_ms"] = timed(kept)
    want = res[0][0].clone()
    kept()  # the reference against itself: float atomics in a different order = the fp32 noise floor
    torch.cuda.synchronize()
    out["rel_l2_ref_vs_ref_rerun"] = float((res[0][0] - want).abs().double().pow(2).sum().sqrt() / want.abs().double().pow(2).sum().sqrt())
    p.destroy()
    res[0] = None
    torch.cuda.empty_cache()
    if typ == 1:
        ours = lambda: J.nufft1(nm, data, *pts, eps=eps, iflag=isign)
    else:
        ours = lambda: J.nufft2(data, *pts, eps=eps, iflag=isign)
    out["ours_ms"] = timed(ours, K=5, W=2)
    got = ours()
    got2 = ours()
    out["rel_l2_ours_vs_ours_rerun"] = float((got2 - got).abs().double().pow(2).sum().sqrt() / got.abs().double().pow(2).sum().sqrt())
    del got2
    out["rel_l2_ours_vs_ref"] = float((got - want).abs().double().pow(2).sum().sqrt() / want.abs().double().pow(2).sum().sqrt())
    out["speedup_vs_ref_percall"] = out["ref_percall_ms"] / out["ours_ms"]
    out["speedup_vs_ref_plan_kept"] = out["ref_plan_kept_ms"] / out["ours_ms"]
    return out


if __name__ == "__main__":
    if not ref.available():
        print(json.dumps({"unavailable": "oracle/_ref/libcufinufft_ref.so not built"}))
        sys.exit(0)
    for name, typ, dist in (("c3_t1", 1, "uniform"), ("c3_t2", 2, "uniform"), ("c3_t1_clustered", 1, "clustered")):
        try:
            r = run(typ, 10 ** 8, (256, 256, 256), 1e-6, dist)
        except Exception as e:
            r = {"error": repr(e)[:300]}
        r["workload"] = name
        print(json.dumps(r), flush=True)
        torch.cuda.empty_cache()
