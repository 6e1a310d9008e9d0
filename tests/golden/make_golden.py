"""Generate tests/golden/ref_cufinufft_golden.npz: outputs of the UNMODIFIED reference cuFINUFFT
(oracle/_ref/libcufinufft_ref.so, built from /root/reference by oracle/Makefile.ref) on the
seeded cases of cases.py.  Needs a GPU:

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/ref_cufinufft_golden.npz'
    cp gpurun_out/ref_cufinufft_golden.npz tests/golden/

Also records, per case, the reference's own error against a float64 direct NUDFT (the yardstick
of SURVEY.md §8c) so the tests can show ours is not worse than the reference.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import cases as G  # noqa: E402
import oracle  # noqa: E402
from oracle import ref_cufinufft as ref  # noqa: E402


def run_reference(case):
    name, typ, dim, nm, M, N, eps, dbl, iflag, ntr, sigma, modeord = case
    inp = G.make_inputs(case)
    dev = "cuda"
    tp = [torch.as_tensor(p, device=dev) for p in inp["pts"]]
    tt = [torch.as_tensor(p, device=dev) for p in inp["tgt"]]
    dt = "complex128" if dbl else "complex64"
    p = ref.RefPlan(typ, dim if typ == 3 else nm, n_trans=ntr, eps=eps, isign=iflag, dtype=dt,
                    upsampfac=sigma, modeord=modeord)
    p.setpts(*(tp + [None] * (3 - dim)), *(tt + [None] * (3 - len(tt))))
    out = p.execute(torch.as_tensor(inp["data"], device=dev)).cpu().numpy()
    p.destroy()
    return inp, out


def nudft(case, inp):
    name, typ, dim, nm, M, N, eps, dbl, iflag, ntr, sigma, modeord = case
    pts = [p.astype(np.float64) for p in inp["pts"]]
    outs = []
    for t in range(ntr):
        d = inp["data"][t].astype(np.complex128)
        if typ == 1:
            outs.append(oracle.dirft1(nm, d, *pts, iflag=iflag, modeord=modeord))
        elif typ == 2:
            outs.append(oracle.dirft2(d, *pts, iflag=iflag, modeord=modeord))
        else:
            outs.append(oracle.dirft3(d, pts, [s.astype(np.float64) for s in inp["tgt"]], iflag=iflag))
    return np.stack(outs)


if __name__ == "__main__":
    dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "ref_cufinufft_golden.npz")
    blob = {}
    for case in G.CASES:
        inp, out = run_reference(case)
        err = oracle.relerr(out, nudft(case, inp))
        blob[case[0]] = out
        blob[case[0] + "__ref_vs_nudft"] = np.float64(err)
        print(f"{case[0]:28s} shape {out.shape} ref-vs-NUDFT {err:.2e}")
    np.savez_compressed(dst, **blob)
    print("wrote", dst, os.path.getsize(dst), "bytes")
