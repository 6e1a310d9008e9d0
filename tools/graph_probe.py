"""CUDA-graph capture of the custom call: warm the plan cache, capture
nufft1 / nufft2 calls into a torch CUDA graph, replay with new data in the same buffers and compare
with eager calls.  Prints one line per case."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import jax_finufft_b200 as J
from jax_finufft_b200 import _lib


def rel(a, b):
    return float(torch.linalg.norm((a - b).flatten().to(torch.complex128)) / torch.linalg.norm(b.flatten().to(torch.complex128)))


def main():
    L = _lib.lib()
    dev = "cuda"
    for cache in (0, 1):
        L.b2n_set_setpts_cache(cache)
        L.b2n_cache_clear()
        for ndim, nm, M in ((2, (128, 96), 50000), (3, (48, 40, 36), 200000)):
            g = torch.Generator(device=dev).manual_seed(1)
            pts = [(torch.rand(M, generator=g, device=dev) * 2 - 1) * np.pi for _ in range(ndim)]
            c = torch.randn(M, generator=g, device=dev, dtype=torch.complex64)
            f_in = torch.randn(nm, generator=g, device=dev, dtype=torch.complex64)
            # warm-up on a side stream (plan creation, cuFFT plans, allocator pools)
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(3):
                    J.nufft1(nm, c, *pts, eps=1e-6)
                    J.nufft2(f_in, *pts, eps=1e-6)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            try:
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    f_out = J.nufft1(nm, c, *pts, eps=1e-6)
                    c_out = J.nufft2(f_in, *pts, eps=1e-6)
                # new data in the captured buffers
                c.copy_(torch.randn(M, generator=g, device=dev, dtype=torch.complex64))
                f_in.copy_(torch.randn(nm, generator=g, device=dev, dtype=torch.complex64))
                for p in pts:
                    p.copy_((torch.rand(M, generator=g, device=dev) * 2 - 1) * np.pi)
                gr.replay()
                torch.cuda.synchronize()
                e1 = rel(f_out, J.nufft1(nm, c, *pts, eps=1e-6))
                e2 = rel(c_out, J.nufft2(f_in, *pts, eps=1e-6))
                gr.replay()
                torch.cuda.synchronize()
                e3 = rel(f_out, J.nufft1(nm, c, *pts, eps=1e-6))
                print(f"graph cache={cache} dim={ndim}: ok rel1={e1:.2e} rel2={e2:.2e} replay2={e3:.2e}")
            except Exception as ex:  # noqa: BLE001
                print(f"graph cache={cache} dim={ndim}: FAILED {type(ex).__name__}: {str(ex)[:300]}")
                torch.cuda.synchronize()
    L.b2n_set_setpts_cache(0)


if __name__ == "__main__":
    sys.exit(main())
