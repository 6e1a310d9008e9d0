"""Runtime behaviour of the custom call beyond single-threaded eager use (SURVEY.md 8(b)
"Threading", 8(f).4): concurrent host threads (the reference module is built FREE_THREADED and
has tests/freethreaded_test.py) and capture into a CUDA graph (no host sync, no retained operand
pointers inside the call -- DESIGN.md section 1)."""
import concurrent.futures

import numpy as np
import pytest
import torch

import jax_finufft_b200 as J
from jax_finufft_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    a, b = a.to(torch.complex128).flatten(), b.to(torch.complex128).flatten()
    return float(torch.linalg.norm(a - b) / torch.linalg.norm(b))


def test_threaded_pool_nufft1():  # freethreaded_test.py:29-66, on streams of a thread pool
    n_transforms, n_points, modes, eps = 200, 1000, 64, 1e-7
    rng = np.random.default_rng(42)
    x = torch.as_tensor(rng.uniform(-np.pi, np.pi, size=(n_transforms, n_points)).astype(np.float32), device=DEV)
    c = torch.as_tensor((rng.normal(size=(n_transforms, n_points)) +
                         1j * rng.normal(size=(n_transforms, n_points))).astype(np.complex64), device=DEV)
    torch.cuda.synchronize()

    def compute(i):
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            out = J.nufft1(modes, c[i], x[i], eps=eps)
        s.synchronize()
        return out

    seq = [compute(i) for i in range(n_transforms)]
    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        par = list(ex.map(compute, range(n_transforms)))
    for a, b in zip(par, seq):
        assert rel(a, b) < 1e-5


def test_threaded_mixed_plans():
    """Different plan keys and the same plan key requested at once by several threads: a cached
    plan is owned by exactly one caller while checked out (csrc/api.cu plan cache)."""
    g = torch.Generator(device=DEV).manual_seed(7)
    jobs = []
    for k in range(24):
        ndim = 1 + k % 3
        M = 3000 + 500 * (k % 4)
        nm = [(96,), (40, 36), (20, 18, 16)][ndim - 1]
        pts = [(torch.rand(M, generator=g, device=DEV) * 2 - 1) * np.pi for _ in range(ndim)]
        cc = torch.randn(M, generator=g, device=DEV, dtype=torch.complex64)
        ff = torch.randn(nm, generator=g, device=DEV, dtype=torch.complex64)
        jobs.append((nm, cc, ff, pts))
    torch.cuda.synchronize()

    def run(job):
        nm, cc, ff, pts = job
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            a = J.nufft1(nm, cc, *pts, eps=1e-6)
            b = J.nufft2(ff, *pts, eps=1e-6)
        s.synchronize()
        return a, b

    seq = [run(j) for j in jobs]
    with concurrent.futures.ThreadPoolExecutor(max_workers=6) as ex:
        par = list(ex.map(run, jobs * 3))
    for i, (a, b) in enumerate(par):
        assert rel(a, seq[i % len(jobs)][0]) < 1e-5 and rel(b, seq[i % len(jobs)][1]) < 1e-5


@pytest.mark.parametrize("cache,warm", [(0, True), (1, True), (0, False)])
@pytest.mark.parametrize("ndim,nm,M", [(2, (128, 96), 50000), (3, (48, 40, 36), 200000)])
def test_cuda_graph_capture(ndim, nm, M, cache, warm):
    """Capture nufft1 + nufft2 (twice: the second pair re-uses the plans of the first inside the
    capture) into one CUDA graph, replay with new points and data in the same buffers, compare
    with eager calls.  warm=False: the plans are built while the stream is capturing."""
    L = _lib.lib()
    prev = L.b2n_set_setpts_cache(cache)
    L.b2n_cache_clear()
    try:
        g = torch.Generator(device=DEV).manual_seed(1)
        pts = [(torch.rand(M, generator=g, device=DEV) * 2 - 1) * np.pi for _ in range(ndim)]
        c = torch.randn(M, generator=g, device=DEV, dtype=torch.complex64)
        f_in = torch.randn(nm, generator=g, device=DEV, dtype=torch.complex64)
        s = torch.cuda.Stream()  # warm-up: plans, cuFFT work areas, allocator pools
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3 if warm else 0):
                J.nufft1(nm, c, *pts, eps=1e-6)
                J.nufft2(f_in, *pts, eps=1e-6)
            torch.fft.fft(c)  # torch's own lazy initialisations happen outside the capture
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            J.nufft1(nm, c, *pts, eps=1e-6)
            J.nufft2(f_in, *pts, eps=1e-6)
            f_out = J.nufft1(nm, c, *pts, eps=1e-6)
            c_out = J.nufft2(f_in, *pts, eps=1e-6)
        for rep in range(3):
            if rep != 1:  # replay 1 repeats the points of replay 0 (a setpts-cache hit when on)
                for p in pts:
                    p.copy_((torch.rand(M, generator=g, device=DEV) * 2 - 1) * np.pi)
            c.copy_(torch.randn(M, generator=g, device=DEV, dtype=torch.complex64))
            f_in.copy_(torch.randn(nm, generator=g, device=DEV, dtype=torch.complex64))
            gr.replay()
            torch.cuda.synchronize()
            assert rel(f_out, J.nufft1(nm, c, *pts, eps=1e-6)) < 5e-6, rep
            assert rel(c_out, J.nufft2(f_in, *pts, eps=1e-6)) < 5e-6, rep
    finally:
        L.b2n_set_setpts_cache(prev)
        L.b2n_cache_clear()
