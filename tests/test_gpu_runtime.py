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


def test_cache_is_bounded_private_and_clearable():
    """plan memory lives in a PRIVATE pool (the device default pool's release threshold
    is left alone), the cache is bounded by bytes, and the package can release it."""
    import ctypes as C

    import jax_finufft_b200 as J
    from jax_finufft_b200 import _lib

    import gc

    L = _lib.lib()
    gc.collect()
    J.clear_cache()
    _, u_live = J.cache_bytes()   # plans other tests of this process still hold
    thr = C.c_uint64(123)
    cudart = C.CDLL("libcudart.so.12")
    pool = C.c_void_p()
    assert cudart.cudaDeviceGetDefaultMemPool(C.byref(pool), torch.cuda.current_device()) == 0
    x = [torch.rand(200000, device="cuda") * 6 - 3 for _ in range(3)]
    c = torch.complex(torch.rand(200000, device="cuda"), torch.rand(200000, device="cuda"))
    J.nufft1((64, 64, 64), c, *x, eps=1e-6)
    torch.cuda.synchronize()
    assert cudart.cudaMemPoolGetAttribute(pool, 4, C.byref(thr)) == 0   # cudaMemPoolAttrReleaseThreshold
    assert thr.value != 2 ** 64 - 1, "the device default pool must not be reconfigured"
    r0, u0 = J.cache_bytes()
    assert r0 >= u0 > u_live
    one = u0 - u_live                           # one parked 64^3 plan
    prev = J.set_cache_limit(u_live + one // 2)  # smaller than one parked plan: only the newest stays
    for n in (48, 56, 72):
        J.nufft1((n, n, n), c, *x, eps=1e-6)
    torch.cuda.synchronize()
    _, u1 = J.cache_bytes()
    assert u1 - u_live < 3 * one, (u_live, one, u1)   # not four plans' worth
    J.set_cache_limit(prev)
    J.clear_cache()
    r2, u2 = J.cache_bytes()
    assert u2 == u_live and r2 < r0, (r2, u2, u_live, r0)


def test_empty_operands_may_be_null_and_warning_repeats_on_cache_hits():
    import jax_finufft_b200 as J
    from jax_finufft_b200 import _lib
    import ctypes as C

    # n_j == 0: zeros out, although torch hands over NULL data pointers for the empty arrays
    e = torch.empty(0, device="cuda")
    f = J.nufft1((8, 6), torch.empty(0, dtype=torch.complex64, device="cuda"), e, e, eps=1e-5)
    assert f.shape == (8, 6) and float(f.abs().max()) == 0.0
    s = torch.rand(50, device="cuda")
    f3 = J.nufft3(torch.empty(0, dtype=torch.complex64, device="cuda"), e, s, eps=1e-5)
    assert f3.shape == (50,) and float(f3.abs().max()) == 0.0
    # eps below float machine precision: warning code 1 on the first call AND on the cache hit
    L = _lib.lib()
    J.clear_cache()
    a = _lib.B2nFfiAttrs()
    a.eps, a.iflag, a.n_tot, a.n_transf, a.n_j, a.n_k_1, a.upsampfac, a.gpu_kerevalmeth, a.gpu_sort = 1e-9, 1, 1, 1, 100, 32, 2.0, 1, 1
    x = torch.rand(100, device="cuda")
    c = torch.complex(torch.rand(100, device="cuda"), torch.rand(100, device="cuda"))
    out = torch.empty(32, dtype=torch.complex64, device="cuda")
    ops = (C.c_void_p * 2)(c.data_ptr(), x.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert L.b2n_ffi_call(b"nufft1d1f", st, C.byref(a), ops, 2, C.c_void_p(out.data_ptr())) == 1
    assert L.b2n_ffi_call(b"nufft1d1f", st, C.byref(a), ops, 2, C.c_void_p(out.data_ptr())) == 1
    torch.cuda.synchronize()


def test_two_pass_sort_learns_to_skip_an_overflowing_attempt():
    """A strongly non-uniform set overflows the bucket regions of the two-pass sort every time; the
    overflow flag comes home asynchronously and later setpts calls on the plan skip the attempt
    (sort.cu: learned_skip).  Whatever path a call takes -- attempt + fallback, fallback alone, or the
    fallback on a uniform set presented while the hold is still on -- the transform must not change."""
    from jax_finufft_b200.plan import Plan

    M, nm = 4_000_000, (48, 48, 48)   # > 48 MB of records: the bucketed sort
    g = torch.Generator(device=DEV).manual_seed(11)
    uni = [(torch.rand(M, device=DEV, generator=g) * 2 - 1) * np.pi for _ in range(3)]
    clu = [-np.pi + torch.rand(M, device=DEV, generator=g) * (2 * np.pi * 6 / 96) for _ in range(3)]
    c = torch.complex(torch.rand(M, device=DEV, generator=g), torch.rand(M, device=DEV, generator=g))[None]
    p = Plan(1, nm, eps=1e-6, isign=1)
    outs = []
    for rep in range(5):
        p.setpts(*clu)
        outs.append(p.execute(c).clone())
        torch.cuda.synchronize()   # lets the flag of this call reach the host before the next setpts
    for o in outs[1:]:
        assert rel(o, outs[0]) < 2e-6
    p.setpts(*uni)                 # hold still on: a uniform set through the three-pass pipeline
    a = p.execute(c).clone()
    p.destroy()
    q = Plan(1, nm, eps=1e-6, isign=1)   # fresh plan: the two-pass attempt succeeds
    q.setpts(*uni)
    b = q.execute(c).clone()
    q.destroy()
    assert rel(a, b) < 2e-6
