"""Setpts cache (include/b200nufft.h: b2n_set_setpts_cache).

The reference re-sorts the points on every custom call (lib/kernels.cc.cu:49-51,64).  With the
cache on, a call whose coordinates equal the ones the cached plan holds sorted skips the bin-sort
(decided on the device from a 64-bit signature).  These tests pin the contract: results with the
cache on equal the results with it off -- for repeated points, changed points (one coordinate of
one point), permuted points, a different point count -- for every kernel family the sort feeds
(GM 1-D, register tile 2-D, sliding window 3-D, tile kernels f64, type 2, type 3), and the skipped
sort is measurably cheaper on the device.
"""
import numpy as np
import pytest
import torch

import jax_finufft_b200 as J
from jax_finufft_b200 import _lib
from jax_finufft_b200.plan import Plan

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture()
def cache_on():
    L = _lib.lib()
    L.b2n_cache_clear()
    prev = L.b2n_set_setpts_cache(1)
    yield L
    L.b2n_set_setpts_cache(prev)
    L.b2n_cache_clear()


def relerr(a, b):
    a = a.to(torch.complex128)
    b = b.to(torch.complex128)
    return float(torch.linalg.norm((a - b).flatten()) / torch.linalg.norm(b.flatten()))


def make_case(ndim, M, x64, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    rd, cd = (torch.float64, torch.complex128) if x64 else (torch.float32, torch.complex64)
    pts = [((torch.rand(M, generator=g, device=DEV, dtype=torch.float64) * 2 - 1) * np.pi).to(rd) for _ in range(ndim)]
    c = torch.complex(torch.randn(M, generator=g, device=DEV, dtype=torch.float64),
                      torch.randn(M, generator=g, device=DEV, dtype=torch.float64)).to(cd)
    return pts, c


CASES = [  # (ndim, modes, M, x64, eps): one per kernel family behind the sort
    (1, (4096,), 20000, True, 1e-9),          # GM kernels (1-D)
    (2, (192, 160), 60000, False, 1e-6),      # register-tile kernels
    (3, (48, 40, 36), 120000, False, 1e-6),   # sliding-window kernels
    (3, (24, 20, 28), 30000, True, 1e-10),    # tile kernels, double
    (3, (64, 64, 64), 3000, False, 1e-5),     # sparse 3-D float -> tile kernels
]


@pytest.mark.parametrize("ndim,nm,M,x64,eps", CASES)
def test_cache_on_equals_cache_off(cache_on, ndim, nm, M, x64, eps):
    L = cache_on
    tol = 1e-12 if x64 else 2e-6  # run-to-run noise of the float atomics, nothing more
    pts, c = make_case(ndim, M, x64, 11)
    pts_b = [p.clone() for p in pts]
    pts_b[0][M // 3] = pts_b[0][M // 3] + 1.0  # ONE coordinate of ONE point: a different bin
    perm = torch.randperm(M, device=DEV)
    pts_p, c_p = [p[perm].contiguous() for p in pts], c[perm].contiguous()
    pts_s, c_s = [p[: M - 7].contiguous() for p in pts], c[: M - 7].contiguous()
    f2 = torch.complex(torch.randn(nm, device=DEV, dtype=torch.float64),
                       torch.randn(nm, device=DEV, dtype=torch.float64)).to(c.dtype)

    L.b2n_set_setpts_cache(0)
    ref = {
        "a": J.nufft1(nm, c, *pts, eps=eps), "b": J.nufft1(nm, c, *pts_b, eps=eps),
        "s": J.nufft1(nm, c_s, *pts_s, eps=eps),
        "a2": J.nufft2(f2, *pts, eps=eps), "b2": J.nufft2(f2, *pts_b, eps=eps),
        "p2": J.nufft2(f2, *pts_p, eps=eps),
    }
    L.b2n_cache_clear()
    L.b2n_set_setpts_cache(1)
    # type 1: first call sorts, second and third are cache hits, then the point set changes
    for k in range(3):
        assert relerr(J.nufft1(nm, c, *pts, eps=eps), ref["a"]) < tol, k
    assert relerr(J.nufft1(nm, c, *pts_b, eps=eps), ref["b"]) < tol      # one coordinate differs
    assert relerr(J.nufft1(nm, c, *pts_b, eps=eps), ref["b"]) < tol      # hit on the new set
    assert relerr(J.nufft1(nm, c, *pts, eps=eps), ref["a"]) < tol        # and back
    assert relerr(J.nufft1(nm, c_p, *pts_p, eps=eps), ref["a"]) < tol    # same set, other order
    assert relerr(J.nufft1(nm, c, *pts, eps=eps), ref["a"]) < tol
    assert relerr(J.nufft1(nm, c_s, *pts_s, eps=eps), ref["s"]) < tol    # other M
    assert relerr(J.nufft1(nm, c, *pts, eps=eps), ref["a"]) < tol
    # same coordinates in arrays that are not 16-byte aligned (scalar signature kernel; the sum is
    # commutative, so the signature -- and the cache hit -- is the same)
    mis = []
    for q in pts:
        buf = torch.empty(M + 1, dtype=q.dtype, device=DEV)
        buf[1:].copy_(q)
        mis.append(buf[1:])
    assert mis[0].data_ptr() % 16 != 0
    for k in range(2):
        assert relerr(J.nufft1(nm, c, *mis, eps=eps), ref["a"]) < tol, k
    # type 2 has its own cached plan (and its own sorted copy)
    for k in range(2):
        assert relerr(J.nufft2(f2, *pts, eps=eps), ref["a2"]) < tol, k
    assert relerr(J.nufft2(f2, *pts_b, eps=eps), ref["b2"]) < tol
    assert relerr(J.nufft2(f2, *pts_p, eps=eps), ref["p2"]) < tol
    assert relerr(J.nufft2(f2, *pts, eps=eps), ref["a2"]) < tol


def test_cache_type3(cache_on):
    L = cache_on
    M, N = 20000, 15000
    pts, c = make_case(3, M, False, 5)
    g = torch.Generator(device=DEV).manual_seed(6)
    tg = [((torch.rand(N, generator=g, device=DEV) * 2 - 1) * 20.0) for _ in range(3)]
    tg_b = [t.clone() for t in tg]
    tg_b[1][17] += 0.5
    L.b2n_set_setpts_cache(0)
    ra = J.nufft3(c, *pts, *tg, eps=1e-6)
    rb = J.nufft3(c, *pts, *tg_b, eps=1e-6)
    L.b2n_cache_clear()
    L.b2n_set_setpts_cache(1)
    for k in range(3):
        assert relerr(J.nufft3(c, *pts, *tg, eps=1e-6), ra) < 2e-6
    assert relerr(J.nufft3(c, *pts, *tg_b, eps=1e-6), rb) < 2e-6
    assert relerr(J.nufft3(c, *pts, *tg, eps=1e-6), ra) < 2e-6


def test_gradient_step_shares_points(cache_on):
    """forward + VJP on fixed points, as a solver iteration does: equal gradients with the cache on."""
    L = cache_on
    nm = (40, 36, 32)
    pts, c = make_case(3, 50000, False, 3)

    def step():
        cc = c.clone().requires_grad_(True)
        xs = [p.clone().requires_grad_(True) for p in pts]
        f = J.nufft1(nm, cc, *xs, eps=1e-6)
        (f.abs() ** 2).sum().backward()
        return [cc.grad] + [x.grad for x in xs]

    L.b2n_set_setpts_cache(0)
    ref = step()
    L.b2n_cache_clear()
    L.b2n_set_setpts_cache(1)
    for _ in range(3):
        got = step()
        for a, b in zip(got, ref):
            assert relerr(a, b) < 5e-6


def test_cached_setpts_is_cheaper():
    """Device time of setpts with a signature hit vs a full bin-sort (plan-level, debug timers)."""
    L = _lib.lib()
    prev = L.b2n_set_setpts_cache(1)
    try:
        M = 8_000_000
        pts, _ = make_case(3, M, False, 9)
        other, _ = make_case(3, M, False, 10)
        p = Plan(1, (128, 128, 128), eps=1e-6, debug=1)
        p.setpts(*pts); torch.cuda.synchronize(); p.timings()
        full, hit = [], []
        for k in range(3):
            p.setpts(*other); torch.cuda.synchronize(); full.append(p.timings()["sort"])
            p.setpts(*pts); torch.cuda.synchronize(); full.append(p.timings()["sort"])
            p.setpts(*pts); torch.cuda.synchronize(); hit.append(p.timings()["sort"])
        p.destroy()
        assert min(hit) < 0.5 * min(full), (hit, full)
    finally:
        L.b2n_set_setpts_cache(prev)
