"""Public-API tests on the GPU, mirroring the reference's tests/ops_test.py case by case
(same seeds, sizes, eps, tolerances; NumPy NUDFT as truth): forward nufft1/2/3 in 1-3-D for
c64/c128 and iflag +-1, JVP + VJP by finite differences / torch.autograd.gradcheck, vmap in its
four flavours, explicit stacked (n_tot, n_transf) inputs, modeord x parity of N, regressions."""
from itertools import product

import numpy as np
import pytest
import torch

import jax_finufft_b200 as J
from jax_finufft_b200 import Opts
from jax_finufft_b200.ops import get_frequency_array

pytestmark = pytest.mark.gpu
DEV = "cuda"


def T(a):
    return torch.as_tensor(np.ascontiguousarray(a), device=DEV)


def check_close(a, b, x64, **kw):  # ops_test.py:13-16
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else a
    b = b.detach().cpu().numpy() if torch.is_tensor(b) else b
    # rtol as ops_test.py:13-16; the tiny atol (relative to the largest entry) keeps one
    # near-zero element of a 26k-element output from failing an eps=1e-10 transform
    scale = float(np.abs(b).max()) if np.size(b) else 1.0
    np.testing.assert_allclose(a, b, **({"rtol": 1e-7, "atol": 1e-9 * scale} if x64 else {"rtol": 1e-4, "atol": 2e-5}) | kw)


def freq_grids(nm, modeord=0):
    return np.meshgrid(*[get_frequency_array(n, modeord) for n in nm], indexing="ij")


@pytest.mark.parametrize("ndim,x64,iflag", list(product([1, 2, 3], [False, True], [-1, 1])))
def test_nufft1_forward(ndim, x64, iflag):  # ops_test.py:25-56
    rng = np.random.default_rng(657)
    eps = 1e-10 if x64 else 1e-7
    rd, cd = (np.float64, np.complex128) if x64 else (np.float32, np.complex64)
    M = 50
    nm = tuple(int(v) for v in (75 // ndim + 5 * np.arange(ndim)))
    x = [rng.uniform(-np.pi, np.pi, M).astype(rd) for _ in range(ndim)]
    c = (rng.normal(size=M) + 1j * rng.normal(size=M)).astype(cd)
    ks = freq_grids(nm)
    f_expect = np.zeros(nm, dtype=cd)
    for n in range(M):
        f_expect += c[n] * np.exp(1j * iflag * sum(k * xx[n] for k, xx in zip(ks, x)))
    f = J.nufft1(nm, T(c), *map(T, x), eps=eps, iflag=iflag)
    assert f.dtype == (torch.complex128 if x64 else torch.complex64) and tuple(f.shape) == nm
    check_close(f, f_expect, x64, **({} if x64 else {"atol": 1e-4}))


@pytest.mark.parametrize("ndim,x64,iflag", list(product([1, 2, 3], [False, True], [-1, 1])))
def test_nufft2_forward(ndim, x64, iflag):  # ops_test.py:59-92
    rng = np.random.default_rng(657)
    eps = 1e-10 if x64 else 1e-7
    rd, cd = (np.float64, np.complex128) if x64 else (np.float32, np.complex64)
    M = 50
    nm = tuple(int(v) for v in (75 // ndim + 5 * np.arange(ndim)))
    x = [rng.uniform(-np.pi, np.pi, M).astype(rd) for _ in range(ndim)]
    f = (rng.normal(size=nm) + 1j * rng.normal(size=nm)).astype(cd)
    ks = freq_grids(nm)
    c_expect = np.zeros(M, dtype=cd)
    for n in range(M):
        c_expect[n] = np.sum(f * np.exp(1j * iflag * sum(k * xx[n] for k, xx in zip(ks, x))))
    c = J.nufft2(T(f), *map(T, x), eps=eps, iflag=iflag)
    assert tuple(c.shape) == (M,)
    check_close(c, c_expect, x64, **({} if x64 else {"atol": 2e-4}))


@pytest.mark.parametrize("ndim,x64,iflag", list(product([1, 2, 3], [False, True], [-1, 1])))
def test_nufft3_forward(ndim, x64, iflag):  # ops_test.py:95-126
    rng = np.random.default_rng(657)
    eps = 1e-10 if x64 else 1e-7
    rd, cd = (np.float64, np.complex128) if x64 else (np.float32, np.complex64)
    M, N = 25, 20
    x = [rng.uniform(-1.0, 1.0, M).astype(rd) for _ in range(ndim)]
    s = [rng.uniform(-1.0, 1.0, N).astype(rd) for _ in range(ndim)]
    c = (rng.normal(size=M) + 1j * rng.normal(size=M)).astype(cd)
    f_expect = np.zeros(N, dtype=cd)
    for k in range(N):
        f_expect[k] = np.sum(c * np.exp(1j * iflag * sum(ss[k] * xx for ss, xx in zip(s, x))))
    f = J.nufft3(T(c), *map(T, x), *map(T, s), eps=eps, iflag=iflag)
    assert tuple(f.shape) == (N,)
    check_close(f, f_expect, x64, **({} if x64 else {"rtol": 1e-3, "atol": 1e-4}))


def _gradcheck(func, args):
    """First-order forward- and reverse-mode check by finite differences (jtu.check_grads order 1,
    modes fwd+rev, ops_test.py:150)."""
    args = [a.clone().requires_grad_(True) for a in args]
    assert torch.autograd.gradcheck(func, args, eps=1e-6, atol=1e-5, rtol=1e-4, check_forward_ad=True,
                                    check_backward_ad=True, fast_mode=True, check_batched_grad=False,
                                    check_batched_forward_grad=False, nondet_tol=1e-9)


@pytest.mark.parametrize("ndim,iflag", list(product([1, 2, 3], [-1, 1])))
def test_nufft1_grad(ndim, iflag):  # ops_test.py:128-158
    rng = np.random.default_rng(657)
    nm = tuple(int(v) for v in (35 // ndim + 5 * np.arange(ndim)))
    x = [T(rng.uniform(-np.pi, np.pi, 50)) for _ in range(ndim)]
    c = T(rng.normal(size=50) + 1j * rng.normal(size=50))
    _gradcheck(lambda c_, *x_: J.nufft1(nm, c_, *x_, eps=1e-10, iflag=iflag), [c, *x])


@pytest.mark.parametrize("ndim,iflag", list(product([1, 2, 3], [-1, 1])))
def test_nufft2_grad(ndim, iflag):  # ops_test.py:161-190
    rng = np.random.default_rng(657)
    nm = tuple(int(v) for v in (35 // ndim + 5 * np.arange(ndim)))
    x = [T(rng.uniform(-np.pi, np.pi, 50)) for _ in range(ndim)]
    f = T(rng.normal(size=nm) + 1j * rng.normal(size=nm))
    _gradcheck(lambda f_, *x_: J.nufft2(f_, *x_, eps=1e-10, iflag=iflag), [f, *x])


@pytest.mark.parametrize("ndim,iflag", list(product([1, 2, 3], [-1, 1])))
def test_nufft3_grad(ndim, iflag):  # ops_test.py:193-220
    rng = np.random.default_rng(657)
    x = [T(rng.uniform(-1.0, 1.0, 50)) for _ in range(ndim)]
    s = [T(rng.uniform(-1.0, 1.0, 35)) for _ in range(ndim)]
    c = T(rng.normal(size=50) + 1j * rng.normal(size=50))
    _gradcheck(lambda c_, *p: J.nufft3(c_, *p, eps=1e-10, iflag=iflag), [c, *x, *s])


@pytest.mark.parametrize("ndim,iflag", list(product([1, 2, 3], [-1, 1])))
def test_nufft1_vmap(ndim, iflag):  # ops_test.py:222-262
    rng = np.random.default_rng(657)
    R, M = 5, 50
    nm = tuple(int(v) for v in (35 // ndim + 5 * np.arange(ndim)))
    x = [T(rng.uniform(-np.pi, np.pi, (R, M))) for _ in range(ndim)]
    c = T(rng.normal(size=(R, M)) + 1j * rng.normal(size=(R, M)))
    func = lambda c_, *x_: J.nufft1(nm, c_, *x_, iflag=iflag)
    expect = torch.stack([func(c[i], *[xx[i] for xx in x]) for i in range(R)])
    got = torch.vmap(func)(c, *x)
    check_close(got, expect, True)
    # unmapped source
    got = torch.vmap(func, in_dims=(None,) + (0,) * ndim)(c[0], *x)
    expect = torch.stack([func(c[0], *[xx[i] for xx in x]) for i in range(R)])
    check_close(got, expect, True)
    # unmapped points: the mapped axis is folded into n_transf (one bin-sort) -- ops.py:323-326
    got = torch.vmap(func, in_dims=(0,) + (None,) * ndim)(c, *[xx[0] for xx in x])
    expect = torch.stack([func(c[i], *[xx[0] for xx in x]) for i in range(R)])
    check_close(got, expect, True)
    if ndim > 1:  # one point axis unmapped
        got = torch.vmap(func, in_dims=(0, None) + (0,) * (ndim - 1))(c, x[0][0], *x[1:])
        expect = torch.stack([func(c[i], x[0][0], *[xx[i] for xx in x[1:]]) for i in range(R)])
        check_close(got, expect, True)


@pytest.mark.parametrize("ndim,iflag", list(product([1, 2, 3], [-1, 1])))
def test_nufft2_vmap(ndim, iflag):  # ops_test.py:265-318
    rng = np.random.default_rng(657)
    R, M = 5, 50
    nm = tuple(int(v) for v in (35 // ndim + 5 * np.arange(ndim)))
    x = [T(rng.uniform(-np.pi, np.pi, (R, M))) for _ in range(ndim)]
    f = T(rng.normal(size=(R,) + nm) + 1j * rng.normal(size=(R,) + nm))
    func = lambda f_, *x_: J.nufft2(f_, *x_, iflag=iflag)
    expect = torch.stack([func(f[i], *[xx[i] for xx in x]) for i in range(R)])
    check_close(torch.vmap(func)(f, *x), expect, True)
    # permuted in_axes (ops_test.py:290-300)
    got = torch.vmap(func, in_dims=(ndim,) + (0,) * ndim)(torch.movedim(f, 0, ndim), *x)
    check_close(got, expect, True)
    got = torch.vmap(func, in_dims=(0,) + (None,) * ndim)(f, *[xx[0] for xx in x])
    expect = torch.stack([func(f[i], *[xx[0] for xx in x]) for i in range(R)])
    check_close(got, expect, True)


@pytest.mark.parametrize("ndim", [1, 2, 3])
def test_nufft3_vmap(ndim):  # ops_test.py:321-378
    rng = np.random.default_rng(657)
    R, M, N = 5, 50, 35
    x = [T(rng.uniform(-1, 1, (R, M))) for _ in range(ndim)]
    s = [T(rng.uniform(-1, 1, (R, N))) for _ in range(ndim)]
    c = T(rng.normal(size=(R, M)) + 1j * rng.normal(size=(R, M)))
    func = lambda c_, *p: J.nufft3(c_, *p)
    expect = torch.stack([func(c[i], *[p[i] for p in x + s]) for i in range(R)])
    check_close(torch.vmap(func)(c, *x, *s), expect, True)
    got = torch.vmap(func, in_dims=(0,) + (None,) * (2 * ndim))(c, *[p[0] for p in x + s])
    expect = torch.stack([func(c[i], *[p[0] for p in x + s]) for i in range(R)])
    check_close(got, expect, True)


def test_multi_transform():  # ops_test.py:380-398
    rng = np.random.default_rng(314)
    n_tot, n_tr, n_j, n_k = 4, 10, 100, 12
    f_shape = (n_tot, n_tr, n_k, n_k)
    c = T(rng.normal(size=(n_tot, n_tr, n_j)) + 1j * rng.normal(size=(n_tot, n_tr, n_j)))
    x = T(rng.uniform(-np.pi, np.pi, (n_tot, n_j)))
    y = T(rng.uniform(-np.pi, np.pi, (n_tot, n_j)))
    f = J.nufft1(f_shape[-2:], c, x[:, None], y[:, None])
    assert tuple(f.shape) == f_shape
    for i in range(n_tot):
        for j in range(n_tr):
            check_close(f[i, j], J.nufft1(f_shape[-2:], c[i, j], x[i], y[i]), True)
    c2 = J.nufft2(f, x[:, None], y[:, None])
    assert tuple(c2.shape) == (n_tot, n_tr, n_j)
    check_close(c2[2, 3], J.nufft2(f[2, 3], x[2], y[2]), True)


@pytest.mark.parametrize("ndim,nufft_type,modeord,even", list(product([1, 2, 3], [1, 2], [0, 1], [True, False])))
def test_modeord_values_and_grads(ndim, nufft_type, modeord, even):  # ops_test.py:494-567
    rng = np.random.default_rng(657)
    M = 40
    nm = tuple(int(v) + (0 if even else 1) for v in (12 + 2 * np.arange(ndim)))
    x = [rng.uniform(-np.pi, np.pi, M) for _ in range(ndim)]
    opts = Opts(modeord=modeord)
    ks = freq_grids(nm, modeord)
    if nufft_type == 1:
        c = rng.normal(size=M) + 1j * rng.normal(size=M)
        expect = np.zeros(nm, dtype=np.complex128)
        for n in range(M):
            expect += c[n] * np.exp(1j * sum(k * xx[n] for k, xx in zip(ks, x)))
        got = J.nufft1(nm, T(c), *map(T, x), eps=1e-10, iflag=1, opts=opts)
        check_close(got, expect, True)
        _gradcheck(lambda c_, *x_: J.nufft1(nm, c_, *x_, eps=1e-10, iflag=1, opts=opts), [T(c), *map(T, x)])
    else:
        f = rng.normal(size=nm) + 1j * rng.normal(size=nm)
        expect = np.array([np.sum(f * np.exp(-1j * sum(k * xx[n] for k, xx in zip(ks, x)))) for n in range(M)])
        got = J.nufft2(T(f), *map(T, x), eps=1e-10, iflag=-1, opts=opts)
        check_close(got, expect, True)
        _gradcheck(lambda f_, *x_: J.nufft2(f_, *x_, eps=1e-10, iflag=-1, opts=opts), [T(f), *map(T, x)])


def test_issue14_and_issue37_regressions():  # ops_test.py:400-450 (shape / finiteness pins)
    rng = np.random.default_rng(1)
    M, N = 100, 200
    x = T(rng.uniform(-np.pi, np.pi, M))
    f = T(rng.normal(size=N) + 1j * rng.normal(size=N)).requires_grad_(True)
    c = J.nufft2(f, x)
    (c.abs() ** 2).sum().backward()
    assert f.grad.shape == f.shape and torch.isfinite(f.grad.abs()).all()
    # batched points + gradient wrt points for type 3 (ops_test.py:570-595)
    xs = T(rng.uniform(-1, 1, (3, 30))).requires_grad_(True)
    ss = T(rng.uniform(-1, 1, (3, 20)))
    cc = T(rng.normal(size=(3, 30)) + 1j * rng.normal(size=(3, 30)))
    out = torch.vmap(lambda c_, x_, s_: J.nufft3(c_, x_, s_))(cc, xs, ss)
    out.abs().sum().backward()
    assert xs.grad.shape == xs.shape and torch.isfinite(xs.grad).all()


def test_cpu_tensors_are_rejected_loudly():
    x = torch.zeros(10)
    c = torch.zeros(10, dtype=torch.complex64)
    with pytest.raises((ValueError, RuntimeError)):
        J.nufft1((8,), c, x)


def test_float32_gradient_of_stacked_transform_shares_points():
    """The D-dim gradient is a D(+1)-deep stacked transform sharing points (ops.py:238-273):
    check the c64 VJP against the c128 one."""
    rng = np.random.default_rng(3)
    M, nm = 3000, (20, 24, 16)
    x = [rng.uniform(-np.pi, np.pi, M) for _ in range(3)]
    c = rng.normal(size=M) + 1j * rng.normal(size=M)
    w = rng.normal(size=nm) + 1j * rng.normal(size=nm)
    grads = []
    for rd, cd in ((torch.float64, torch.complex128), (torch.float32, torch.complex64)):
        xt = [T(v).to(rd).requires_grad_(True) for v in x]
        ct = T(c).to(cd).requires_grad_(True)
        f = J.nufft1(nm, ct, *xt, eps=1e-6)
        (f * T(w).to(cd)).real.sum().backward()
        grads.append([ct.grad.cpu().numpy()] + [v.grad.cpu().numpy() for v in xt])
    for a, b in zip(*grads):
        assert np.linalg.norm(a - b) / np.linalg.norm(a) < 1e-4
