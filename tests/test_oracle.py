"""Pins the CPU oracle (oracle/nufft_oracle.c): (1) against a float64 direct NUDFT with the
reference's own seeds / sizes / tolerances (tests/ops_test.py:25-126,
V/test/cuda/cufinufft3d_test.cu:186-254) and (2) against outputs of the unmodified reference
cuFINUFFT recorded on a B200 (tests/golden/ref_cufinufft_golden.npz, made by
tests/golden/make_golden.py).  CPU only."""
import os
import sys

import numpy as np
import pytest

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import cases as G  # noqa: E402

GOLDEN = os.path.join(HERE, "golden", "ref_cufinufft_golden.npz")


@pytest.mark.parametrize("ndim", [1, 2, 3])
@pytest.mark.parametrize("x64", [False, True])
@pytest.mark.parametrize("iflag", [-1, 1])
def test_oracle_vs_nudft_reference_seeds(ndim, x64, iflag):
    rng = np.random.default_rng(657)                      # ops_test.py:30
    eps = 1e-10 if x64 else 1e-7                          # ops_test.py:32
    rtol = 1e-7 if x64 else 1e-4                          # ops_test.py:14-16 (checked as relative l2)
    M = 50
    nm = tuple(int(v) for v in (75 // ndim + 5 * np.arange(ndim)))   # x-fastest order
    x = rng.uniform(-np.pi, np.pi, size=(ndim, M))
    c = rng.normal(size=M) + 1j * rng.normal(size=M)
    f = oracle.nufft1(nm, c, *x, eps=eps, iflag=iflag)
    assert oracle.relerr(f, oracle.dirft1(nm, c, *x, iflag=iflag)) < rtol
    fk = rng.normal(size=nm[::-1]) + 1j * rng.normal(size=nm[::-1])
    c2 = oracle.nufft2(fk, *x, eps=eps, iflag=iflag)
    assert oracle.relerr(c2, oracle.dirft2(fk, *x, iflag=iflag)) < rtol
    s = rng.uniform(-20, 20, size=(ndim, 20))              # type 3: 25 -> 20 (ops_test.py:100-116)
    f3 = oracle.nufft3(c[:25], list(x[:, :25]), list(s), eps=eps, iflag=iflag)
    assert oracle.relerr(f3, oracle.dirft3(c[:25], list(x[:, :25]), list(s), iflag=iflag)) < (1e-3 if not x64 else rtol)


@pytest.mark.parametrize("eps,check", [(1e-4, 2e-3), (1e-12, 1e-11)])
def test_oracle_cufinufft3d_test_matrix(eps, check):
    """V/test/cuda/cufinufft3d_test.cu: N=2x5x10, M=20, x in [-pi,pi), relative l2 vs dirft3d."""
    rng = np.random.default_rng(1)
    nm, M = (2, 5, 10), 20
    x = rng.uniform(-np.pi, np.pi, size=(3, M))
    c = rng.uniform(-1, 1, M) + 1j * rng.uniform(-1, 1, M)
    assert oracle.relerr(oracle.nufft1(nm, c, *x, eps=eps), oracle.dirft1(nm, c, *x)) < check
    fk = rng.uniform(-1, 1, nm[::-1]) + 1j * rng.uniform(-1, 1, nm[::-1])
    assert oracle.relerr(oracle.nufft2(fk, *x, eps=eps), oracle.dirft2(fk, *x)) < check
    s = rng.uniform(-np.pi, np.pi, size=(3, M)) * 3
    assert oracle.relerr(oracle.nufft3(c, list(x), list(s), eps=eps), oracle.dirft3(c, list(x), list(s))) < check


def test_oracle_modeord_and_sigma125():
    rng = np.random.default_rng(5)
    x = rng.uniform(-np.pi, np.pi, size=(2, 300))
    c = rng.normal(size=300) + 1j * rng.normal(size=300)
    for nm in ((21, 16), (20, 17)):
        for modeord in (0, 1):
            f = oracle.nufft1(nm, c, *x, eps=1e-9, modeord=modeord)
            assert oracle.relerr(f, oracle.dirft1(nm, c, *x, modeord=modeord)) < 1e-8
    f = oracle.nufft1((20, 30), c, *x, eps=1e-8, upsampfac=1.25)
    assert oracle.relerr(f, oracle.dirft1((20, 30), c, *x)) < 1e-7   # CTest: tol 1e-8 -> check 1e-7


def test_oracle_fold_rescale_rounding():
    """V/include/cufinufft/spreadinterp.h:30-57: result in [0, N); float path rounds down."""
    for prec in (0, 1):
        xs = np.array([-np.pi, np.pi, 0.0, 3 * np.pi, -3 * np.pi, 1e3, -1e3, np.nextafter(np.pi, 0), 7.5])
        r = oracle.fold_rescale(xs, 512, prec)
        assert (r >= 0).all() and (r < 512).all()
    assert oracle.fold_rescale(0.0, 512)[0] == 256.0


def test_oracle_binsort_contract():
    rng = np.random.default_rng(3)
    x = rng.uniform(-np.pi, np.pi, size=(3, 5000))
    binid, hist = oracle.binsort(list(x), (64, 48, 40), (16, 16, 8))
    assert hist.sum() == 5000 and hist.size == 4 * 3 * 5
    assert (np.bincount(binid, minlength=hist.size) == hist).all()
    # x-fastest numbering (3d/spreadinterp3d.cuh:41-52)
    bx = np.floor(oracle.fold_rescale(x[0], 64) / 16).astype(int)
    by = np.floor(oracle.fold_rescale(x[1], 48) / 16).astype(int)
    bz = np.floor(oracle.fold_rescale(x[2], 40) / 8).astype(int)
    assert (binid == bx + 4 * (by + 3 * bz)).all()


@pytest.mark.parametrize("case", G.CASES, ids=[c[0] for c in G.CASES])
def test_oracle_vs_reference_golden(case):
    if not os.path.exists(GOLDEN):
        pytest.fail("tests/golden/ref_cufinufft_golden.npz missing: run tests/golden/make_golden.py on a GPU box")
    name, typ, dim, nm, M, N, eps, dbl, iflag, ntr, sigma, modeord = case
    gold = np.load(GOLDEN)[name]
    inp = G.make_inputs(case)
    pts = [p.astype(np.float64) for p in inp["pts"]]
    prec = 0 if dbl else 1
    if typ == 1:
        out = oracle.nufft1(nm, inp["data"], *pts, iflag=iflag, eps=eps, upsampfac=sigma, modeord=modeord, prec=prec)
    elif typ == 2:
        out = oracle.nufft2(inp["data"], *pts, iflag=iflag, eps=eps, upsampfac=sigma, modeord=modeord, prec=prec)
    else:
        out = oracle.nufft3(inp["data"], pts, [s.astype(np.float64) for s in inp["tgt"]], iflag=iflag, eps=eps,
                            upsampfac=sigma, prec=prec)
    assert out.shape == gold.shape
    assert oracle.relerr(out, gold) < G.tolerance(case)
