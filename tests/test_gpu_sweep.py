"""Seeded sweep of the register kernels (3-D sliding window, 2-D register tile, stacked 2-D) over
kernel widths, grid shapes, transform counts and mode orderings, against the CPU oracle
(float64 restatement of the reference algorithm, oracle/nufft_oracle.c).  Point sets are dense
enough that the register kernels are selected (checked through b2n_plan_info).  Tolerance:
relative l2 <= max(2 eps, fp32 rounding floor) = tests/golden/cases.py::parity_tol."""
import os
import sys

import numpy as np
import pytest
import torch

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases as G  # noqa: E402

pytestmark = pytest.mark.gpu


def T(a):
    return torch.as_tensor(np.ascontiguousarray(a), device="cuda")


def _run(typ, nm, M, eps, ntr, iflag, modeord, seed, seam=False, upsampfac=2.0):
    from jax_finufft_b200.plan import Plan

    dim = len(nm)
    rng = np.random.default_rng(seed)
    pts = [rng.uniform(-np.pi, np.pi, M).astype(np.float32) for _ in range(dim)]
    if seam:  # a third of the points pile up on the periodic seam, a few far outside [-pi, pi)
        k = M // 3
        for d in range(dim):
            pts[d][:k] = (np.pi + rng.uniform(-0.2, 0.2, k)).astype(np.float32)
            pts[d][k:k + 50] += np.float32(2 * np.pi * 3)
    if typ == 1:
        data = (rng.uniform(-1, 1, (ntr, M)) + 1j * rng.uniform(-1, 1, (ntr, M))).astype(np.complex64)
    else:
        data = (rng.uniform(-1, 1, (ntr,) + nm[::-1]) + 1j * rng.uniform(-1, 1, (ntr,) + nm[::-1])).astype(np.complex64)
    p = Plan(typ, nm, n_trans=ntr, eps=eps, isign=iflag, modeord=modeord, upsampfac=upsampfac)
    p.setpts(*[T(x) for x in pts])
    out = p.execute(T(data)).cpu().numpy()
    info = p.info()
    p.destroy()
    p64 = [x.astype(np.float64) for x in pts]
    if typ == 1:
        want = oracle.nufft1(nm, data, *p64, iflag=iflag, eps=eps, modeord=modeord, prec=1, upsampfac=upsampfac)
    else:
        want = oracle.nufft2(data, *p64, iflag=iflag, eps=eps, modeord=modeord, prec=1, upsampfac=upsampfac)
    return out, want, info


# (dim, nm (backend order, x first), M): grids with odd sizes, sizes that are not multiples of
# the bins, and the smallest grid the register kernels take (nf = 32)
GRIDS = [
    ((16, 16), 3000), ((37, 50), 20000), ((64, 23), 20000), ((128, 96), 150000),
    ((16, 16, 16), 20000), ((20, 33, 17), 40000), ((48, 18, 40), 120000),
]
EPS = [1e-2, 1e-3, 1e-4, 1e-5, 1e-6, 1e-7]   # ns = 3 .. 8


@pytest.mark.parametrize("typ", [1, 2])
@pytest.mark.parametrize("eps", EPS)
@pytest.mark.parametrize("nm,M", GRIDS, ids=[f"{'x'.join(map(str, g))}" for g, _ in GRIDS])
def test_register_kernels_vs_oracle(nm, M, eps, typ):
    seed = hash((nm, eps, typ)) % (2 ** 31)
    out, want, info = _run(typ, nm, M, eps, 1, 1 if typ == 1 else -1, 0, seed, seam=True)
    assert info.method == 3, "expected the register kernels for a dense float point set"
    tol = G.parity_tol(eps, False)   # max(2*eps, fp32 floor): golden/cases.py
    err = oracle.relerr(out, want)
    print(f"\nPARITY sweep {'x'.join(map(str, nm))} eps={eps:g} t{typ}: {err:.3e} ({err / eps:.2f} eps)")
    assert err < tol, (nm, eps, typ, err)


@pytest.mark.parametrize("ntr", [2, 3, 4, 5, 9])
@pytest.mark.parametrize("typ", [1, 2])
@pytest.mark.parametrize("dim", [2, 3])
def test_stacked_transforms_register_kernels(dim, typ, ntr):
    """vmap-stacked transforms sharing the points: the 2-D stacked spreader handles 4 per pass
    (remainders 1..3 exercised), the others run one transform per grid row."""
    nm = (40, 28) if dim == 2 else (24, 18, 20)
    out, want, info = _run(typ, nm, 30000, 1e-5, ntr, -1 if typ == 1 else 1, 1, 100 + ntr + 10 * dim, seam=False)
    assert info.method == 3
    assert out.shape == want.shape
    for t in range(ntr):
        assert oracle.relerr(out[t], want[t]) < G.parity_tol(1e-5, False), (dim, typ, ntr, t)


@pytest.mark.parametrize("eps", [1e-2, 1e-3, 1e-4, 1e-5, 1e-6])   # ns = 3 .. 7: every instantiation of k_rt2s_spread
@pytest.mark.parametrize("ntr", [2, 8, 11])
def test_stacked_2d_type1_narrow_window_all_widths(eps, ntr):
    """2-D type 1 with stacked transforms runs the narrow-window stacked spreader (8 transforms per
    pass, packed strengths): one full pass, a partial one, and more transforms than a batch holds."""
    out, want, info = _run(1, (44, 30), 40000, eps, ntr, 1, 0, 300 + ntr, seam=True)
    assert info.method == 3
    for t in range(ntr):
        err = oracle.relerr(out[t], want[t])
        assert err < G.parity_tol(eps, False), (eps, ntr, t, err)


@pytest.mark.parametrize("nm,M", [((128, 96), 120000), ((128, 96), 40000), ((44, 30), 40000)],
                         ids=["bins_of_1_2_batches", "sparse_bins", "bins_of_3_batches"])
@pytest.mark.parametrize("ntr", [16, 19])
def test_stacked_2d_type1_groups_of_batches(nm, M, ntr):
    """More stacked transforms than one batch (8): one spread launch takes the whole group and runs
    its passes back to back (api.cu:exec1_stacked, k_rt2s_spread).  Bins of <= 64 points keep their
    weight rows in shared memory across the passes, larger ones re-evaluate them; 19 = two full
    passes and a partial one.  Every transform is checked against the oracle."""
    out, want, info = _run(1, nm, M, 1e-6, ntr, 1, 0, 900 + ntr + M % 7, seam=True)
    assert info.method == 3
    worst = max(oracle.relerr(out[t], want[t]) for t in range(ntr))
    print(f"\nPARITY stacked groups {nm} M={M} ntr={ntr}: worst {worst:.3e}")
    assert worst < G.parity_tol(1e-6, False), worst


@pytest.mark.parametrize("typ", [1, 2])
@pytest.mark.parametrize("nm", [(40, 36), (40, 36, 32)])
def test_register_kernels_low_upsampling(nm, typ):
    """upsampfac = 1.25 gives kernel tables with another coefficient count than the ones the 3-D
    spreader unrolls: the run-time Horner loop of swr2_weights."""
    out, want, info = _run(typ, nm, 60000, 1e-4, 1, -1, 0, 77, seam=True, upsampfac=1.25)
    err = oracle.relerr(out, want)
    print(f"\nPARITY sigma=1.25 {'x'.join(map(str, nm))} t{typ}: {err:.3e} method {info.method}")
    assert info.method == 3, "expected the register kernels"
    assert err < G.parity_tol(1e-4, False), err
