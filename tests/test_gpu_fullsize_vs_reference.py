"""Full-size parity against the LIVE reference cuFINUFFT (oracle/_ref) on BASELINE.json's configs:
C2 (2-D type 2), C3 (3-D type 1 / type 2, uniform and clustered), C4 (64 stacked 2-D type 1),
C5 (3-D type 3, targets in [-64, 64)^3).  Same tensors through both libraries, full-array
relative l2 (V/test/utils/norms.hpp:15-37) in float64 on the device.

Bound (north star): ours<->ref <= 2*eps.  The reference accumulates with float atomics and so
differs from ITSELF between two runs on the same inputs; that self-difference is measured in the
same test and the bound is max(2*eps, 1.2 * ref<->ref) -- no blanket fp32 floor.  Each test
prints a ``PARITY`` line with the numbers.  Mirrors V/test/cuda/cufinufft3d_test.cu:186-254
(library vs library instead of library vs direct sum, which is infeasible at these sizes; a
sampled float64 NUDFT is the yardstick where the fp32 arithmetic, not eps, sets the error).
"""
import numpy as np
import pytest
import torch

from oracle import ref_cufinufft as ref

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")]


def relerr(a, b):
    a = a.reshape(-1).to(torch.complex128)
    b = b.reshape(-1).to(torch.complex128)
    return float(torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b))


def urand(g, *shape):
    return torch.rand(*shape, device="cuda", generator=g)


def cplx(g, *shape):
    return torch.complex(urand(g, *shape) * 2 - 1, urand(g, *shape) * 2 - 1)


def ours_and_ref(typ, nm_or_dim, pts, tgt, data, eps, isign, n_trans=1, **opts):
    """-> (ours, ref run 1, ref run 2); nm x-fastest; pts = (x, y, z)."""
    from jax_finufft_b200.plan import Plan

    dim = len(pts)
    pad = [None] * (3 - dim)
    p = Plan(typ, nm_or_dim, n_trans=n_trans, eps=eps, isign=isign, **opts)
    p.setpts(*pts, *pad, *tgt, *([None] * (3 - len(tgt))))
    mine = p.execute(data).clone()
    p.destroy()
    outs = []
    for _ in range(2):
        r = ref.RefPlan(typ, nm_or_dim, n_trans=n_trans, eps=eps, isign=isign, **opts)
        r.setpts(*pts, *pad, *tgt, *([None] * (3 - len(tgt))))
        outs.append(r.execute(data).clone())
        r.destroy()
    torch.cuda.synchronize()
    return mine, outs[0], outs[1]


def check(name, eps, mine, r1, r2):
    e = relerr(mine, r1)
    rr = relerr(r2, r1)
    bound = max(2 * eps, 1.2 * rr)
    print(f"\nPARITY {name}: ours<->ref {e:.3e}  ref<->ref {rr:.3e}  2*eps {2 * eps:.1e}  bound {bound:.3e}")
    assert mine.shape == r1.shape
    assert e <= bound, (name, e, rr, bound)
    return e, rr


def test_c2_2d_type2():
    """BASELINE configs[1]: 2-D type 2, M=1e7, N=2048x2048, eps=1e-5, complex64."""
    g = torch.Generator(device="cuda").manual_seed(1)
    M, nm, eps = 10 ** 7, (2048, 2048), 1e-5
    pts = [(urand(g, M) * 2 - 1) * np.pi for _ in range(2)]
    f = cplx(g, 1, *nm)
    check("C2 2-D t2 M=1e7 N=2048^2 eps=1e-5", eps, *ours_and_ref(2, nm, pts, [], f, eps, -1))


def c3_points(g, M, dist):
    if dist == "uniform":
        return [(urand(g, M) * 2 - 1) * np.pi for _ in range(3)]
    h = 2 * np.pi / 512  # all points in an 8^3-cell corner box of the 512^3 fine grid
    return [-np.pi + urand(g, M) * (8 * h) for _ in range(3)]


@pytest.mark.parametrize("dist", ["uniform", "clustered"])
@pytest.mark.parametrize("typ", [1, 2])
def test_c3_3d(typ, dist):
    """BASELINE configs[2]: 3-D, M=1e8 uniform and clustered points, N=256^3, eps=1e-6, complex64."""
    g = torch.Generator(device="cuda").manual_seed(1 if dist == "uniform" else 2)
    M, nm, eps = 10 ** 8, (256, 256, 256), 1e-6
    pts = c3_points(g, M, dist)
    data = cplx(g, 1, M) if typ == 1 else cplx(g, 1, *nm)
    check(f"C3 3-D t{typ} {dist} M=1e8 N=256^3 eps=1e-6", eps,
          *ours_and_ref(typ, nm, pts, [], data, eps, 1 if typ == 1 else -1))


def test_c4_2d_type1_stacked():
    """BASELINE configs[3]: 64 stacked 2-D type-1 transforms sharing M=1e7 points, N=1024^2 (all 64)."""
    g = torch.Generator(device="cuda").manual_seed(3)
    M, nm, eps, ntr = 10 ** 7, (1024, 1024), 1e-6, 64
    pts = [(urand(g, M) * 2 - 1) * np.pi for _ in range(2)]
    c = cplx(g, ntr, M)
    mine, r1, r2 = ours_and_ref(1, nm, pts, [], c, eps, 1, n_trans=ntr)
    check("C4 2-D t1 x64 M=1e7 N=1024^2 eps=1e-6 (whole stack)", eps, mine, r1, r2)
    worst = max(relerr(mine[t], r1[t]) for t in range(ntr))
    rr = max(relerr(r2[t], r1[t]) for t in range(ntr))
    print(f"PARITY C4 worst single transform: ours<->ref {worst:.3e}  ref<->ref {rr:.3e}")
    assert worst <= max(2 * eps, 1.2 * rr)


def test_c5_3d_type3():
    """BASELINE configs[4]: 3-D type 3, 1e7 sources in [-pi, pi)^3 -> 1e7 targets in [-64, 64)^3, eps=1e-6.

    Type 3 evaluates phases up to S*X = 64*pi rad per dimension from float32 coordinates, so the
    rounding of the rescaled coordinates (~6e-8 * S*X*sqrt(3) ~ 2e-5), not eps, sets the distance
    between ANY two float pipelines.  The yardstick is then SURVEY.md 8c's third comparison: a
    float64 direct sum on a sample of targets -- ours must not be further from it than the reference is."""
    g = torch.Generator(device="cuda").manual_seed(4)
    M = N = 10 ** 7
    eps, S = 1e-6, 64.0
    pts = [(urand(g, M) * 2 - 1) * np.pi for _ in range(3)]
    tgt = [(urand(g, N) * 2 - 1) * S for _ in range(3)]
    c = cplx(g, 1, M)
    mine, r1, r2 = ours_and_ref(3, 3, pts, tgt, c, eps, 1, upsampfac=2.0)
    e, rr = relerr(mine, r1), relerr(r2, r1)
    # float64 direct sum at 48 sampled targets
    js = torch.as_tensor(np.random.default_rng(5).integers(0, N, size=48), device="cuda")
    x64 = [p.to(torch.float64) for p in pts]
    c128 = c[0].to(torch.complex128)
    truth = []
    for j in js.tolist():
        ph = sum(float(tgt[d][j]) * x64[d] for d in range(3))
        truth.append(torch.sum(c128 * torch.polar(torch.ones_like(ph), ph)))
    truth = torch.stack(truth)
    e_ours = relerr(mine[0][js], truth)
    e_ref = relerr(r1[0][js], truth)
    print(f"\nPARITY C5 3-D t3 M=N=1e7 S=64 eps=1e-6: ours<->ref {e:.3e}  ref<->ref {rr:.3e}  2*eps {2 * eps:.1e}; "
          f"vs float64 direct sum (48 targets): ours {e_ours:.3e}  ref {e_ref:.3e}")
    if e > max(2 * eps, 1.2 * rr):   # the fp32 coordinate-rounding regime: be no worse than the reference
        assert e_ours <= 1.2 * e_ref + eps, (e, rr, e_ours, e_ref)
        assert e <= e_ours + e_ref + 2 * eps
