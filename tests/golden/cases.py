"""Seeded small cases shared by the golden-vector generator and the parity tests.

The reference stores no golden vectors: its tests compare against an on-the-fly
NUDFT with fixed seeds.  These cases follow that matrix -- V/test/cuda/CMakeLists.txt:26-236
(1-D / 2-D / 3-D, types 1-3, float tol 1e-4, double tol 1e-12, sigma=1.25 at 1e-8, many-vector
ntransf) and tests/ops_test.py:25-126 (M=50, N~75/ndim+5k, eps 1e-7 / 1e-10, iflag +-1) -- with
inputs regenerated from the seed, and the OUTPUTS of the unmodified reference cuFINUFFT (run on
a B200 by make_golden.py) stored in ref_cufinufft_golden.npz.
Backend conventions: n_modes = (ms, mt, mu) with x fastest; pts = (x, y, z).
"""
import numpy as np

# name, type, dim, n_modes (x fastest), M, N_targets, eps, double, iflag, ntransf, upsampfac, modeord
CASES = [
    ("t1_1d_f_1e-4", 1, 1, (200,), 400, 0, 1e-4, False, 1, 1, 2.0, 0),
    ("t2_1d_f_1e-4", 2, 1, (200,), 400, 0, 1e-4, False, -1, 1, 2.0, 0),
    ("t3_1d_f_1e-4", 3, 1, (), 400, 300, 1e-4, False, 1, 1, 2.0, 0),
    ("t1_2d_f_1e-4", 1, 2, (20, 30), 500, 0, 1e-4, False, 1, 1, 2.0, 0),
    ("t2_2d_f_1e-4", 2, 2, (20, 30), 500, 0, 1e-4, False, -1, 1, 2.0, 0),
    ("t3_2d_f_1e-4", 3, 2, (), 500, 300, 1e-4, False, 1, 1, 2.0, 0),
    ("t1_3d_f_1e-4_ref", 1, 3, (2, 5, 10), 20, 0, 1e-4, False, 1, 1, 2.0, 0),   # cufinufft3d_test sizes
    ("t2_3d_f_1e-4_ref", 2, 3, (2, 5, 10), 20, 0, 1e-4, False, -1, 1, 2.0, 0),
    ("t3_3d_f_1e-4_ref", 3, 3, (), 20, 20, 1e-4, False, 1, 1, 2.0, 0),
    ("t1_3d_f_1e-6", 1, 3, (12, 10, 8), 800, 0, 1e-6, False, 1, 1, 2.0, 0),
    ("t2_3d_f_1e-6", 2, 3, (12, 10, 8), 800, 0, 1e-6, False, -1, 1, 2.0, 0),
    ("t3_3d_f_1e-6", 3, 3, (), 800, 500, 1e-6, False, -1, 1, 2.0, 0),
    ("t1_2d_f_many", 1, 2, (24, 18), 600, 0, 1e-5, False, 1, 5, 2.0, 0),       # ntransf=5 as the CTest matrix
    ("t2_2d_f_many", 2, 2, (24, 18), 600, 0, 1e-5, False, -1, 5, 2.0, 0),
    ("t1_2d_f_modeord1_odd", 1, 2, (21, 16), 400, 0, 1e-5, False, -1, 1, 2.0, 1),
    ("t2_3d_f_modeord1_odd", 2, 3, (9, 8, 7), 400, 0, 1e-5, False, 1, 1, 2.0, 1),
    ("t1_1d_d_1e-12", 1, 1, (200,), 400, 0, 1e-12, True, 1, 1, 2.0, 0),
    ("t2_2d_d_1e-12", 2, 2, (20, 30), 500, 0, 1e-12, True, -1, 1, 2.0, 0),
    ("t1_3d_d_1e-12", 1, 3, (12, 10, 8), 800, 0, 1e-12, True, 1, 1, 2.0, 0),
    ("t2_3d_d_1e-10", 2, 3, (12, 10, 8), 800, 0, 1e-10, True, -1, 1, 2.0, 0),
    ("t3_3d_d_1e-10", 3, 3, (), 800, 500, 1e-10, True, 1, 1, 2.0, 0),
    ("t3_2d_d_1e-12", 3, 2, (), 500, 300, 1e-12, True, -1, 2, 2.0, 0),
    ("t1_2d_d_1e-8_s125", 1, 2, (20, 30), 500, 0, 1e-8, True, 1, 1, 1.25, 0),
    ("t2_3d_d_1e-8_s125", 2, 3, (12, 10, 8), 800, 0, 1e-8, True, -1, 1, 1.25, 0),
]


def make_inputs(case):
    """-> dict(pts=[x,y,z][:dim], tgt=[s,t,u][:dim] or [], data=(ntransf, ...) complex)."""
    name, typ, dim, nm, M, N, eps, dbl, iflag, ntr, sigma, modeord = case
    seed = 1000 + [c[0] for c in CASES].index(name)
    rng = np.random.default_rng(seed)
    rd = np.float64 if dbl else np.float32
    cd = np.complex128 if dbl else np.complex64
    pts = [rng.uniform(-np.pi, np.pi, M).astype(rd) for _ in range(dim)]
    tgt = [rng.uniform(-25.0, 25.0, N).astype(rd) for _ in range(dim)] if typ == 3 else []
    if typ == 2:
        shape = (ntr,) + tuple(nm[::-1])
    else:
        shape = (ntr, M)
    data = (rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)).astype(cd)
    return dict(pts=pts, tgt=tgt, data=data)


FLOOR = {False: 1e-6, True: 2e-14}   # rounding floor of the arithmetic (fp32 / fp64), NOT added to 2*eps


def parity_tol(eps, dbl):
    """The parity contract (BASELINE.json north_star): relative l2 vs the reference <= 2*eps.
    Below the resolution of the arithmetic the bound cannot follow eps any further: two fp32
    pipelines that differ only in summation order and cuFFT plan disagree at a few 1e-7
    (SURVEY.md 8c parity protocol 3), so the bound is max(2*eps, floor) -- at the headline
    eps = 1e-6 in complex64 that is exactly 2*eps."""
    return max(2.0 * eps, FLOOR[bool(dbl)])


def tolerance(case):
    typ, dim, eps, dbl = case[1], case[2], case[6], case[7]
    tol = parity_tol(eps, dbl)
    if typ == 3:
        # type 3 evaluates phases up to S*X = 25*pi rad per dim; the rounding of the rescaled
        # coordinates alone costs ~machine-eps * S*X*sqrt(dim) (the reference's own float type-3
        # error vs NUDFT on these cases is 1.4e-5 at eps=1e-6: golden key <name>__ref_vs_nudft)
        tol += (1e-15 if dbl else 2.5e-7) * 25 * np.pi * np.sqrt(dim)
    return tol
