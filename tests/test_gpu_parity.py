"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI of
libb200nufft.so (plan interface or b2n_run) and is compared with
  * the recorded outputs of the unmodified reference cuFINUFFT (tests/golden),
  * the CPU oracle (float64 restatement) on seeded mid-size inputs,
  * the reference library itself when oracle/_ref travelled to the box,
  * a float64 direct NUDFT,
with the contract tolerance: relative l2 <= 2*eps (+ fp32 rounding floor), written in
golden/cases.py::tolerance."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch

import oracle
from oracle import ref_cufinufft as ref

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import cases as G  # noqa: E402

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(HERE, "golden", "ref_cufinufft_golden.npz")


def T(a, dev="cuda"):
    return torch.as_tensor(np.ascontiguousarray(a), device=dev)


def run_plan(typ, dim, nm, pts, tgt, data, eps, iflag, dbl, **opts):
    from jax_finufft_b200.plan import Plan

    p = Plan(typ, dim if typ == 3 else nm, n_trans=data.shape[0], eps=eps, isign=iflag,
             dtype="complex128" if dbl else "complex64", **opts)
    p.setpts(*([T(x) for x in pts] + [None] * (3 - dim)), *([T(s) for s in tgt] + [None] * (3 - len(tgt))))
    out = p.execute(T(data)).cpu().numpy()
    info = p.info()
    p.destroy()
    return out, info


@pytest.mark.parametrize("method", [0, 1, 2, 3])
@pytest.mark.parametrize("case", G.CASES, ids=[c[0] for c in G.CASES])
def test_vs_reference_golden(case, method):
    """ours vs recorded reference cuFINUFFT outputs, every gpu_method value the FFI can pass
    (the reference rejects some method/type pairs; we accept 0-3 everywhere)."""
    name, typ, dim, nm, M, N, eps, dbl, iflag, ntr, sigma, modeord = case
    blob = np.load(GOLDEN)
    gold = blob[name]
    inp = G.make_inputs(case)
    out, info = run_plan(typ, dim, nm, inp["pts"], inp["tgt"], inp["data"], eps, iflag, dbl,
                         upsampfac=sigma, modeord=modeord, gpu_method=method)
    assert out.shape == gold.shape
    err = oracle.relerr(out, gold)
    print(f"\nPARITY golden {name} method={method}: {err:.3e} ({err / eps:.2f} eps)")
    assert err < G.tolerance(case), (name, method, err)


@pytest.mark.parametrize("case", G.CASES, ids=[c[0] for c in G.CASES])
def test_not_worse_than_reference_vs_nudft(case):
    """Yardstick of SURVEY.md §8c: our distance to the float64 NUDFT must not exceed the
    reference's own (recorded) distance by more than rounding."""
    name, typ, dim, nm, M, N, eps, dbl, iflag, ntr, sigma, modeord = case
    blob = np.load(GOLDEN)
    ref_err = float(blob[name + "__ref_vs_nudft"])
    inp = G.make_inputs(case)
    out, _ = run_plan(typ, dim, nm, inp["pts"], inp["tgt"], inp["data"], eps, iflag, dbl, upsampfac=sigma, modeord=modeord)
    pts = [p.astype(np.float64) for p in inp["pts"]]
    truth = []
    for t in range(ntr):
        d = inp["data"][t].astype(np.complex128)
        if typ == 1:
            truth.append(oracle.dirft1(nm, d, *pts, iflag=iflag, modeord=modeord))
        elif typ == 2:
            truth.append(oracle.dirft2(d, *pts, iflag=iflag, modeord=modeord))
        else:
            truth.append(oracle.dirft3(d, pts, [s.astype(np.float64) for s in inp["tgt"]], iflag=iflag))
    err = oracle.relerr(out, np.stack(truth))
    assert err < 1.5 * ref_err + (1e-14 if dbl else 1e-6), (name, err, ref_err)
    assert err < 20 * eps + (0 if dbl else 2e-6)   # the reference's own checktol/tol ratio


MID = [  # (dim, n_modes x-fastest, M, eps, double)
    (1, (5000,), 30000, 1e-6, False),
    (2, (200, 180), 100000, 1e-5, False),
    (2, (128, 96), 60000, 1e-6, False),
    (3, (48, 40, 36), 200000, 1e-6, False),
    (3, (30, 32, 34), 100000, 1e-4, False),
    (3, (24, 20, 28), 50000, 1e-3, False),
    (2, (100, 90), 50000, 1e-11, True),
    (3, (24, 20, 28), 50000, 1e-9, True),
]


@pytest.mark.parametrize("dim,nm,M,eps,dbl", MID, ids=[f"{d}d_{e:g}_{'d' if b else 'f'}" for d, _, _, e, b in MID])
def test_mid_size_vs_oracle_and_live_reference(dim, nm, M, eps, dbl):
    rng = np.random.default_rng(42 + dim)
    rd, cd = (np.float64, np.complex128) if dbl else (np.float32, np.complex64)
    pts = [rng.uniform(-np.pi, np.pi, M).astype(rd) for _ in range(dim)]
    c = (rng.uniform(-1, 1, (2, M)) + 1j * rng.uniform(-1, 1, (2, M))).astype(cd)
    fk = (rng.uniform(-1, 1, (2,) + nm[::-1]) + 1j * rng.uniform(-1, 1, (2,) + nm[::-1])).astype(cd)
    tol = G.parity_tol(eps, dbl)
    p64 = [p.astype(np.float64) for p in pts]
    prec = 0 if dbl else 1
    f, info = run_plan(1, dim, nm, pts, [], c, eps, 1, dbl)
    e1 = oracle.relerr(f, oracle.nufft1(nm, c, *p64, eps=eps, iflag=1, prec=prec))
    c2, _ = run_plan(2, dim, nm, pts, [], fk, eps, -1, dbl)
    e2 = oracle.relerr(c2, oracle.nufft2(fk, *p64, eps=eps, iflag=-1, prec=prec))
    print(f"\nPARITY mid {dim}d eps={eps:g} {'f64' if dbl else 'f32'} vs oracle: t1 {e1 / eps:.2f} eps  t2 {e2 / eps:.2f} eps")
    assert e1 < tol and e2 < tol, (e1, e2, tol)
    if ref.available():
        tp = [T(p) for p in pts] + [None] * (3 - dim)
        r = ref.RefPlan(1, nm, n_trans=2, eps=eps, isign=1, dtype="complex128" if dbl else "complex64").setpts(*tp)
        fr = r.execute(T(c)).cpu().numpy()
        r.destroy()
        print(f"PARITY mid {dim}d vs live reference: t1 {oracle.relerr(f, fr) / eps:.2f} eps")
        assert oracle.relerr(f, fr) < tol
        r = ref.RefPlan(2, nm, n_trans=2, eps=eps, isign=-1, dtype="complex128" if dbl else "complex64").setpts(*tp)
        cr = r.execute(T(fk)).cpu().numpy()
        r.destroy()
        assert oracle.relerr(c2, cr) < tol


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_type3_mid_vs_oracle(dim):
    rng = np.random.default_rng(7)
    M, N, eps = 40000, 30000, 1e-6
    pts = [rng.uniform(-np.pi, np.pi, M).astype(np.float32) for _ in range(dim)]
    tgt = [rng.uniform(-40, 40, N).astype(np.float32) + (3.0 if d == 0 else 0.0) for d in range(dim)]
    c = (rng.uniform(-1, 1, (2, M)) + 1j * rng.uniform(-1, 1, (2, M))).astype(np.complex64)
    f, info = run_plan(3, dim, (), pts, tgt, c, eps, -1, False, upsampfac=2.0)
    fo = oracle.nufft3(c, [p.astype(np.float64) for p in pts], [s.astype(np.float64) for s in tgt], eps=eps, iflag=-1, prec=1)
    err = oracle.relerr(f, fo)
    # float type 3: phases reach S*X ~ 43*pi rad per dim, so the fp32 rounding of the rescaled
    # coordinates costs ~2.5e-7 * S*X*sqrt(dim) whatever the kernel (ops_test.py:120 allows 1e-3)
    assert err < 2 * eps + 2.5e-7 * 43 * np.pi * np.sqrt(dim)
    if ref.available():   # ... and the reference library pays the same price on the same inputs
        r = ref.RefPlan(3, dim, n_trans=2, eps=eps, isign=-1, upsampfac=2.0)
        r.setpts(*([T(x) for x in pts] + [None] * (3 - dim)), *([T(s) for s in tgt] + [None] * (3 - dim)))
        fr = r.execute(T(c)).cpu().numpy()
        r.destroy()
        assert err < 1.5 * oracle.relerr(fr, fo) + 1e-6
        assert oracle.relerr(f, fr) < 2 * eps + 2.5e-7 * 43 * np.pi * np.sqrt(dim)


@pytest.mark.parametrize("dist", ["uniform", "clustered", "mixed"])
@pytest.mark.parametrize("method", [2, 0])
def test_binsort_contract_vs_oracle(method, dist):
    """histogram, exclusive-scan offsets and per-bin point SETS must match;
    the order inside a bin is unspecified in the reference (atomicAdd ranks).  gpu_method=2 (tile
    kernels) bins by floor(x') exactly like the reference; the default 3-D float path (sliding-
    window kernels) bins by the window's anchor cell and orders each bin by anchor z."""
    from jax_finufft_b200.plan import Plan

    rng = np.random.default_rng(11)
    M, nm = 300000, (40, 36, 30)
    pts = [rng.uniform(-np.pi, np.pi, M).astype(np.float32) for _ in range(3)]
    if dist != "uniform":
        # SURVEY.md §8(d) "clustered": iid uniform in a corner box 8 fine-grid cells wide -- every warp
        # of the sort sees repeated keys, which switches the histogram to its per-CTA hot-key table
        # and the placement pass to its aggregated variant (sort.cu); "mixed" = half and half
        n = M if dist == "clustered" else M // 2
        for d in range(3):
            h = 2 * np.pi / (2 * nm[d])
            pts[d][:n] = (-np.pi + rng.uniform(0, 8 * h, n)).astype(np.float32)
        perm = rng.permutation(M)
        pts = [x[perm] for x in pts]
    pts[0][:5] = [-np.pi, np.pi, np.nextafter(np.float32(np.pi), np.float32(0)), 0.0, 3 * np.pi]
    pts[1][:5] = pts[0][:5]
    pts[2][:5] = pts[0][:5]
    p = Plan(1, nm, eps=1e-6, gpu_method=method).setpts(*[T(x) for x in pts])
    info = p.info()
    idx, bstart = p.sort_arrays()
    idx, bstart = idx.cpu().numpy(), bstart.cpu().numpy()
    p.destroy()
    nf = [int(info.nf[d]) for d in range(3)]
    bins = [int(info.binsize[d]) for d in range(3)]
    p64 = [x.astype(np.float64) for x in pts]
    if method == 2:
        assert info.method == 2
        binid, hist = oracle.binsort(p64, nf, bins, prec=1)
    else:
        assert info.method == 3
        binid, hist, uz = oracle.binsort_anchor(p64, nf, bins, int(info.ns), prec=1)
    assert bstart[0] == 0 and bstart[-1] == M
    assert (np.diff(bstart) == hist).all()                       # histogram + offsets
    assert (np.sort(idx) == np.arange(M)).all()                  # a permutation
    assert (binid[idx] == np.repeat(np.arange(hist.size), hist)).all()   # per-bin sets
    if method == 0:   # anchor z is non-decreasing inside every bin
        d = np.diff(uz[idx])
        same_bin = np.diff(binid[idx]) == 0
        assert (d[same_bin] >= 0).all()


@pytest.mark.parametrize("dist", ["uniform", "mixed"])
def test_binsort_contract_partitioned_sort(dist):
    """Point sets large enough for the partition pass P1 (records > 48 MB, sort.cu): same contract."""
    from jax_finufft_b200.plan import Plan

    rng = np.random.default_rng(12)
    M, nm = 4_000_000, (64, 56, 60)
    pts = [rng.uniform(-np.pi, np.pi, M).astype(np.float32) for _ in range(3)]
    if dist == "mixed":  # a quarter of the points in a corner box: hot keys, below the aggregation threshold
        n = M // 4
        for d in range(3):
            h = 2 * np.pi / (2 * nm[d])
            pts[d][:n] = (-np.pi + rng.uniform(0, 8 * h, n)).astype(np.float32)
        perm = rng.permutation(M)
        pts = [x[perm] for x in pts]
    p = Plan(1, nm, eps=1e-6).setpts(*[T(x) for x in pts])
    info = p.info()
    assert info.method == 3
    idx, bstart = p.sort_arrays()
    idx, bstart = idx.cpu().numpy(), bstart.cpu().numpy()
    p.destroy()
    nf = [int(info.nf[d]) for d in range(3)]
    bins = [int(info.binsize[d]) for d in range(3)]
    binid, hist, uz = oracle.binsort_anchor([x.astype(np.float64) for x in pts], nf, bins, int(info.ns), prec=1)
    assert bstart[0] == 0 and bstart[-1] == M
    assert (np.diff(bstart) == hist).all()
    assert (np.sort(idx) == np.arange(M)).all()
    assert (binid[idx] == np.repeat(np.arange(hist.size), hist)).all()
    d = np.diff(uz[idx])
    assert (d[np.diff(binid[idx]) == 0] >= 0).all()


@pytest.mark.parametrize("dim", [2, 3])
def test_clustered_at_the_periodic_seam(dim):
    """All points in a box 8 fine cells wide that STRADDLES x = +pi (the seam of the periodic grid):
    the sort takes its clustered path (hot-key table + aggregated placement, also with the records
    folded afresh because M is small), and the register kernels' anchor cells wrap."""
    import jax_finufft_b200 as J

    rng = np.random.default_rng(5)
    nm = (40, 36, 30)[:dim]
    M = 60000
    pts = []
    for d in range(dim):
        h = 2 * np.pi / (2 * nm[d])
        x = np.pi + rng.uniform(-4 * h, 4 * h, M)          # outside [-pi, pi) on purpose: folded by the library
        pts.append(x.astype(np.float32))
    c = (rng.uniform(-1, 1, M) + 1j * rng.uniform(-1, 1, M)).astype(np.complex64)
    f = J.nufft1(nm[::-1], T(c), *[T(x) for x in pts[::-1]], eps=1e-6, iflag=1).cpu().numpy()
    fo = oracle.nufft1(nm, c, *[x.astype(np.float64) for x in pts], eps=1e-6, iflag=1, prec=1)
    assert oracle.relerr(f, fo) < 2e-5
    fk = (rng.uniform(-1, 1, nm[::-1]) + 1j * rng.uniform(-1, 1, nm[::-1])).astype(np.complex64)
    c2 = J.nufft2(T(fk), *[T(x) for x in pts[::-1]], eps=1e-6, iflag=-1).cpu().numpy()
    co = oracle.nufft2(fk, *[x.astype(np.float64) for x in pts], eps=1e-6, iflag=-1, prec=1)
    assert oracle.relerr(c2, co) < 2e-5


def test_edge_cases_empty_single_far_and_clustered():
    from jax_finufft_b200.plan import Plan

    nm = (16, 12, 10)
    # M = 0: type 1 gives zeros, type 2 gives an empty vector
    p = Plan(1, nm, eps=1e-6).setpts(*[torch.empty(0, device="cuda") for _ in range(3)])
    f = p.execute(torch.empty((1, 0), dtype=torch.complex64, device="cuda"))
    assert f.shape == (1, 10, 12, 16) and float(f.abs().max()) == 0.0
    p.destroy()
    # M = 1 at the periodic seam, points far outside [-3pi, 3pi], all points in one fine cell
    rng = np.random.default_rng(0)
    for pts in ([np.array([np.pi], np.float32)] * 3,
                [rng.uniform(-300, 300, 2000).astype(np.float32) for _ in range(3)],
                [(1e-3 * rng.uniform(-1, 1, 5000) - np.pi).astype(np.float32) for _ in range(3)]):
        M = pts[0].size
        c = (rng.uniform(-1, 1, (1, M)) + 1j * rng.uniform(-1, 1, (1, M))).astype(np.complex64)
        f, _ = run_plan(1, 3, nm, pts, [], c, 1e-6, 1, False)
        fo = oracle.nufft1(nm, c, *[x.astype(np.float64) for x in pts], eps=1e-6, prec=1)
        tol = G.parity_tol(1e-6, False) if np.abs(pts[0]).max() < 10 else 1e-4   # |x|~300 in fp32: the fold itself loses 5 digits
        assert oracle.relerr(f, fo) < tol
        fk = (rng.uniform(-1, 1, (1,) + nm[::-1]) + 1j * rng.uniform(-1, 1, (1,) + nm[::-1])).astype(np.complex64)
        c2, _ = run_plan(2, 3, nm, pts, [], fk, 1e-6, -1, False)
        co = oracle.nufft2(fk, *[x.astype(np.float64) for x in pts], eps=1e-6, prec=1)
        assert oracle.relerr(c2, co) < tol


def test_many_transforms_more_than_one_batch_and_repeated_setpts():
    """ntransf > batch (impl.h:123-127) and two setpts with different M on one plan
    (V/test/cuda/cufinufft2d1nupts_test.cu -- the n_tot>1 pattern of run_nufft)."""
    from jax_finufft_b200.plan import Plan

    rng = np.random.default_rng(9)
    nm, ntr = (40, 30), 11
    p = Plan(1, nm, n_trans=ntr, eps=1e-5, gpu_maxbatchsize=4)
    for M in (7000, 2500):
        pts = [rng.uniform(-np.pi, np.pi, M).astype(np.float32) for _ in range(2)]
        c = (rng.uniform(-1, 1, (ntr, M)) + 1j * rng.uniform(-1, 1, (ntr, M))).astype(np.complex64)
        p.setpts(T(pts[0]), T(pts[1]))
        f = p.execute(T(c)).cpu().numpy()
        fo = oracle.nufft1(nm, c, *[x.astype(np.float64) for x in pts], eps=1e-5, prec=1)
        assert oracle.relerr(f, fo) < G.parity_tol(1e-5, False)
    p.destroy()


@pytest.mark.parametrize("kem,sort", [(0, 1), (1, 0), (0, 0)])
def test_kerevalmeth_direct_and_unsorted(kem, sort):
    rng = np.random.default_rng(13)
    nm, M = (20, 24, 18), 20000
    pts = [rng.uniform(-np.pi, np.pi, M).astype(np.float32) for _ in range(3)]
    c = (rng.uniform(-1, 1, (1, M)) + 1j * rng.uniform(-1, 1, (1, M))).astype(np.complex64)
    f, _ = run_plan(1, 3, nm, pts, [], c, 1e-5, 1, False, gpu_kerevalmeth=kem, gpu_sort=sort)
    fo = oracle.nufft1(nm, c, *[x.astype(np.float64) for x in pts], eps=1e-5, kerevalmeth=kem, prec=1)
    assert oracle.relerr(f, fo) < G.parity_tol(1e-5, False)


def test_spreadinterponly_matches_oracle_spread():
    """gpu_spreadinterponly=1 (impl.h:115-117): the output IS the fine grid."""
    rng = np.random.default_rng(17)
    nf, M = (64, 48, 40), 50000
    pts = [rng.uniform(-np.pi, np.pi, M).astype(np.float32) for _ in range(3)]
    c = (rng.uniform(-1, 1, (1, M)) + 1j * rng.uniform(-1, 1, (1, M))).astype(np.complex64)
    fw, info = run_plan(1, 3, nf, pts, [], c, 1e-6, 1, False, gpu_spreadinterponly=1)
    fo = oracle.spread([x.astype(np.float64) for x in pts], c[0], nf, info.ns, info.beta, prec=1)
    assert oracle.relerr(fw[0], fo) < 1e-6
    g = (rng.uniform(-1, 1, (1,) + nf[::-1]) + 1j * rng.uniform(-1, 1, (1,) + nf[::-1])).astype(np.complex64)
    ci, _ = run_plan(2, 3, nf, pts, [], g, 1e-6, -1, False, gpu_spreadinterponly=1)
    co = oracle.interp([x.astype(np.float64) for x in pts], g[0], info.ns, info.beta, prec=1)
    assert oracle.relerr(ci[0], co) < 1e-6


@pytest.mark.parametrize("dim,nf,ntr", [(2, (33, 20), 1), (2, (33, 21), 3), (3, (21, 16, 12), 1), (3, (15, 9, 7), 2),
                                        (3, (34, 33, 33), 2)])
@pytest.mark.parametrize("method", [0, 2])
def test_spreadinterponly_odd_grids(dim, nf, ntr, method):
    """The caller's grid with gpu_spreadinterponly has the MODE sizes: odd rows, odd sizes per
    transform.  The float tile kernels use 16-byte cell-pair accesses, so such grids must take the
    scalar kernels (misaligned-address faults / spills into the next row otherwise)."""
    rng = np.random.default_rng(19)
    M = 20000
    pts = [rng.uniform(-np.pi, np.pi, M).astype(np.float32) for _ in range(dim)]
    c = (rng.uniform(-1, 1, (ntr, M)) + 1j * rng.uniform(-1, 1, (ntr, M))).astype(np.complex64)
    fw, info = run_plan(1, dim, nf, pts, [], c, 1e-5, 1, False, gpu_spreadinterponly=1, gpu_method=method)
    p64 = [x.astype(np.float64) for x in pts]
    for t in range(ntr):
        fo = oracle.spread(p64, c[t], nf, info.ns, info.beta, prec=1)
        assert oracle.relerr(fw[t], fo) < 3e-6, (dim, nf, t)   # up to 20 points per cell: fp32 summation order
    g = (rng.uniform(-1, 1, (ntr,) + nf[::-1]) + 1j * rng.uniform(-1, 1, (ntr,) + nf[::-1])).astype(np.complex64)
    ci, _ = run_plan(2, dim, nf, pts, [], g, 1e-5, -1, False, gpu_spreadinterponly=1, gpu_method=method)
    for t in range(ntr):
        co = oracle.interp(p64, g[t], info.ns, info.beta, prec=1)
        assert oracle.relerr(ci[t], co) < 3e-6, (dim, nf, t)


def test_run_host_entry_point():
    """b2n_run_host: the C-ABI call a foreign-language host makes with HOST buffers."""
    from jax_finufft_b200 import _lib

    L = _lib.lib()
    rng = np.random.default_rng(21)
    n_tot, ntr, M, nk = 2, 3, 4000, (18, 14, 10)
    pts = [np.ascontiguousarray(rng.uniform(-np.pi, np.pi, (n_tot, M)).astype(np.float32)) for _ in range(3)]
    c = np.ascontiguousarray((rng.uniform(-1, 1, (n_tot, ntr, M)) + 1j * rng.uniform(-1, 1, (n_tot, ntr, M))).astype(np.complex64))
    out = np.zeros((n_tot, ntr) + nk[::-1], np.complex64)
    o = _lib.default_opts()
    o.upsampfac = 2.0
    n_k = (C.c_int64 * 3)(*nk)
    pp = (C.c_void_p * 3)(*[p.ctypes.data for p in pts])
    rc = L.b2n_run_host(1, 3, 0, 1e-6, 1, n_tot, ntr, M, n_k, C.byref(o), c.ctypes.data, pp, None, out.ctypes.data)
    assert rc == 0
    for i in range(n_tot):
        fo = oracle.nufft1(nk, c[i], *[p[i].astype(np.float64) for p in pts], eps=1e-6, prec=1)
        assert oracle.relerr(out[i], fo) < G.parity_tol(1e-6, False)


@pytest.mark.parametrize("typ", [1, 2])
def test_run_host_chunked_pipeline_matches_device_path(typ):
    """b2n_run_host splits large single transforms into point chunks (copy of chunk k+1 overlaps the
    bin-sort + spread/interp of chunk k, uniform-grid stages once).  The result must equal the
    device-resident b2n_run on the whole point set up to summation order (type 1) / exactly per
    point (type 2), at a size where 8 chunks are used (M >= 2^23)."""
    import jax_finufft_b200 as J
    from jax_finufft_b200 import _lib

    L = _lib.lib()
    M, nk = (1 << 23) + 12345, (48, 40, 36)      # odd tail: the last chunk is ragged
    g = torch.Generator(device="cuda").manual_seed(77)
    pts = [((torch.rand(M, device="cuda", generator=g) * 2 - 1) * np.pi) for _ in range(3)]
    o = _lib.default_opts()
    o.upsampfac = 2.0
    n_k = (C.c_int64 * 3)(*nk)
    hp = [p.cpu().pin_memory() for p in pts]
    pp = (C.c_void_p * 3)(*[p.data_ptr() for p in hp])
    if typ == 1:
        c = torch.complex(torch.rand(M, device="cuda", generator=g) * 2 - 1, torch.rand(M, device="cuda", generator=g) * 2 - 1)
        want = J.nufft1(nk[::-1], c, *pts[::-1], eps=1e-6, iflag=1).cpu().numpy()
        hc = c.cpu().pin_memory()
        out = torch.zeros(nk[::-1], dtype=torch.complex64).pin_memory()
        rc = L.b2n_run_host(1, 3, 0, 1e-6, 1, 1, 1, M, n_k, C.byref(o), C.c_void_p(hc.data_ptr()), pp, None,
                            C.c_void_p(out.data_ptr()))
        assert rc == 0
        assert oracle.relerr(out.numpy(), want) < 2e-6
    else:
        f = torch.complex(torch.rand(nk[::-1], device="cuda", generator=g) * 2 - 1,
                          torch.rand(nk[::-1], device="cuda", generator=g) * 2 - 1)
        want = J.nufft2(f, *pts[::-1], eps=1e-6, iflag=-1).cpu().numpy()
        hf = f.cpu().pin_memory()
        out = torch.zeros(M, dtype=torch.complex64).pin_memory()
        rc = L.b2n_run_host(2, 3, 0, 1e-6, -1, 1, 1, M, n_k, C.byref(o), C.c_void_p(hf.data_ptr()), pp, None,
                            C.c_void_p(out.data_ptr()))
        assert rc == 0
        assert oracle.relerr(out.numpy(), want) < 1e-6


@pytest.mark.parametrize("typ", [1, 2])
def test_full_size_properties_c3(typ):
    """BASELINE config C3 (3-D, M=1e8, N=256^3, eps=1e-6, c64): size-independent checks --
    a random sample of outputs against a float64 direct NUDFT, and the adjoint identity
    <nufft1(c), f> = <c, nufft2(f)> (type 1 / type 2 with the same points and opposite iflag
    are exact adjoints), which exercises every point and every mode."""
    from jax_finufft_b200.plan import Plan

    M, nm, eps = 10 ** 8, (256, 256, 256), 1e-6
    g = torch.Generator(device="cuda").manual_seed(1)
    pts = [(torch.rand(M, device="cuda", generator=g) * 2 - 1) * np.pi for _ in range(3)]
    c = torch.complex(torch.rand(M, device="cuda", generator=g) * 2 - 1, torch.rand(M, device="cuda", generator=g) * 2 - 1)[None]
    f = torch.complex(torch.rand(nm, device="cuda", generator=g) * 2 - 1, torch.rand(nm, device="cuda", generator=g) * 2 - 1)[None]
    p1 = Plan(1, nm, eps=eps, isign=1).setpts(*pts)
    F = p1.execute(c)
    p1.destroy()
    p2 = Plan(2, nm, eps=eps, isign=-1).setpts(*pts)
    Cc = p2.execute(f)
    p2.destroy()
    lhs = torch.sum(F.to(torch.complex128) * f.to(torch.complex128).conj())
    rhs = torch.sum(c.to(torch.complex128) * Cc.to(torch.complex128).conj())
    assert abs(lhs - rhs) / abs(lhs) < 2e-5
    rng = np.random.default_rng(3)
    if typ == 1:   # 24 random modes, each an M-term sum in float64
        ks = rng.integers(-128, 128, size=(24, 3))
        x64 = [p.to(torch.float64) for p in pts]
        c64 = c[0].to(torch.complex128)
        errs = []
        for kx, ky, kz in ks:
            ph = kx * x64[0] + ky * x64[1] + kz * x64[2]
            truth = torch.sum(c64 * torch.polar(torch.ones_like(ph), ph))
            got = F[0, kz + 128, ky + 128, kx + 128]
            errs.append(abs(complex(got) - complex(truth)))
        scale = float(torch.sqrt(torch.mean(torch.abs(F) ** 2)))
        assert max(errs) / scale < 2e-5
    else:          # 4000 random points, each a 256^3-term sum (separable in float64)
        js = torch.as_tensor(rng.integers(0, M, size=4000), device="cuda")
        k = torch.arange(-128, 128, device="cuda", dtype=torch.float64)
        ex = [torch.polar(torch.ones(4000, 256, device="cuda", dtype=torch.float64), -pts[d][js].to(torch.float64)[:, None] * k[None]) for d in range(3)]
        f128 = f[0].to(torch.complex128)             # [z, y, x]
        t = torch.einsum("zyx,jx->jzy", f128, ex[0])
        t = torch.einsum("jzy,jy->jz", t, ex[1])
        truth = torch.einsum("jz,jz->j", t, ex[2])
        got = Cc[0][js].to(torch.complex128)
        assert float(torch.linalg.norm(got - truth) / torch.linalg.norm(truth)) < 2e-5


@pytest.mark.parametrize("frac", [0.0, 0.5, 1.0], ids=["uniform", "half_clustered", "clustered"])
def test_two_pass_sort_and_its_overflow_fallback(frac):
    """Point sets large enough for the bucketed sort (> 48 MB of records).  Uniform input takes the
    two-pass sort (fixed-capacity bucket regions, histogram taken while partitioning); a bucket that
    outgrows its region flips a device flag MID-KERNEL and the three-pass pipeline takes over.  (The
    first version let every thread read that flag on its own: threads of one CTA parted ways in front
    of a barrier and wrote out of bounds -- only at full speed, never under the sanitizer.)"""
    from jax_finufft_b200.plan import Plan

    M, nm = 3_600_000, (64, 64, 64)
    rng = np.random.default_rng(31)
    k = int(M * frac)
    pts = []
    for d in range(3):
        u = rng.uniform(-np.pi, np.pi, M).astype(np.float32)
        u[:k] = (-np.pi + rng.uniform(0, 8 * 2 * np.pi / 128, k)).astype(np.float32)
        pts.append(u)
    perm = rng.permutation(M)
    pts = [u[perm] for u in pts]
    c = (rng.uniform(-1, 1, (1, M)) + 1j * rng.uniform(-1, 1, (1, M))).astype(np.complex64)
    tp = [T(u) for u in pts]
    for rep in range(3):   # the race needed several tries to show
        p = Plan(1, nm, eps=1e-5, isign=1)
        p.setpts(*tp)
        f = p.execute(T(c)).cpu().numpy()
        idx, bs = p.sort_arrays()
        p.destroy()
        assert int(bs[-1]) == M
        assert torch.equal(torch.sort(idx.long())[0], torch.arange(M, device="cuda"))
    fo = oracle.nufft1(nm, c, *[u.astype(np.float64) for u in pts], eps=1e-5, prec=1)
    assert oracle.relerr(f, fo) < G.parity_tol(1e-5, False)
