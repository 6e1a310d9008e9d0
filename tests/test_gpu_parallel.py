"""Multi-GPU paths on real devices (NCCL).  Single-GPU cases run the same code with world = 1;
the world_size-2 cases need two visible GPUs (gpurun --gpus 2) and are skipped otherwise."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _data(M, nm, seed, dev):
    rng = np.random.default_rng(seed)
    pts = rng.uniform(-np.pi, np.pi, size=(3, M)).astype(np.float32)
    c = (rng.uniform(-1, 1, M) + 1j * rng.uniform(-1, 1, M)).astype(np.complex64)
    return [torch.as_tensor(p, device=dev) for p in pts], torch.as_tensor(c, device=dev), pts, c


@pytest.mark.parametrize("nm,iflag", [((24, 20, 16), 1), ((16, 18, 30), -1)])
def test_slab_split_single_gpu(nm, iflag):
    """combine="slab" with world = 1: the whole grid is one slab whose halo wraps onto itself."""
    import jax_finufft_b200 as J
    import oracle
    from jax_finufft_b200 import parallel as P

    dev = torch.device("cuda:0")
    tp, tc, pts, c = _data(30000, nm, 4, dev)
    tp[0][:4] = torch.tensor([-np.pi, 0.0, np.nextafter(np.float32(np.pi), np.float32(0)), 3 * np.pi], device=dev)
    a = P.nufft1_sharded_points(nm, tc, *tp, combine="slab", iflag=iflag, eps=1e-6)
    b = J.nufft1(nm, tc, *tp, iflag=iflag, eps=1e-6)
    assert a.shape == b.shape == tuple(nm)
    assert oracle.relerr(a.cpu().numpy(), b.cpu().numpy()) < 2e-6
    # double precision (tile kernels behind the spread-only plan), eps = 1e-10
    tp64, tc64 = [p.double() for p in tp], tc.to(torch.complex128)
    a = P.nufft1_sharded_points(nm, tc64, *tp64, combine="slab", iflag=iflag, eps=1e-10)
    b = J.nufft1(nm, tc64, *tp64, iflag=iflag, eps=1e-10)
    assert a.dtype == torch.complex128
    assert oracle.relerr(a.cpu().numpy(), b.cpu().numpy()) < 2e-10


def test_slab_partition_groups_points_by_owner():
    """b2n_slab_partition (csrc/slab.cu) against the torch restatement in exchange_points_by_slab:
    same owners, same counts, same re-based coordinates, every point exactly once."""
    import ctypes as C

    from jax_finufft_b200 import _lib
    from jax_finufft_b200 import parallel as P

    dev = torch.device("cuda:0")
    M, nf0, world, ns = 200_003, 96, 4, 7
    h = P.slab_halo(ns)
    g = torch.Generator(device=dev).manual_seed(2)
    pts = [(torch.rand(M, generator=g, device=dev) * 2 - 1) * 3 * np.pi for _ in range(3)]  # beyond one period
    pts[0][:5] = torch.tensor([-np.pi, np.pi, 0.0, np.pi / 2, -np.pi / 2], device=dev)       # on the slab edges
    c = torch.complex(torch.arange(M, device=dev, dtype=torch.float32), torch.rand(M, generator=g, device=dev))
    oz, oy, ox = (torch.empty(M, dtype=torch.float32, device=dev) for _ in range(3))
    oc = torch.empty(M, dtype=torch.complex64, device=dev)
    cnt = torch.empty(2 * world, dtype=torch.int64, device=dev)
    vp = C.c_void_p
    ier = _lib.lib().b2n_slab_partition(0, vp(torch.cuda.current_stream().cuda_stream), M, vp(pts[0].data_ptr()),
                                        vp(pts[1].data_ptr()), vp(pts[2].data_ptr()), vp(c.data_ptr()), nf0, world, h,
                                        vp(oz.data_ptr()), vp(oy.data_ptr()), vp(ox.data_ptr()), vp(oc.data_ptr()),
                                        vp(cnt.data_ptr()))
    assert ier == 0
    torch.cuda.synchronize()
    L, Lz = nf0 // world, nf0 // world + 2 * h
    t = pts[0].double() / (2 * np.pi) + 0.5
    zf = (t - t.floor()) * nf0
    owner = (zf / L).floor().long().clamp(0, world - 1)
    z_in = ((zf - owner * L + h) * (2 * np.pi / Lz) - np.pi).float()
    counts = cnt[:world].cpu()
    assert torch.equal(counts, torch.bincount(owner, minlength=world).cpu())
    orig = oc.real.round().long()                                          # Re c = original index
    assert torch.equal(torch.sort(orig).values, torch.arange(M, device=dev))
    edges = torch.cat([torch.zeros(1, dtype=torch.int64), counts.cumsum(0)])
    blk = torch.bucketize(torch.arange(M), edges[1:], right=True).to(dev)  # owner block of each output row
    assert torch.equal(blk, owner[orig])
    assert torch.allclose(oz, z_in[orig], rtol=0, atol=1e-6)
    assert torch.equal(oy, pts[1][orig]) and torch.equal(ox, pts[2][orig])
    assert torch.equal(oc.imag, c.imag[orig])


@pytest.mark.parametrize("nm,iflag", [((24, 20, 16), 1), ((16, 18, 30), -1)])
def test_native_type1_pipeline_single_gpu(nm, iflag):
    """spread-only plan + slab/pencil FFT + deconvolve == the fused single-GPU nufft1 (<= 2 eps)
    and the oracle."""
    import jax_finufft_b200 as J
    import oracle
    from jax_finufft_b200 import parallel as P

    dev = torch.device("cuda:0")
    tp, tc, pts, c = _data(30000, nm, 3, dev)
    a = P.nufft1_sharded_points(nm, tc, *tp, combine="reduce_scatter", iflag=iflag, eps=1e-6)
    b = J.nufft1(nm, tc, *tp, iflag=iflag, eps=1e-6)
    assert a.shape == b.shape == tuple(nm)
    assert oracle.relerr(a.cpu().numpy(), b.cpu().numpy()) < 2e-6
    want = oracle.nufft1(tuple(nm[::-1]), c, *pts[::-1].astype(np.float64), iflag=iflag, eps=1e-6, prec=1)
    assert oracle.relerr(a.cpu().numpy(), want) < 2e-5


@pytest.mark.parametrize("N,modeord,iflag,dt", [((12, 10, 16), 0, 1, torch.complex64), ((9, 15, 7), 1, -1, torch.complex64),
                                                 ((8, 8, 8), 0, -1, torch.complex128)])
def test_native_slab_pencil_fft_equals_the_torch_restatement(N, modeord, iflag, dt):
    """csrc/slab.cu (b2n_slab_fft_xy / b2n_slab_fft_z: cuFFT + fused crop/deconvolve kernels) against the
    torch implementation of the same stages that the gloo tests exercise (tests/test_parallel.py)."""
    from jax_finufft_b200 import parallel as P

    ns, beta, nf = P.fine_grid_geometry(N, 1e-6, dt == torch.complex64)
    g = torch.Generator().manual_seed(3)
    rdt = torch.float32 if dt == torch.complex64 else torch.float64
    slab = torch.complex(torch.rand(nf, generator=g, dtype=rdt) - 0.5, torch.rand(nf, generator=g, dtype=rdt) - 0.5)
    want = P.slab_pencil_fft(slab.clone(), N, nf, iflag, ns, beta, modeord=modeord)            # CPU tensors: torch glue
    got = P.slab_pencil_fft(slab.clone().cuda(), N, nf, iflag, ns, beta, modeord=modeord)       # CUDA: the library
    assert got.shape == want.shape == tuple(N)
    err = float(torch.linalg.vector_norm(got.cpu() - want) / torch.linalg.vector_norm(want))
    assert err < (2e-6 if dt == torch.complex64 else 1e-13), err


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import jax_finufft_b200 as J
        import oracle
        from jax_finufft_b200 import parallel as P

        nm, M = (32, 24, 20), 40000
        tp, tc, pts, c = _data(M, nm, 5, dev)
        lo, hi = P.shard_range(M, world, rank)
        loc = [p[lo:hi].contiguous() for p in tp]
        full = J.nufft1(nm, tc, *tp, eps=1e-6)
        res = {}
        a = P.nufft1_sharded_points(nm, tc[lo:hi].contiguous(), *loc, combine="reduce_scatter", eps=1e-6)
        res["rs"] = oracle.relerr(a.cpu().numpy(), full.cpu().numpy())
        for gather in (True, False):
            a = P.nufft1_sharded_points(nm, tc[lo:hi].contiguous(), *loc, combine="slab", eps=1e-6, gather=gather)
            if not gather:
                ylo, yhi = P.shard_range(nm[1], world, rank)
                res["slab_sharded"] = oracle.relerr(a.cpu().numpy(), full[:, ylo:yhi].cpu().numpy())
            else:
                res["slab"] = oracle.relerr(a.cpu().numpy(), full.cpu().numpy())
        b = P.nufft1_sharded_points(nm, tc[lo:hi].contiguous(), *loc, combine="psum", eps=1e-6)
        res["psum"] = oracle.relerr(b.cpu().numpy(), full.cpu().numpy())
        f = torch.as_tensor((np.random.default_rng(9).uniform(-1, 1, nm) + 0j).astype(np.complex64), device=dev)
        c2 = P.nufft2_sharded_points(f, *loc, eps=1e-6)
        c2_full = J.nufft2(f, *tp, eps=1e-6)
        res["t2"] = oracle.relerr(c2.cpu().numpy(), c2_full[lo:hi].cpu().numpy())
        stack = torch.stack([tc * (k + 1) for k in range(5)])
        s = P.nufft1_stacked(nm, stack, *tp, gather=True, eps=1e-6)
        s_full = J.nufft1(nm, stack, *[p[None] for p in tp], eps=1e-6)
        res["stack"] = oracle.relerr(s.cpu().numpy(), s_full.cpu().numpy())
        tg = [torch.as_tensor(np.random.default_rng(11 + d).uniform(-15, 15, 3000).astype(np.float32), device=dev) for d in range(3)]
        t3 = P.nufft3_sharded_sources(tc[lo:hi].contiguous(), *loc, *tg, eps=1e-6)
        t3_full = J.nufft3(tc, *tp, *tg, eps=1e-6)
        res["t3"] = oracle.relerr(t3.cpu().numpy(), t3_full.cpu().numpy())
        ret[rank] = res
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_world2_nccl_paths_match_single_gpu():
    world = 2
    ret = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    for r in range(world):
        for k, v in ret[r].items():
            assert v < (1e-5 if k == "t3" else 3e-6), (r, dict(ret[r]))
