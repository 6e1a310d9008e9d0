"""Multi-GPU host logic (jax_finufft_b200/parallel.py) on CPU: world_size-2 `gloo` processes.

The spread itself has no CPU path; here each rank's private fine grid comes from the ORACLE's
spreader (test infrastructure), and what is under test is the decomposition that follows it:
reduce-scatter over z-slabs -> 2-D FFT + crop -> transpose -> 1-D FFT + crop + deconvolve, which
must reproduce the oracle's single-process type-1 transform of ALL points.  The `-m gpu` twin
(test_gpu_parallel.py) runs the same functions over NCCL with the CUDA spreader.
"""
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


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _inputs(M, seed):
    rng = np.random.default_rng(seed)
    pts = rng.uniform(-np.pi, np.pi, size=(3, M))
    c = rng.uniform(-1, 1, M) + 1j * rng.uniform(-1, 1, M)
    return pts, c


def _worker(rank, world, port, nm, M, iflag, modeord, gather, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from jax_finufft_b200 import parallel as P

        eps = 1e-6
        pts, c = _inputs(M, 7)
        lo, hi = P.shard_range(M, world, rank)
        ns, beta, nf = P.fine_grid_geometry(nm, eps, single=False)
        # nm is slowest-first (JAX order); the oracle takes the backend's x-fastest order
        grid = oracle.spread(list(pts[::-1, lo:hi]), c[lo:hi], list(nf[::-1]), ns, beta)
        slab = P.reduce_scatter_slabs(torch.from_numpy(grid), None)
        assert slab.shape == (nf[0] // world, nf[1], nf[2])
        out = P.slab_pencil_fft(slab, nm, nf, iflag, ns, beta, modeord=modeord, gather=gather)
        if not gather:
            ylo, yhi = P.shard_range(nm[1], world, rank)
            assert out.shape == (nm[0], yhi - ylo, nm[2])
            full = P._all_gather_cat(out, None, 1, [b - a for a, b in (P.shard_range(nm[1], world, r) for r in range(world))])
        else:
            full = out
        want = oracle.nufft1(tuple(nm[::-1]), c, *pts[::-1], iflag=iflag, eps=eps, modeord=modeord)
        err = oracle.relerr(full.numpy(), want)
        ret[rank] = err
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nm,iflag,modeord,gather", [
    ((12, 10, 14), 1, 0, True),
    ((9, 11, 8), -1, 1, False),     # odd sizes: uneven y-pencils; FFT-order modes
])
def test_reduce_scatter_path_matches_single_process_oracle(nm, iflag, modeord, gather):
    world = 2
    ret = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), nm, 500, iflag, modeord, gather, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        assert ret[r] < 1e-12, dict(ret)   # same arithmetic as the oracle, only the FFT library differs


def _worker_slab(rank, world, port, nm, M, iflag, ret):
    """Spatial split: exchange by z-slab -> (oracle) spread into slab + halo -> halo add -> slab/pencil FFT."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from jax_finufft_b200 import parallel as P

        eps = 1e-6
        pts, c = _inputs(M, 8)
        # points on the slab boundaries, at the periodic seam and just beside them
        pts[0, :6] = [-np.pi, 0.0, np.nextafter(0.0, -1), np.nextafter(np.pi, 0), 1e-9, -np.pi + 1e-9]
        lo, hi = P.shard_range(M, world, rank)
        ns, beta, nf = P.fine_grid_geometry(nm, eps, single=False)
        h = P.slab_halo(ns)
        src, loc, Lz = P.exchange_points_by_slab(torch.from_numpy(c[lo:hi]), [torch.from_numpy(p[lo:hi]) for p in pts],
                                                 nf[0], ns)
        assert Lz == nf[0] // world + 2 * h
        n_all = torch.tensor([src.numel()])
        dist.all_reduce(n_all)
        assert int(n_all) == M                                  # nobody lost or duplicated
        L = nf[0] // world
        cell = (loc[0].numpy() + np.pi) / (2 * np.pi) * Lz      # local plane coordinate
        assert cell.min() >= h - 1e-9 and cell.max() < L + h + 1e-9
        lnf = [Lz, nf[1], nf[2]]
        grid = oracle.spread([q.numpy() for q in loc[::-1]], src.numpy(), lnf[::-1], ns, beta)
        local = torch.from_numpy(grid)
        assert local[:h - 4].abs().max() == 0 and local[Lz - h + 4:].abs().max() == 0  # the safety margin stays empty
        slab = P.halo_add(local, h)
        assert slab.shape == (L, nf[1], nf[2])
        out = P.slab_pencil_fft(slab.contiguous(), nm, nf, iflag, ns, beta, gather=True)
        want = oracle.nufft1(tuple(nm[::-1]), c, *pts[::-1], iflag=iflag, eps=eps)
        ret[rank] = oracle.relerr(out.numpy(), want)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nm,iflag", [((16, 10, 12), 1), ((24, 9, 8), -1)])
def test_slab_split_matches_single_process_oracle(nm, iflag):
    world = 2
    ret = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker_slab, args=(world, _free_port(), nm, 700, iflag, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        assert ret[r] < 1e-11, dict(ret)


def _worker_split(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from jax_finufft_b200 import parallel as P

        src = torch.arange(7 * 5, dtype=torch.float32).reshape(7, 5)
        loc = P.split_transforms(src)
        sizes = [b - a for a, b in (P.shard_range(7, world, r) for r in range(world))]
        back = P._all_gather_cat(loc * 1.0, None, 0, sizes)
        ret[rank] = (tuple(loc.shape), bool(torch.equal(back, src)))
    finally:
        dist.destroy_process_group()


def test_split_transforms_and_gather_roundtrip():
    world = 2
    ret = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker_split, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert ret[0] == ((4, 5), True) and ret[1] == ((3, 5), True)


def test_shard_range_partitions():
    from jax_finufft_b200.parallel import shard_range

    for n in (0, 1, 7, 64, 100000001):
        for w in (1, 2, 3, 4, 8):
            edges = [shard_range(n, w, r) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[r][1] == edges[r + 1][0] for r in range(w - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


def test_product_path_needs_cuda():
    """No CPU fallback: the sharded entry points refuse CPU tensors like the single-GPU ones."""
    from jax_finufft_b200 import parallel as P

    x = torch.zeros(10, dtype=torch.float32)
    c = torch.zeros(10, dtype=torch.complex64)
    with pytest.raises((ValueError, RuntimeError)):
        P.nufft1_sharded_points((8, 8, 8), c, x, x, x, combine="psum")
