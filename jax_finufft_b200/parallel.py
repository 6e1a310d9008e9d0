"""Multi-GPU execution on one 8xB200 box: one process per GPU, ``torch.distributed`` (NCCL over
NVLink 5 / NVSwitch).  SURVEY.md §8(e); DESIGN.md §6.

The reference has no multi-GPU code of its own: it relies on JAX sharding (``shard_map`` over a
device mesh, ref ``tests/sharding_test.py:119-334``, README.md:356-380).  The rules there are

* leading batch dimensions / stacked transforms shard trivially            (README.md:364-367)
* type 2 with the POINTS sharded needs no communication                    (sharding_test.py:197-241)
* type 1 with the POINTS sharded needs a ``psum`` of the output modes      (sharding_test.py:119-194)
* type 3: shard the sources (+ ``psum``) or the targets (nothing); both = unsupported
                                                                           (sharding_test.py:244-334)

Every function below is the body a ``shard_map`` would run on one device: it takes this rank's
LOCAL shard, calls the single-GPU backend, and performs the one collective the rule needs.
Work only shards where it splits naturally; the uniform grid itself is never distributed on the
caller's side ("replicas only").

What is new relative to the reference is the *native* type-1 path (``combine="reduce_scatter"``):
instead of every GPU running the full FFT and all-reducing the output modes, each GPU spreads its
points into a private fine grid, the fine grids are summed with ONE ``reduce_scatter`` along z
(8*nf bytes, each GPU receives nf3/G summed z-slabs), and the FFT is done slab/pencil-wise so
that its cost is divided by G instead of repeated G times:

    spread (local points, full private fine grid)                [libb200nufft, spread-only plan]
    reduce_scatter over z-slabs                                  [NCCL]
    2-D FFT (y, x) of the local slabs, crop to the N1 x N2 central modes (4x less data)
    all_to_all transpose: z-slabs -> y-pencils                   [NCCL]
    1-D FFT along z, crop to N3, divide by the kernel Fourier series (deconvolve)

``combine="slab"`` goes one step further (the *spatial* split): the points are first exchanged so
that every GPU holds the points of its own nf3/G z-slabs (ONE all_to_all of 20 bytes per point),
each GPU spreads into a grid of just its slabs plus a halo of ceil(ns/2)+2 planes on either side,
and only the halo planes (2 x 6 planes, 25 MB at C3 instead of the 1.07 GB private grid) travel to
the two neighbours.  Sort, spread and grid memory are all divided by G; the slab/pencil FFT is
the same.

The output is sharded over the y mode axis (``gather=True`` all-gathers it).  The psum path is
kept as the parity check.  All collectives are enqueued on the current CUDA stream; there is no
host synchronisation inside.  With the ``gloo`` backend (CPU tests) the post-spread stages run on
CPU tensors so the decomposition logic is testable without a GPU; the spread itself has no CPU
path.
"""

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .ops import get_frequency_array, nufft1, nufft2, nufft3

__all__ = [
    "shard_range", "split_transforms", "nufft1_stacked", "nufft2_stacked", "nufft1_sharded_points",
    "nufft2_sharded_points", "nufft3_sharded_sources", "nufft3_sharded_targets", "fine_grid_geometry",
    "slab_pencil_fft", "reduce_scatter_slabs", "exchange_points_by_slab", "halo_add", "slab_halo",
]


# ------------------------------------------------------------------------------------ partitioning
def _world(group):
    if not dist.is_available() or not dist.is_initialized():
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def shard_range(n, world, rank):
    """[lo, hi) of rank's share of n items: contiguous blocks, the first n % world one longer."""
    q, r = divmod(int(n), int(world))
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def split_transforms(source, group=None, axis=0):
    """This rank's slice of a stack of transforms that share points (split along ``axis``)."""
    world, rank = _world(group)
    lo, hi = shard_range(source.shape[axis], world, rank)
    return source.narrow(axis, lo, hi - lo)


def _all_gather_cat(x, group, axis, sizes):
    """all_gather of shards whose extent along ``axis`` differs per rank (sizes[r])."""
    world, _ = _world(group)
    if world == 1:
        return x
    mx = max(sizes)
    xm = torch.movedim(x, axis, 0).contiguous()
    if xm.shape[0] < mx:
        pad = torch.zeros((mx - xm.shape[0],) + tuple(xm.shape[1:]), dtype=x.dtype, device=x.device)
        xm = torch.cat([xm, pad], 0)
    parts = [torch.empty_like(xm) for _ in range(world)]
    dist.all_gather(parts, xm, group=group)
    return torch.movedim(torch.cat([p[:s] for p, s in zip(parts, sizes)], 0), 0, axis)


# ------------------------------------------------------------------------------------ stacked transforms
def nufft1_stacked(output_shape, source, *points, group=None, gather=False, **kw):
    """n_transf transforms sharing one point set (BASELINE config 4): ``source`` is the FULL stack
    (n_transf, M), replicated like the points; each rank transforms its slice of the stack (and
    repeats the cheap bin-sort).  No data-path collective.  Returns (n_transf_local, *output_shape)
    or, with ``gather``, the full stack on every rank."""
    world, _ = _world(group)
    local = split_transforms(source, group, axis=0)
    out = nufft1(output_shape, local, *points, **kw) if local.shape[0] else \
        torch.empty((0,) + tuple(output_shape), dtype=source.dtype, device=source.device)
    if not gather or world == 1:
        return out
    sizes = [shard_range(source.shape[0], world, r) for r in range(world)]
    return _all_gather_cat(out, group, 0, [b - a for a, b in sizes])


def nufft2_stacked(source, *points, group=None, gather=False, **kw):
    """Type-2 twin of :func:`nufft1_stacked`: ``source`` (n_transf, N1[, N2[, N3]]) full stack."""
    world, _ = _world(group)
    local = split_transforms(source, group, axis=0)
    out = nufft2(local, *points, **kw) if local.shape[0] else \
        torch.empty((0, points[0].shape[-1]), dtype=source.dtype, device=source.device)
    if not gather or world == 1:
        return out
    sizes = [shard_range(source.shape[0], world, r) for r in range(world)]
    return _all_gather_cat(out, group, 0, [b - a for a, b in sizes])


# ------------------------------------------------------------------------------------ sharded points
def nufft2_sharded_points(source, *points_local, group=None, **kw):
    """Type 2, points split across ranks, modes replicated: purely local (sharding_test.py:197-241)."""
    return nufft2(source, *points_local, **kw)


def nufft3_sharded_targets(source, *points, group=None, **kw):
    """Type 3, targets split across ranks (sources replicated): purely local."""
    return nufft3(source, *points, **kw)


def nufft3_sharded_sources(source_local, *points, group=None, **kw):
    """Type 3, sources split across ranks (targets replicated): local transform + sum over ranks
    (sharding_test.py:244-290).  Each rank derives its own grid from its own source extent, which
    is harmless: the sum of exact partial sums is the exact sum."""
    world, _ = _world(group)
    out = nufft3(source_local, *points, **kw)
    if world > 1:
        dist.all_reduce(torch.view_as_real(out), group=group)
    return out


def fine_grid_geometry(output_shape, eps, single, upsampfac=2.0, kerevalmeth=1):
    """(ns, beta, nf[...]) of the type-1/2 fine grid for JAX-ordered ``output_shape`` -- the same
    plan arithmetic the backend uses (V/src/cuda/spreadinterp.cpp:16-90, common.cu:166-177)."""
    L = _lib.lib()
    ns, beta = C.c_int(), C.c_double()
    ier = L.b2n_setup_spreader(float(eps), float(upsampfac), int(kerevalmeth), 0 if single else 1,
                               C.byref(ns), C.byref(beta))
    if ier > 1:
        raise RuntimeError(f"setup_spreader failed with code {ier}")
    nf = tuple(int(L.b2n_set_nf_type12(int(n), float(upsampfac), ns.value)) for n in output_shape)
    return ns.value, beta.value, nf


def _kernel_ft(nf, ns, beta):
    out = (C.c_double * (nf // 2 + 1))()
    _lib.lib().b2n_fseries(int(nf), int(ns), float(beta), out)
    return np.frombuffer(out, dtype=np.float64).copy()


def _mode_index(n, nf, modeord):
    """For each output mode (in the requested order): its FFT-ordered fine-grid index and |k|."""
    k = get_frequency_array(int(n), int(modeord))
    return np.where(k >= 0, k, k + nf), np.abs(k)


def reduce_scatter_slabs(grid, group=None):
    """Sum the ranks' private fine grids (nf3, nf2, nf1) and leave rank r with z-slabs
    [r*nf3/G, (r+1)*nf3/G).  nf3 must divide by the world size (always true for the even 2^a3^b5^c
    fine-grid sizes when G is 2, 4 or 8 and nf3 >= 16)."""
    world, rank = _world(group)
    if world == 1:
        return grid
    nf3 = grid.shape[0]
    if nf3 % world:
        raise ValueError(f"reduce_scatter path needs nf3 ({nf3}) divisible by the world size ({world})")
    flat = torch.view_as_real(grid.contiguous()).reshape(-1)
    out = torch.empty(flat.numel() // world, dtype=flat.dtype, device=flat.device)
    if grid.is_cuda:
        dist.reduce_scatter_tensor(out, flat, group=group)
    else:  # gloo (CPU tests of the host logic): no reduce_scatter -> all_reduce + slice
        dist.all_reduce(flat, group=group)
        out = flat.reshape(world, -1)[rank].clone()
    return torch.view_as_complex(out.reshape((nf3 // world,) + tuple(grid.shape[1:]) + (2,)))


def slab_pencil_fft(slab, output_shape, nf, iflag, ns, beta, modeord=0, group=None, gather=True):
    """Steps 3-5 of the native type-1 path on this rank's summed z-slabs ``slab`` (nf3/G, nf2, nf1):
    2-D FFT + crop, all_to_all transpose, 1-D FFT along z + crop + deconvolve (the index maps of
    V/src/cuda/deconvolve_wrapper.cu:76-118).  Returns modes (N3, N2_local, N1) in JAX axis order
    (output_shape = (N3, N2, N1) here, slowest first), y-sharded, or the full array with ``gather``."""
    world, rank = _world(group)
    N3, N2, N1 = (int(n) for n in output_shape)
    nf3, nf2, nf1 = (int(n) for n in nf)
    if slab.is_cuda:
        return _slab_pencil_fft_native(slab, (N3, N2, N1), (nf3, nf2, nf1), iflag, ns, beta, modeord, group, gather)
    rdt = torch.float32 if slab.dtype == torch.complex64 else torch.float64
    dev = slab.device
    idx = {}
    dec = {}
    for name, n, f in (("x", N1, nf1), ("y", N2, nf2), ("z", N3, nf3)):
        ii, ak = _mode_index(n, f, modeord)
        idx[name] = torch.as_tensor(ii, device=dev)
        dec[name] = torch.as_tensor(1.0 / _kernel_ft(f, ns, beta)[ak], dtype=rdt, device=dev)
    # 2-D FFT over (y, x); sign as cufft_ex(iflag) (V/include/cufinufft/types.h:108-115), unnormalised
    if iflag >= 0:
        s2 = torch.fft.ifft2(slab, dim=(1, 2), norm="forward")
    else:
        s2 = torch.fft.fft2(slab, dim=(1, 2))
    s2 = s2.index_select(2, idx["x"]).index_select(1, idx["y"])          # (nz_loc, N2, N1)
    s2 = s2 * (dec["y"][None, :, None] * dec["x"][None, None, :])
    # transpose z-slabs -> y-pencils
    ysz = [b - a for a, b in (shard_range(N2, world, r) for r in range(world))]
    if world > 1:
        send = [s2[:, a:b, :].contiguous() for a, b in (shard_range(N2, world, r) for r in range(world))]
        nzl = nf3 // world
        recv = [torch.empty((nzl, ysz[rank], N1), dtype=s2.dtype, device=dev) for _ in range(world)]
        if dev.type == "cuda":
            dist.all_to_all(recv, send, group=group)
        else:  # gloo has no all_to_all: emulate with one all_gather per destination
            for dst in range(world):
                parts = [torch.empty((nzl, ysz[dst], N1), dtype=s2.dtype, device=dev) for _ in range(world)]
                dist.all_gather(parts, send[dst], group=group)
                if dst == rank:
                    recv = parts
        pencil = torch.cat(recv, 0)                                       # (nf3, N2_loc, N1)
    else:
        pencil = s2
    pencil = torch.fft.ifft(pencil, dim=0, norm="forward") if iflag >= 0 else torch.fft.fft(pencil, dim=0)
    out = pencil.index_select(0, idx["z"]) * dec["z"][:, None, None]
    if gather and world > 1:
        out = _all_gather_cat(out, group, 1, ysz)
    return out


def _slab_pencil_fft_native(slab, N, nf, iflag, ns, beta, modeord, group, gather):
    """slab_pencil_fft on the device: the two stages are library calls (csrc/slab.cu: cuFFT plans cached
    per geometry, fused crop + deconvolve kernels), the transpose between them is ONE all_to_all whose
    split sizes follow from the shapes -- nothing is read back to the host, the slab is transformed in
    place."""
    world, rank = _world(group)
    N3, N2, N1 = N
    nf3, nf2, nf1 = nf
    L = _lib.lib()
    vp = C.c_void_p
    dbl = int(slab.dtype == torch.complex128)
    slab = slab.contiguous()
    nzl = slab.shape[0]
    assert slab.shape[1:] == (nf2, nf1) and nzl * world == nf3
    st = vp(torch.cuda.current_stream(slab.device).cuda_stream)
    ysz = [b - a for a, b in (shard_range(N2, world, r) for r in range(world))]
    send = torch.empty(nzl * N2 * N1, dtype=slab.dtype, device=slab.device)
    ier = L.b2n_slab_fft_xy(dbl, st, vp(slab.data_ptr()), nzl, nf2, nf1, N2, N1, world, int(iflag), int(modeord),
                            int(ns), float(beta), vp(send.data_ptr()))
    if ier:
        raise RuntimeError(f"b2n_slab_fft_xy failed with code {ier}")
    if world > 1:
        recv = torch.empty(nf3 * ysz[rank] * N1, dtype=slab.dtype, device=slab.device)
        dist.all_to_all_single(torch.view_as_real(recv), torch.view_as_real(send),
                               output_split_sizes=[nzl * ysz[rank] * N1] * world,
                               input_split_sizes=[nzl * ysz[d] * N1 for d in range(world)], group=group)
    else:
        recv = send
    out = torch.empty((N3, ysz[rank], N1), dtype=slab.dtype, device=slab.device)
    ier = L.b2n_slab_fft_z(dbl, st, vp(recv.data_ptr()), nf3, N3, ysz[rank] * N1, int(iflag), int(modeord), int(ns),
                           float(beta), vp(out.data_ptr()))
    if ier:
        raise RuntimeError(f"b2n_slab_fft_z failed with code {ier}")
    if gather and world > 1:
        out = _all_gather_cat(out, group, 1, ysz)
    return out


# ---------------------------------------------------------------------------- spatial (slab) split
def slab_halo(ns):
    """Halo planes on either side of a rank's z-slabs: the ns-wide window of a point inside the
    slab reaches at most ceil(ns/2) planes out; +2 keeps float32 rounding of the re-based
    coordinate away from the local grid's periodic seam."""
    return (int(ns) + 1) // 2 + 2


def _rows_all_to_all(rows, counts, group):
    """rows (M, W) grouped by destination rank, counts[r] rows for rank r -> the rows every rank
    sent to this one.  NCCL: all_to_all_single; gloo (CPU tests): all_gather + slice."""
    world, rank = _world(group)
    if world == 1:
        return rows
    cnt = counts.to(torch.int64) if torch.is_tensor(counts) else torch.as_tensor(counts, dtype=torch.int64, device=rows.device)
    if rows.is_cuda:
        got = torch.empty_like(cnt)
        dist.all_to_all_single(got, cnt, group=group)
        sent, got = torch.stack([cnt, got]).tolist()  # the one host read of this path: the split sizes
        out = torch.empty((sum(got), rows.shape[1]), dtype=rows.dtype, device=rows.device)
        dist.all_to_all_single(out, rows, output_split_sizes=got, input_split_sizes=sent, group=group)
        return out
    allc = [torch.empty_like(cnt) for _ in range(world)]
    dist.all_gather(allc, cnt, group=group)
    mmax = int(max(int(c.sum()) for c in allc))
    pad = torch.zeros((mmax, rows.shape[1]), dtype=rows.dtype)
    pad[: rows.shape[0]] = rows
    allr = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(allr, pad, group=group)
    parts = []
    for src in range(world):
        off = int(allc[src][:rank].sum())
        parts.append(allr[src][off: off + int(allc[src][rank])])
    return torch.cat(parts, 0)


def exchange_points_by_slab(source_local, points_local, nf0, ns, group=None):
    """Send every point to the rank that owns its fine-grid plane along axis 0 (the slowest grid
    axis, points_local[0]); rank r owns planes [r*L, (r+1)*L), L = nf0 / G.

    Returns (source, points, Lz): the strengths and points this rank received, with the axis-0
    coordinate RE-BASED to the rank's local grid of Lz = L + 2*halo planes (plane p of the local grid
    = global plane r*L - halo + p) and expressed as an angle in [-pi, pi) of that local grid, so
    that the unmodified spreader (which folds its input periodically) places it correctly.  The
    re-basing is done in float64 by the sender; the float32 the spreader sees is as accurate,
    in cells, as the original coordinate was."""
    world, rank = _world(group)
    L = int(nf0) // world
    h = slab_halo(ns)
    Lz = L + 2 * h
    z = points_local[0]
    rdt = z.dtype
    if z.is_cuda and len(points_local) == 3:  # one native pass (csrc/slab.cu) instead of ~20 torch kernels
        M = z.numel()
        cdt = torch.complex64 if rdt == torch.float32 else torch.complex128
        pp = [q.to(rdt).contiguous() for q in points_local]
        cc = source_local.to(cdt).contiguous()
        outs = [torch.empty(M, dtype=rdt, device=z.device) for _ in range(3)] + [torch.empty(M, dtype=cdt, device=z.device)]
        cnt = torch.empty(2 * world, dtype=torch.int64, device=z.device)
        vp = C.c_void_p
        ier = _lib.lib().b2n_slab_partition(int(rdt == torch.float64), vp(torch.cuda.current_stream(z.device).cuda_stream),
                                            M, vp(pp[0].data_ptr()), vp(pp[1].data_ptr()), vp(pp[2].data_ptr()),
                                            vp(cc.data_ptr()), int(nf0), world, h, *[vp(o.data_ptr()) for o in outs],
                                            vp(cnt.data_ptr()))
        if ier:
            raise RuntimeError(f"b2n_slab_partition failed with code {ier}")
        if world > 1:  # the arrays travel as they are (structure of arrays): nothing to unpack on arrival
            send = cnt[:world]
            got = torch.empty_like(send)
            dist.all_to_all_single(got, send, group=group)
            sent, got = torch.stack([send, got]).tolist()  # the one host read of this path: the split sizes
            recv = []
            for o in outs:
                flat = torch.view_as_real(o) if o.is_complex() else o
                w = 2 if o.is_complex() else 1
                r = torch.empty((sum(got),) + tuple(flat.shape[1:]), dtype=flat.dtype, device=z.device)
                dist.all_to_all_single(r, flat, output_split_sizes=got, input_split_sizes=sent, group=group)
                recv.append(torch.view_as_complex(r) if w == 2 else r)
            outs = recv
        return outs[3], outs[:3], Lz
    two_pi = 2.0 * np.pi
    t = z.to(torch.float64) / two_pi + 0.5
    zf = (t - torch.floor(t)) * float(nf0)                       # fold_rescale (csrc/common.cuh), in float64
    owner = torch.clamp((zf / L).floor().to(torch.int64), 0, world - 1)
    z_in = ((zf - (owner * L).to(torch.float64) + h) * (two_pi / Lz) - np.pi).to(rdt)
    rows = torch.stack([z_in] + [q.to(rdt) for q in points_local[1:]] +
                       [source_local.real.to(rdt), source_local.imag.to(rdt)], dim=1)
    if world > 1:
        order = torch.argsort(owner)
        rows = rows.index_select(0, order)
        counts = torch.bincount(owner, minlength=world).tolist()
        rows = _rows_all_to_all(rows.contiguous(), counts, group)
    nd = len(points_local)
    pts = [rows[:, d].contiguous() for d in range(nd)]
    src = torch.complex(rows[:, nd].contiguous(), rows[:, nd + 1].contiguous())
    return src, pts, Lz


def halo_add(local, h, group=None):
    """local (L + 2h, ...) spread by this rank -> its L summed planes: the h planes below / above the
    slab belong to the previous / next rank (periodically) and are added there."""
    world, rank = _world(group)
    Lz = local.shape[0]
    L = Lz - 2 * h
    if L < h:
        raise ValueError(f"slab of {L} planes is thinner than the halo ({h})")
    lo, hi = local[:h], local[L + h:]
    if world == 1:
        local[L: L + h] += lo
        local[h: 2 * h] += hi
        return local[h: L + h]
    prev, nxt = (rank - 1) % world, (rank + 1) % world
    lo_s, hi_s = lo.contiguous(), hi.contiguous()
    from_next, from_prev = torch.empty_like(lo_s), torch.empty_like(hi_s)
    # gloo has no complex send/recv: move the real views
    v = torch.view_as_real
    ops = [dist.P2POp(dist.isend, v(lo_s), prev, group=group, tag=1),
           dist.P2POp(dist.isend, v(hi_s), nxt, group=group, tag=2),
           dist.P2POp(dist.irecv, v(from_next), nxt, group=group, tag=1),
           dist.P2POp(dist.irecv, v(from_prev), prev, group=group, tag=2)]
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    slab = local[h: L + h]
    slab[L - h:] += from_next   # the next rank's low halo = my top planes
    slab[:h] += from_prev       # the previous rank's high halo = my bottom planes
    return slab


_SPREAD_PLANS = {}  # spread-only plans, kept across calls (sort workspaces are grow-only)


def _spread_only(nf, c, pts, eps, iflag, upsampfac):
    """Private fine grid (nf3, nf2, nf1) of this rank's points: the backend's spreader alone
    (gpu_spreadinterponly, V/include/cufinufft/impl.h:115-117) with the type-1 plan's kernel."""
    from .plan import Plan

    key = (tuple(int(n) for n in nf), str(c.dtype), float(eps), int(iflag), float(upsampfac), c.device.index,
           torch.cuda.current_stream(c.device).cuda_stream)  # a plan enqueues on the stream it was made on
    p = _SPREAD_PLANS.pop(key, None)
    if p is None:
        p = Plan(1, tuple(nf[::-1]), n_trans=1, eps=eps, isign=iflag, dtype=str(c.dtype).replace("torch.", ""),
                 gpu_spreadinterponly=1, upsampfac=float(upsampfac))
    try:
        p.setpts(*pts[::-1])  # backend order: x (fastest) first
        out = p.execute(c.reshape(1, -1))[0]
    except Exception:
        p.destroy()
        raise
    while len(_SPREAD_PLANS) >= 4:
        _SPREAD_PLANS.pop(next(iter(_SPREAD_PLANS))).destroy()
    _SPREAD_PLANS[key] = p
    return out


def nufft1_sharded_points(output_shape, source_local, *points_local, group=None, combine="reduce_scatter",
                          gather=True, iflag=1, eps=1e-6, opts=None):
    """3-D (or 2-D/1-D with ``combine="psum"``) type 1 with the POINTS split across ranks by index
    range; ``source_local`` (M_local,) and ``points_local`` are this rank's shard.

    combine="psum"            reference-equivalent: local nufft1 + all_reduce of the modes
                              (sharding_test.py:163-165).  Output replicated (with an explicit
                              combine="psum", whatever ``gather`` says; the collective is outside autograd).
    combine="reduce_scatter"  native path (module docstring).  3-D, single transform.  Output
                              y-sharded (N3, N2_local, N1) unless ``gather``.
    combine="slab"            spatial split (module docstring): points exchanged by z-slab, spread
                              into slab + halo, halo planes added at the neighbours.  Same output.
    """
    world, _ = _world(group)
    if combine == "auto":
        # measured on B200: the exchange of the slab split costs about what one
        # 512^3 fine-grid FFT does, so it pays once the fine grid is larger than that
        nftot = 1
        for n in output_shape:
            nftot *= 2 * int(n)
        combine = "slab" if (len(points_local) == 3 and source_local.ndim == 1 and nftot > 300_000_000) else "psum"
    if combine == "psum" or len(points_local) != 3 or source_local.ndim != 1:
        out = nufft1(output_shape, source_local, *points_local, iflag=iflag, eps=eps, opts=opts)
        if world > 1:
            # the collective is not differentiable: reduce a detached copy, never the autograd output in place
            out = out.detach().clone() if out.requires_grad else out
            dist.all_reduce(torch.view_as_real(out), group=group)
        return out
    if combine not in ("reduce_scatter", "slab"):
        raise ValueError("combine must be 'psum', 'reduce_scatter', 'slab' or 'auto'")
    from . import options

    o = options.unpack_opts(opts, 1, True) or options.Opts()
    single = source_local.dtype == torch.complex64
    ns, beta, nf = fine_grid_geometry(output_shape, eps, single, upsampfac=o.gpu_upsampfac,
                                      kerevalmeth=int(o.gpu_kerevalmeth))
    if nf[0] % world:  # the slabs do not divide: psum, but keep the layout the caller asked for
        out = nufft1_sharded_points(output_shape, source_local, *points_local, group=group, combine="psum",
                                    iflag=iflag, eps=eps, opts=opts)
        if not gather and world > 1:
            ylo, yhi = shard_range(int(output_shape[1]), world, _world(group)[1])
            out = out[:, ylo:yhi].contiguous()
        return out
    if combine == "slab" and nf[0] // world >= 2 * slab_halo(ns):
        src, pts, Lz = exchange_points_by_slab(source_local, list(points_local), nf[0], ns, group)
        local = _spread_only((Lz, nf[1], nf[2]), src, pts, eps, iflag, o.gpu_upsampfac)
        slab = halo_add(local, slab_halo(ns), group)
    else:
        grid = _spread_only(nf, source_local, list(points_local), eps, iflag, o.gpu_upsampfac)
        slab = reduce_scatter_slabs(grid, group)
        del grid
    return slab_pencil_fft(slab, output_shape, nf, iflag, ns, beta, modeord=int(o.modeord), group=group,
                           gather=gather)
