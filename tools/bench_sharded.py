#!/usr/bin/env python
"""ONE 3-D type-1 transform whose M points are split across the ranks by index range (strong
scaling), through the three paths of jax_finufft_b200/parallel.py:
slab (spatial split), reduce_scatter (private fine grids + NCCL reduce-scatter), psum (what
jax-finufft's shard_map gives).  Run under torchrun; rank 0 prints one JSON line per grid size.
Timing: CUDA events after warm-up, barrier on both sides, max over ranks."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jax_finufft_b200 import _lib  # noqa: E402
from jax_finufft_b200 import parallel as P  # noqa: E402


def main():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    M = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
    sizes = [int(a) for a in sys.argv[2:]] or [256, 512]
    Ml = M // world
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    pts = [(torch.rand(Ml, generator=g, device=dev) * 2 - 1) * np.pi for _ in range(3)]
    c = torch.complex(torch.rand(Ml, generator=g, device=dev) * 2 - 1, torch.rand(Ml, generator=g, device=dev) * 2 - 1)

    def timed(fn, K=3, W=2):
        for _ in range(W):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / K], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    for n in sizes:
        nm = (n, n, n)
        res = {}
        for mode in ("slab", "reduce_scatter", "psum"):
            torch.cuda.empty_cache()
            ms = timed(lambda: P.nufft1_sharded_points(nm, c, *pts, combine=mode, gather=False, eps=1e-6, iflag=1))
            res[mode] = round(ms, 3)
            _lib.lib().b2n_cache_clear()
            for p in list(P._SPREAD_PLANS.values()):
                p.destroy()
            P._SPREAD_PLANS.clear()
        if rank == 0:
            print(json.dumps({"workload": f"3-D type 1, ONE transform, M={world * Ml} split by index range over {world} GPU(s), "
                                          f"N={n}^3, eps=1e-6, c64", "n_gpus": world, "ms_per_step": res,
                              "points_per_s": {k: world * Ml / (v * 1e-3) for k, v in res.items()}}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
