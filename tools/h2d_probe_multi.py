#!/usr/bin/env python
"""Concurrent pinned host->device bandwidth of N ranks (torchrun): what bounds bench.py's e2e at N>1.
Every rank copies 1 GiB from its own pinned buffer at the same time (barrier, 5 copies), rank 0
prints one JSON line with per-rank and aggregate GB/s, the NUMA node of every GPU and of the buffers."""
import json, os, time
import torch, torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 28
h = torch.empty(n, dtype=torch.float32).pin_memory()
h.fill_(1.0)
d = torch.empty(n, dtype=torch.float32, device="cuda")
for _ in range(2):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
gbs = 5 * n * 4 / (time.perf_counter() - t0) / 1e9
vals = [None] * world
if world > 1:
    dist.all_gather_object(vals, gbs)
else:
    vals = [gbs]
if rank == 0:
    numa = {}
    try:
        import subprocess
        q = subprocess.run(["nvidia-smi", "--query-gpu=index,pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True).stdout
        for ln in q.strip().splitlines():
            i, bus = [t.strip() for t in ln.split(",")]
            pth = f"/sys/bus/pci/devices/{bus[4:].lower()}/numa_node"
            numa[i] = open(pth).read().strip() if os.path.exists(pth) else "?"
    except Exception as e:  # noqa: BLE001
        numa = {"error": str(e)}
    nodes = len([d_ for d_ in os.listdir("/sys/devices/system/node") if d_.startswith("node")]) if os.path.isdir("/sys/devices/system/node") else None
    print(json.dumps({"ranks": world, "h2d_gbs_per_rank": [round(v, 1) for v in vals], "h2d_gbs_aggregate": round(sum(vals), 1),
                      "gpu_numa_node": numa, "host_numa_nodes": nodes, "host_cores": os.cpu_count()}))
if world > 1:
    dist.destroy_process_group()
