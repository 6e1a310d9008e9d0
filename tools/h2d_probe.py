"""Host<->device copy bandwidth of the box (pinned), to put bench.py's e2e number in context."""
import time
import torch

n = 1 << 28  # 1 GiB of float32
h = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device="cuda")
for name, a, b in (("h2d", d, h), ("d2h", h, d)):
    for _ in range(2):
        a.copy_(b, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        a.copy_(b, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(f"{name}: {n * 4 / dt / 1e9:.1f} GB/s pinned")
t0 = time.perf_counter()
h2 = torch.empty(n, dtype=torch.float32).pin_memory()
print(f"pin 1 GiB: {time.perf_counter() - t0:.2f} s")
