"""Launches the small kernels added at the end of round 1 at C3 size (M = 1e8) so that ncu can
time them: k_pts_signature / k_sig_decide (setpts cache) and k_slab_count / k_slab_scatter
(spatial multi-GPU split, as if for 8 ranks).
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      -k regex:'k_pts_signature|k_sig_decide|k_slab' --csv python tools/probe_new_kernels.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jax_finufft_b200 import _lib  # noqa: E402
from jax_finufft_b200.plan import Plan  # noqa: E402

L = _lib.lib()
M = 100_000_000
g = torch.Generator(device="cuda").manual_seed(1)
pts = [(torch.rand(M, generator=g, device="cuda") * 2 - 1) * np.pi for _ in range(3)]
c = torch.randn(M, generator=g, device="cuda", dtype=torch.complex64)
L.b2n_set_setpts_cache(1)
p = Plan(1, (256, 256, 256), eps=1e-6)
for _ in range(3):
    p.setpts(*pts)
torch.cuda.synchronize()
p.destroy()
L.b2n_set_setpts_cache(0)
outs = [torch.empty(M, device="cuda") for _ in range(3)] + [torch.empty(M, device="cuda", dtype=torch.complex64)]
cnt = torch.empty(16, dtype=torch.int64, device="cuda")
vp = C.c_void_p
for _ in range(3):
    L.b2n_slab_partition(0, vp(torch.cuda.current_stream().cuda_stream), M, *[vp(q.data_ptr()) for q in pts], vp(c.data_ptr()),
                         512, 8, 6, *[vp(o.data_ptr()) for o in outs], vp(cnt.data_ptr()))
torch.cuda.synchronize()
print("counts", cnt[:8].tolist())
