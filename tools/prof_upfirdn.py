#!/usr/bin/env python
"""Run the generator's largest blur (upfirdn2d up=down=1, pad (1,1), [32*128, 257, 257]) a few times
(for `ncu --set full -k regex:upfirdn2d`).  usage: prof_upfirdn.py [f32|bf16]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-fm-gan_b200"))
import torch  # noqa: E402
from fm3d import ops  # noqa: E402

dt = torch.bfloat16 if len(sys.argv) > 1 and sys.argv[1] == "bf16" else torch.float32
dev = torch.device("cuda:0")
k = torch.tensor([1., 3., 3., 1.]); k = (k[None] * k[:, None]); k = (k / k.sum() * 4).to(dev)
x = torch.randn(32, 128, 257, 257, device=dev, dtype=dt)
for _ in range(3):
    y = ops.upfirdn2d_planes(x, k, 1, 1, 1, 1, 1, 1, 1, 1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    y = ops.upfirdn2d_planes(x, k, 1, 1, 1, 1, 1, 1, 1, 1)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"upfirdn2d {dt}: {ms:.3f} ms  {(x.numel() + y.numel()) * x.element_size() / ms / 1e6:.0f} GB/s")
