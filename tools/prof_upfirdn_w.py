#!/usr/bin/env python
"""upfirdn2d (up = down = 1, pad (1,1): the Blur after an up-conv) at a given plane size.  usage: prof_upfirdn_w.py W planes [f32|bf16]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-fm-gan_b200"))
import torch  # noqa: E402
from fm3d import ops  # noqa: E402

W, planes = int(sys.argv[1]), int(sys.argv[2])
dt = torch.bfloat16 if len(sys.argv) > 3 and sys.argv[3] == "bf16" else torch.float32
dev = torch.device("cuda:0")
k = torch.tensor([1., 3., 3., 1.]); k = (k[None] * k[:, None]); k = (k / k.sum() * 4).to(dev)
x = torch.randn(planes, W, W, device=dev, dtype=dt)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
for _ in range(3):
    y = ops.upfirdn2d_planes(x, k, 1, 1, 1, 1, 1, 1, 1, 1)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    y = ops.upfirdn2d_planes(x, k, 1, 1, 1, 1, 1, 1, 1, 1)
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = sorted(ts)[2]
print(f"upfirdn2d {dt} {planes}x{W}x{W}: {ms * 1e3:.1f} us  {(x.numel() + y.numel()) * x.element_size() / ms / 1e6:.0f} GB/s", flush=True)
