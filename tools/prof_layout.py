#!/usr/bin/env python
"""fp32 NCHW -> bf16 NHWC layout pass of the training path (fm_nchw_to_nhwc_bf16).  usage: prof_layout.py H C [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-fm-gan_b200"))
import torch  # noqa: E402
from fm3d import ops  # noqa: E402

H, C = int(sys.argv[1]), int(sys.argv[2])
B = int(sys.argv[3]) if len(sys.argv) > 3 else 32
dev = torch.device("cuda:0")
x = torch.randn(B, C, H, H, device=dev)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
ref = x.permute(0, 2, 3, 1).to(torch.bfloat16)
y = ops.nchw_to_nhwc_bf16(x)
assert torch.equal(y[..., :C], ref), "layout pass differs from the torch permute + cast"
for _ in range(3):
    y = ops.nchw_to_nhwc_bf16(x)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    y = ops.nchw_to_nhwc_bf16(x)
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = sorted(ts)[2]
print(f"nchw->nhwc bf16 [{B},{C},{H},{H}]: {ms * 1e3:.1f} us  {x.numel() * 6 / ms / 1e6:.0f} GB/s", flush=True)
