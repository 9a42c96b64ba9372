#!/usr/bin/env python
"""Run blur_act_nhwc and rgb_finalize at the 256x256 generator shapes (for ncu)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-fm-gan_b200"))
import torch  # noqa: E402
from fm3d import ops  # noqa: E402

dev = torch.device("cuda:0")
B, H, C = 32, 256, 128
t = torch.randn(B, H + 1, H + 1, C, device=dev).to(torch.bfloat16)
k = torch.tensor([1., 3., 3., 1.]); k = torch.outer(k, k); k = (k / k.sum() * 4).to(dev)
tab = torch.zeros(B, C, 8, device=dev); tab[..., 0] = 1; tab[..., 2] = 0.2; tab[..., 3] = 1
noise = torch.randn(B, H, H, device=dev); nw = torch.ones(1, device=dev)
out = torch.empty(B, H, H, C, device=dev, dtype=torch.bfloat16)
acc = torch.randn(B, H, H, 4, device=dev); skip = torch.randn(B, 3, H // 2, H // 2, device=dev); bias = torch.zeros(3, device=dev)
rgb = torch.empty(B, 3, H, H, device=dev)


def run():
    ops.blur_act_nhwc(t, k, tab, noise, True, nw, C, out=out)
    ops.rgb_finalize(acc, bias, skip, k, out=rgb)


for _ in range(3):
    run()
torch.cuda.synchronize()
for name, fn, nbytes in (("blur_act", lambda: ops.blur_act_nhwc(t, k, tab, noise, True, nw, C, out=out), t.numel() * 2 + out.numel() * 2),
                         ("rgb_finalize", lambda: ops.rgb_finalize(acc, bias, skip, k, out=rgb), acc.numel() * 8 + rgb.numel() * 4 + skip.numel() * 4)):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name}: {ms:.3f} ms  {nbytes / ms / 1e6:.0f} GB/s")
