#!/usr/bin/env python
"""Run one implicit-GEMM conv shape a few times (for `ncu --set full -k regex:igemm`).
usage: prof_conv.py H Cin Cout [rgb] [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-fm-gan_b200"))
import torch  # noqa: E402
from fm3d import ops  # noqa: E402

H, Cin, Cout = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
use_rgb = len(sys.argv) > 4 and sys.argv[4] == "rgb"
B = int(sys.argv[5]) if len(sys.argv) > 5 else 32
dev = torch.device("cuda:0")
x = torch.randn(B, H, H, Cin, device=dev).to(torch.bfloat16)
w = (torch.randn(9, Cout, Cin, device=dev) / (Cin * 9) ** 0.5).to(torch.bfloat16)
tab = torch.zeros(B, Cout, 8, device=dev); tab[..., 0] = 1; tab[..., 2] = 0.2; tab[..., 3] = 1; tab[..., 4:7] = 0.01
out = torch.empty(B, H, H, Cout, device=dev, dtype=torch.bfloat16)
noise = torch.randn(B, H, H, device=dev)
nw = torch.ones(1, device=dev)
rgb = torch.zeros(B, H, H, 4, device=dev) if use_rgb else None
for _ in range(4):
    ops.conv_igemm(x, w, ops.conv_taps(3, 3, 1), out, tab, B=B, H=H, W=H, Cin=Cin, Cout=Cout, OH=H, OW=H,
                   tab_per_sample=True, noise=noise, noise_w=nw, rgb=rgb)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.conv_igemm(x, w, ops.conv_taps(3, 3, 1), out, tab, B=B, H=H, W=H, Cin=Cin, Cout=Cout, OH=H, OW=H,
                   tab_per_sample=True, noise=noise, noise_w=nw, rgb=rgb)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"conv {H}x{H} {Cin}->{Cout} rgb={use_rgb}: {ms:.3f} ms  {2.0 * B * H * H * Cin * Cout * 9 / ms / 1e9:.1f} TFLOP/s")
