#!/usr/bin/env python
"""Run one weight-gradient GEMM shape a few times (for `ncu --set full -k regex:wgrad_kernel`).
usage: prof_wgrad.py H Cin Cout [k] [stride] [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-fm-gan_b200"))
import torch  # noqa: E402
from fm3d import convgrad  # noqa: E402

H, Cin, Cout = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
k = int(sys.argv[4]) if len(sys.argv) > 4 else 3
s = int(sys.argv[5]) if len(sys.argv) > 5 else 1
B = int(sys.argv[6]) if len(sys.argv) > 6 else 32
p = k // 2 if s == 1 else 0
OH = (H + 2 * p - k) // s + 1
dev = torch.device("cuda:0")
x = torch.randn(B, H, H, Cin, device=dev).to(torch.bfloat16)
g = torch.randn(B, OH, OH, Cout, device=dev).to(torch.bfloat16)
shifts = [(ky - p, kx - p) for ky in range(k) for kx in range(k)]
if Cout >= Cin:
    fn = lambda: convgrad.conv_wgrad(g, x, Cout, Cin, B, OH, OH, [(0, 0, dy, dx) for (dy, dx) in shifts], 1, s)
else:
    fn = lambda: convgrad.conv_wgrad(x, g, Cin, Cout, B, OH, OH, [(dy, dx, 0, 0) for (dy, dx) in shifts], s, 1)
for _ in range(4):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    fn()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"wgrad {H}x{H} {Cin}->{Cout} k{k} s{s} B{B}: {ms:.3f} ms (incl. the memset of dW)  {2.0 * B * OH * OH * Cin * Cout * k * k / ms / 1e9:.1f} TFLOP/s")
