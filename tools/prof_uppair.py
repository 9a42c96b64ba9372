#!/usr/bin/env python
"""Paired-parity launch of the last up-conv (128x128, 256 -> 128, B=32), as the engine issues it
(for `ncu --set full -k regex:igemm_conv`).  usage: prof_uppair.py [py]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-fm-gan_b200"))
import torch  # noqa: E402
from fm3d import ops, engine  # noqa: E402

py = int(sys.argv[1]) if len(sys.argv) > 1 else 0
B, h, Cin, Cout = 32, 128, 256, 128
dev = torch.device("cuda:0")
x = torch.randn(B, h, h, Cin, device=dev).to(torch.bfloat16)
views = engine._PAIR_VIEWS[py]
w = (torch.randn(len(views), 2 * Cout, Cin, device=dev) / (Cin * 9) ** 0.5).to(torch.bfloat16)
tab = torch.zeros(1, 2 * Cout, 8, device=dev); tab[..., 0] = 1; tab[..., 2] = 1; tab[..., 3] = 1
t = torch.empty(B, 2 * h + 1, 2 * h + 1, Cout, device=dev, dtype=torch.bfloat16)
taps = [(dy, dx, v) for v, ((dy, dx), _, _) in enumerate(views)]


def run():
    ops.conv_igemm(x, w, taps, t, None, B=B, H=h, W=h, Cin=Cin, Cout=2 * Cout, OH=h + 1 - py, OW=h + 1, out_H=2 * h + 1,
                   out_W=2 * h + 1, out_y0=py, out_x0=0, out_ys=2, out_xs=2, tab_per_sample=False, out_cgroup=Cout,
                   out_gstride=Cout, out_cstride=Cout, out_cgroup_ow_shrink=1)


for _ in range(4):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    run()
e1.record(); torch.cuda.synchronize()
print(f"uppair py={py}: {e0.elapsed_time(e1) / 5:.3f} ms")
