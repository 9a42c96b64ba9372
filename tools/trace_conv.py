#!/usr/bin/env python
"""Per-CTA event timeline of one igemm launch (FM3D_TRACE=1): where a tile's time goes.
usage: FM3D_TRACE=1 python tools/trace_conv.py H Cin Cout [B]   (plain 3x3 conv, shared epilogue table)"""
import ctypes as C
import os
import sys

os.environ.setdefault("FM3D_TRACE", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-fm-gan_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from fm3d import _lib, ops  # noqa: E402

H, Cin, Cout = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
B = int(sys.argv[4]) if len(sys.argv) > 4 else 32
dev = torch.device("cuda:0")
x = torch.randn(B, H, H, Cin, device=dev).to(torch.bfloat16)
w = (torch.randn(9, Cout, Cin, device=dev) / (Cin * 9) ** 0.5).to(torch.bfloat16)
tab = torch.zeros(1, Cout, 8, device=dev); tab[..., 0] = 1; tab[..., 2] = 0.0; tab[..., 3] = 1
out = torch.empty(B, H, H, Cout, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    ops.conv_igemm(x, w, ops.conv_taps(3, 3, 1), out, tab, B=B, H=H, W=H, Cin=Cin, Cout=Cout, OH=H, OW=H, tab_per_sample=False)
torch.cuda.synchronize()
n = 148
buf = np.zeros((n, 64), dtype=np.int64)
slots = C.c_int(0)
_lib.check(_lib.lib().fm_igemm_trace(buf.ctypes.data, n, C.byref(slots)), "fm_igemm_trace")
t0 = buf[:, 0].min()
for cta in (0, 1, 73, 147):
    r = buf[cta]
    print(f"CTA {cta}: start {r[0] - t0}, first loads issued {r[1] - t0}, end {r[63] - t0}")
    for t in range(15):
        s = r[2 + 4 * t: 6 + 4 * t]
        if s[3] == 0 and s[1] == 0:
            break
        print(f"   tile {t}: epi done {s[0] - r[0]:7d}  acc free {s[1] - r[0]:7d}  patch landed {s[2] - r[0]:7d}  acc complete {s[3] - r[0]:7d}")
r = buf[0]
if r[32] > 0:
    print("CTA 0, tile 1, chunk 0, per tap: weights landed / MMAs issued (clk since kernel start):")
    print("   " + "  ".join(f"{r[32 + 2 * i] - r[0]}/{r[33 + 2 * i] - r[0]}" for i in range(9) if r[32 + 2 * i] > 0))
print("kernel span (clk):", buf[:, 63].max() - t0)
