#!/usr/bin/env python
"""Per-CTA event timeline of the pSp stem launch (3x3 conv on 3 channels, 256^2, strip mode) with FM3D_TRACE=1."""
import ctypes as C
import os
import sys

os.environ.setdefault("FM3D_TRACE", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-fm-gan_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from fm3d import _lib, ops  # noqa: E402

B, S = 32, 256
dev = torch.device("cuda:0")
Hp, Wp = S + 2, S + 8
packed = torch.zeros(B, Hp, Wp, 8, device=dev, dtype=torch.bfloat16)
packed[:, 1:S + 1, 1:S + 1, :3] = torch.randn(B, S, S, 3, device=dev).to(torch.bfloat16)
w = (torch.randn(3, 64, 64, device=dev) * 0.1).to(torch.bfloat16)
tab = torch.zeros(1, 64, 8, device=dev); tab[..., 0] = 1; tab[..., 2] = 0.25; tab[..., 3] = 1
out = torch.empty(B, S, S, 64, device=dev, dtype=torch.bfloat16)


def run():
    ops.conv_igemm(packed, w, [(ky, 0, ky) for ky in range(3)], out, tab, B=B, H=Hp, W=S, Cin=24, Cout=64, OH=S, OW=S,
                   stride_x=1, stride_y=1, x_pixstride=8, x_rowstride=Wp * 8, x_imgstride=Hp * Wp * 8)


for _ in range(3):
    run()
torch.cuda.synchronize()
n = 148
buf = np.zeros((n, 64), dtype=np.int64)
slots = C.c_int(0)
_lib.check(_lib.lib().fm_igemm_trace(buf.ctypes.data, n, C.byref(slots)), "fm_igemm_trace")
t0 = buf[:, 0].min()
for cta in (0, 73, 147):
    r = buf[cta]
    print(f"CTA {cta}: start {r[0] - t0}, first loads issued {r[1] - t0}, end {r[63] - t0}")
    for t in range(15):
        s = r[2 + 4 * t: 6 + 4 * t]
        if s[3] == 0 and s[1] == 0:
            break
        print(f"   tile {t}: epi done {s[0] - r[0]:7d}  acc free {s[1] - r[0]:7d}  strip landed {s[2] - r[0]:7d}  acc complete {s[3] - r[0]:7d}")
print("kernel span (clk):", buf[:, 63].max() - t0)
