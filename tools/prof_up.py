#!/usr/bin/env python
"""The engine's stride-2 transposed conv (merged-phase launch on the tall image) at one generator shape, for
`ncu --set full -k regex:igemm_conv`.  usage: prof_up.py h Cin Cout [B]"""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-fm-gan_b200"))
import torch  # noqa: E402
from fm3d import engine  # noqa: E402

h, Cin, Cout = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
B = int(sys.argv[4]) if len(sys.argv) > 4 else 32
dev = torch.device("cuda:0")
x = torch.zeros(B, h + 1, h + 1, Cin, device=dev, dtype=torch.bfloat16)
x[:, :h, :h] = torch.randn(B, h, h, Cin, device=dev).to(torch.bfloat16)
L = types.SimpleNamespace(cin=Cin, cout=Cout, wpair=None,
                          wq=(torch.randn(9, Cout, Cin, device=dev) / (Cin * 9) ** 0.5).to(torch.bfloat16))
t = torch.zeros(B, 2 * h + 2, 2 * h + 2, Cout, device=dev, dtype=torch.bfloat16)
xt, tt = x.view(1, B * (h + 1), h + 1, Cin), t.view(1, B * (2 * h + 2), 2 * h + 2, Cout)


def run():
    engine.SynthesisPlan._up_conv(None, L, xt, tt, B, h)


for _ in range(4):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"up-conv {h}->{2 * h} {Cin}->{Cout}: {ms:.3f} ms  {2.0 * B * h * h * Cin * Cout * 9 / ms / 1e9:.1f} TFLOP/s")
