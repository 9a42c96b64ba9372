#!/usr/bin/env python
"""Run one of the encoders' small-K layer shapes a few times (for `ncu --set full -k regex:igemm`), checked once against
fp32 F.conv2d on the same bf16 operands.
usage: prof_enc.py res64 | lat2 | lat1 | sc64 | stem | c64_256 | all   [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-fm-gan_b200"))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from fm3d import ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def table(cout, slope):
    t = torch.zeros(1, cout, 8, device=dev)
    t[..., 0] = torch.rand(cout, device=dev) + 0.5
    t[..., 1] = torch.randn(cout, device=dev) * 0.1
    t[..., 2] = slope
    t[..., 3] = 1.0
    return t


def ref_conv(x, w, k, stride, pad, tab, residual=None):
    """x bf16 NHWC, w bf16 [k*k, O, I] -> fp32 NHWC of the epilogue's value."""
    O, I = w.shape[1], w.shape[2]
    w4 = w.float().view(k, k, O, I).permute(2, 3, 0, 1).contiguous()
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w4, stride=stride, padding=pad).permute(0, 2, 3, 1)
    y = y * tab[0, :, 0] + tab[0, :, 1]
    if residual is not None:
        y = y + residual
    return torch.where(y > 0, y, y * tab[0, :, 2]) * tab[0, :, 3]


def timeit(fn, name, flops, bytes_):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    print(f"{name}: {us:.1f} us  {flops / us / 1e6:.1f} TFLOP/s  {bytes_ / us / 1e3:.2f} GB/s (algorithmic in+out)", flush=True)


def check(out, ref, name, nb=2):
    err = float((out[:nb].float() - ref).abs().max() / ref.abs().max())
    assert err < 1e-2, f"{name}: rel err {err}"
    print(f"{name}: parity ok ({err:.2e})", flush=True)


def run(kind):
    if kind in ("res64", "c64_256", "c128_64", "c256_32"):
        H, Cc = {"res64": (64, 64), "c64_256": (256, 64), "c128_64": (64, 128), "c256_32": (32, 256)}[kind]
        x = torch.randn(B, H, H, Cc, device=dev).to(torch.bfloat16)
        w = (torch.randn(9, Cc, Cc, device=dev) / (Cc * 9) ** 0.5).to(torch.bfloat16)
        res = torch.randn(B, H, H, Cc, device=dev).to(torch.bfloat16) if kind == "res64" else None
        tab = table(Cc, 0.0)
        out = torch.empty(B, H, H, Cc, device=dev, dtype=torch.bfloat16)
        fn = lambda: ops.conv_igemm(x, w, ops.conv_taps(3, 3, 1), out, tab, B=B, H=H, W=H, Cin=Cc, Cout=Cc, OH=H, OW=H, residual=res)
        fn()
        check(out, ref_conv(x[:2], w, 3, 1, 1, tab, None if res is None else res[:2].float()), kind)
        timeit(fn, kind, 2.0 * B * H * H * Cc * Cc * 9, B * H * H * Cc * 2 * (3 if res is not None else 2))
    elif kind in ("lat2", "lat1"):
        H, Cin = (64, 128) if kind == "lat2" else (32, 256)
        x = torch.randn(B, H, H, Cin, device=dev).to(torch.bfloat16)
        w = (torch.randn(1, 512, Cin, device=dev) / Cin ** 0.5).to(torch.bfloat16)
        low = torch.randn(B, H // 2, H // 2, 512, device=dev).to(torch.bfloat16)
        tab = table(512, 1.0)
        out = torch.empty(B, H, H, 512, device=dev, dtype=torch.bfloat16)
        fn = lambda: ops.conv_igemm(x, w, [(0, 0, 0)], out, tab, B=B, H=H, W=H, Cin=Cin, Cout=512, OH=H, OW=H, residual_up=low)
        fn()
        up = F.interpolate(low[:2].float().permute(0, 3, 1, 2), size=(H, H), mode="bilinear", align_corners=True).permute(0, 2, 3, 1)
        check(out, ref_conv(x[:2], w, 1, 1, 0, tab, up), kind)
        timeit(fn, kind, 2.0 * B * H * H * Cin * 512, B * H * H * (Cin + 512 + 128) * 2)
    elif kind in ("sc64", "sc128"):
        H, Cin = (64, 64) if kind == "sc64" else (128, 64)
        Cout = 128
        x = torch.randn(B, H, H, Cin, device=dev).to(torch.bfloat16)
        w = (torch.randn(1, Cout, Cin, device=dev) / Cin ** 0.5).to(torch.bfloat16)
        tab = table(Cout, 1.0)
        out = torch.empty(B, H // 2, H // 2, Cout, device=dev, dtype=torch.bfloat16)
        fn = lambda: ops.conv_igemm(x, w, [(0, 0, 0)], out, tab, B=B, H=H, W=H, Cin=Cin, Cout=Cout, OH=H // 2, OW=H // 2, stride=2)
        fn()
        check(out, ref_conv(x[:2], w, 1, 2, 0, tab), kind)
        timeit(fn, kind, 2.0 * B * (H // 2) ** 2 * Cin * Cout, B * (H // 2) ** 2 * (Cin + Cout) * 2)
    else:
        raise SystemExit(f"unknown case {kind}")


for k in (["res64", "c64_256", "c128_64", "c256_32", "lat2", "lat1", "sc64", "sc128"] if which == "all" else [which]):
    run(k)
