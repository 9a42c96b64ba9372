#!/usr/bin/env python
"""Kernel micro-benchmarks (CUDA events, L2 flushed between iterations): achieved GB/s of
bias_act / upfirdn2d and TFLOP/s of the implicit-GEMM conv at the generator's layer shapes."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-fm-gan_b200"))
from fm3d import ops  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3


def main():
    res = []
    # ---- bias_act
    for shape, dt in [((32, 128, 256, 256), torch.float32), ((32, 128, 256, 256), torch.bfloat16),
                      ((32, 512, 64, 64), torch.float32)]:
        x = torch.randn(*shape, device=dev, dtype=dt)
        b = torch.randn(shape[1], device=dev, dtype=dt)
        t = timeit(lambda: ops.bias_act(x, b))
        gb = 2 * x.numel() * x.element_size() / t / 1e9
        res.append(dict(kernel="bias_act_fwd", shape=shape, dtype=str(dt), ms=t * 1e3, GBs=gb))
        t = timeit(lambda: ops.bias_act(x, None, x, 3, 1))
        res.append(dict(kernel="bias_act_grad", shape=shape, dtype=str(dt), ms=t * 1e3,
                        GBs=3 * x.numel() * x.element_size() / t / 1e9))
        t = timeit(lambda: ops.bias_act_grad_bias(x, x))
        res.append(dict(kernel="bias_act_grad_bias", shape=shape, dtype=str(dt), ms=t * 1e3,
                        GBs=3 * x.numel() * x.element_size() / t / 1e9))
        del x
    # ---- upfirdn2d
    k = torch.tensor([1., 3., 3., 1.]); k = (k[None] * k[:, None]); k = (k / k.sum() * 4).to(dev)
    for shape, cfg, dt in [((32, 128, 257, 257), (1, 1, 1, 1, 1, 1, 1, 1), torch.float32),
                           ((32, 128, 257, 257), (1, 1, 1, 1, 1, 1, 1, 1), torch.bfloat16),
                           ((32, 512, 65, 65), (1, 1, 1, 1, 1, 1, 1, 1), torch.float32),
                           ((32, 3, 128, 128), (2, 2, 1, 1, 2, 1, 2, 1), torch.float32)]:
        x = torch.randn(*shape, device=dev, dtype=dt)
        y = ops.upfirdn2d_planes(x, k, *cfg)
        t = timeit(lambda: ops.upfirdn2d_planes(x, k, *cfg))
        res.append(dict(kernel="upfirdn2d", shape=shape, cfg=cfg, dtype=str(dt), ms=t * 1e3,
                        GBs=(x.numel() + y.numel()) * x.element_size() / t / 1e9))
        del x, y
    # ---- igemm conv at the generator's shapes (B=32)
    B = 32
    for (H, Cin, Cout) in [(4, 512, 512), (8, 512, 512), (16, 512, 512), (32, 512, 512), (64, 512, 512),
                           (128, 256, 256), (256, 128, 128)]:
        x = torch.randn(B, H, H, Cin, device=dev).to(torch.bfloat16)
        w = (torch.randn(9, Cout, Cin, device=dev) / (Cin * 9) ** 0.5).to(torch.bfloat16)
        tab = torch.zeros(B, Cout, 8, device=dev); tab[..., 0] = 1; tab[..., 2] = 0.2; tab[..., 3] = 1
        out = torch.empty(B, H, H, Cout, device=dev, dtype=torch.bfloat16)
        noise = torch.randn(B, H, H, device=dev)
        nw = torch.ones(1, device=dev)
        fn = lambda: ops.conv_igemm(x, w, ops.conv_taps(3, 3, 1), out, tab, B=B, H=H, W=H, Cin=Cin, Cout=Cout, OH=H, OW=H,
                                    tab_per_sample=True, noise=noise, noise_w=nw)
        t = timeit(fn)
        fl = 2.0 * B * H * H * Cin * Cout * 9
        res.append(dict(kernel="igemm3x3", H=H, Cin=Cin, Cout=Cout, ms=t * 1e3, TFLOPs=fl / t / 1e12))
        del x, out
    for r in res:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
