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


def bench_synth(res):
    """blur_act_nhwc / rgb_finalize at the generator's shapes (B=32), fused up-conv vs 4 phase launches."""
    B = 32
    k = torch.tensor([1., 3., 3., 1.]); k = (k[None] * k[:, None]); k = (k / k.sum() * 4).to(dev)
    for (OH, C) in [(256, 128), (128, 256), (64, 512), (32, 512)]:
        t_in = torch.randn(B, OH + 1, OH + 1, C, device=dev).to(torch.bfloat16)
        tab = torch.zeros(B, C, 8, device=dev); tab[..., 0] = 1; tab[..., 2] = 0.2; tab[..., 3] = 1
        noise = torch.randn(B, 1, OH, OH, device=dev)
        nw = torch.ones(1, device=dev)
        out = torch.empty(B, OH, OH, C, device=dev, dtype=torch.bfloat16)
        t = timeit(lambda: ops.blur_act_nhwc(t_in, k, tab, noise, True, nw, C, out=out))
        res.append(dict(kernel="blur_act_nhwc", OH=OH, C=C, ms=t * 1e3, GBs=(t_in.numel() + out.numel()) * 2 / t / 1e9))
        del t_in, out
    for H in (256, 128, 64):
        acc = torch.randn(B, H, H, 4, device=dev)
        skip = torch.randn(B, 3, H // 2, H // 2, device=dev)
        bias = torch.zeros(3, device=dev)
        out = torch.empty(B, 3, H, H, device=dev)
        t = timeit(lambda: ops.rgb_finalize(acc, bias, skip, k, out=out))
        byts = acc.numel() * 4 * 2 + skip.numel() * 4 + out.numel() * 4
        res.append(dict(kernel="rgb_finalize", H=H, ms=t * 1e3, GBs=byts / t / 1e9))


def bench_upconv(res):
    """Stride-2 transposed 3x3 conv: per-image parity-phase launches vs the engine's tall-image launches."""
    import types
    sys.path.insert(0, os.path.join(ROOT, "3d-fm-gan_b200"))
    from fm3d import engine
    B = 32
    for (h, Cin, Cout) in [(128, 256, 128), (64, 512, 256), (32, 512, 512), (16, 512, 512), (8, 512, 512)]:
        x = torch.randn(B, h, h, Cin, device=dev).to(torch.bfloat16)
        w = (torch.randn(9, Cout, Cin, device=dev) / (Cin * 9) ** 0.5).to(torch.bfloat16)
        t_out = torch.empty(B, 2 * h + 1, 2 * h + 1, Cout, device=dev, dtype=torch.bfloat16)
        tw_ = 16 if h + 1 > 8 else 8
        th_ = 128 // tw_ if h + 1 > 8 else 8

        def phases():
            for py in (0, 1):
                for px in (0, 1):
                    ops.conv_igemm(x, w, engine._up_phase_taps(py, px), t_out, None, B=B, H=h, W=h, Cin=Cin, Cout=Cout,
                                   OH=h + 1 - py, OW=h + 1 - px, out_H=2 * h + 1, out_W=2 * h + 1, out_y0=py, out_x0=px,
                                   out_ys=2, out_xs=2, tab_per_sample=False, tile_w=tw_, tile_h=th_)

        xp = torch.zeros(B, h + 1, h + 1, Cin, device=dev, dtype=torch.bfloat16)
        xp[:, :h, :h] = x
        tp = torch.empty(B, 2 * h + 2, 2 * h + 2, Cout, device=dev, dtype=torch.bfloat16)
        L = types.SimpleNamespace(cin=Cin, cout=Cout, wq=w, wpair=None, wpair_all=None)
        if Cout <= 128:
            n0 = len(engine._PAIR_VIEWS[0])
            L.wpair_all = torch.zeros(n0 + len(engine._PAIR_VIEWS[1]), 2 * Cout, Cin, device=dev, dtype=torch.bfloat16)
            L.wpair = {0: L.wpair_all[:n0], 1: L.wpair_all[n0:]}
            for py, views in engine._PAIR_VIEWS.items():
                wp = L.wpair[py]
                for v, (_, t0, t1) in enumerate(views):
                    wp[v, :Cout] = w[t0]
                    if t1 is not None:
                        wp[v, Cout:] = w[t1]

        def tall():
            engine.SynthesisPlan._up_conv(None, L, xp, tp, B, h)
        fl = 2.0 * B * h * h * Cin * Cout * 9
        for name, fn in (("upconv_phases_per_image", phases), ("upconv_tall_image", tall)):
            t = timeit(fn)
            res.append(dict(kernel=name, h=h, Cin=Cin, Cout=Cout, ms=t * 1e3, TFLOPs=fl / t / 1e12))
        del x, t_out, xp, tp


def bench_train(res):
    """The three maps of a differentiable conv (fm3d/convgrad.py) at the generator's / discriminator's layer shapes, B = 32:
    forward and data gradient on fm_conv_igemm, weight gradient on fm_conv_wgrad (NHWC bf16 operands already in place:
    the layout passes from the NCHW fp32 autograd tensors are timed separately as 'nchw->nhwc')."""
    from fm3d import convgrad
    B = 32
    for (H, Cin, Cout, k, s) in [(64, 512, 512, 3, 1), (128, 256, 256, 3, 1), (256, 128, 128, 3, 1), (32, 512, 512, 3, 1),
                                 (257, 128, 256, 3, 2), (129, 256, 512, 3, 2), (256, 3, 128, 1, 1), (256, 128, 3, 1, 1)]:
        p = k // 2 if s == 1 else 0
        OH = (H + 2 * p - k) // s + 1
        cs = lambda c: (c + 7) // 8 * 8
        x = torch.zeros(B, H, H, cs(Cin), device=dev, dtype=torch.bfloat16)
        x[..., :Cin] = torch.randn(B, H, H, Cin, device=dev).to(torch.bfloat16)
        g = torch.zeros(B, OH, OH, cs(Cout), device=dev, dtype=torch.bfloat16)
        g[..., :Cout] = torch.randn(B, OH, OH, Cout, device=dev).to(torch.bfloat16)
        shifts = [(ky - p, kx - p) for ky in range(k) for kx in range(k)]
        fl = 2.0 * B * OH * OH * Cin * Cout * k * k
        if Cout >= Cin:
            fn = lambda: convgrad.conv_wgrad(g, x, Cout, Cin, B, OH, OH, [(0, 0, dy, dx) for (dy, dx) in shifts], 1, s)
        else:
            fn = lambda: convgrad.conv_wgrad(x, g, Cin, Cout, B, OH, OH, [(dy, dx, 0, 0) for (dy, dx) in shifts], s, 1)
        t = timeit(fn, iters=5, warm=2)
        res.append(dict(kernel="conv_wgrad", H=H, Cin=Cin, Cout=Cout, k=k, stride=s, ms=t * 1e3, TFLOPs=fl / t / 1e12))
        if H <= 128 or Cin <= 128:
            xf = torch.randn(B, Cin, H, H, device=dev)
            gf = torch.randn(B, Cout, OH, OH, device=dev)
            w = torch.randn(Cout, Cin, k, k, device=dev) / (Cin * k * k) ** 0.5
            t = timeit(lambda: convgrad.conv_forward(xf, w, s, p), iters=5, warm=2)
            res.append(dict(kernel="conv_forward (nchw fp32 in/out)", H=H, Cin=Cin, Cout=Cout, k=k, stride=s, ms=t * 1e3, TFLOPs=fl / t / 1e12))
            t = timeit(lambda: convgrad.conv_backward_data(gf, w, s, p, (H, H)), iters=5, warm=2)
            res.append(dict(kernel="conv_backward_data (nchw fp32 in/out)", H=H, Cin=Cin, Cout=Cout, k=k, stride=s, ms=t * 1e3, TFLOPs=fl / t / 1e12))
            t = timeit(lambda: convgrad.conv_backward_weight(xf, gf, s, p, k), iters=5, warm=2)
            res.append(dict(kernel="conv_backward_weight (nchw fp32 in)", H=H, Cin=Cin, Cout=Cout, k=k, stride=s, ms=t * 1e3, TFLOPs=fl / t / 1e12))
            t = timeit(lambda: ops.nchw_to_nhwc_bf16(xf), iters=5, warm=2)
            res.append(dict(kernel="nchw->nhwc bf16", H=H, C=Cin, ms=t * 1e3, GBs=xf.numel() * 6 / t / 1e9))
            del xf, gf
        del x, g


def main():
    res = []
    which = sys.argv[1:] or ["ops", "igemm", "synth", "upconv"]
    if "train" in which:
        bench_train(res)
    if "synth" in which:
        bench_synth(res)
    if "upconv" in which:
        bench_upconv(res)
    if "ops" in which:
        bench_ops(res)
    if "igemm" in which:
        bench_igemm(res)
    for r in res:
        print(json.dumps(r))


def bench_ops(res):
    # ---- bias_act
    for shape, dt in [((32, 128, 256, 256), torch.float32), ((32, 128, 256, 256), torch.bfloat16),
                      ((32, 512, 64, 64), torch.float32)]:
        x = torch.randn(*shape, device=dev, dtype=dt)
        b = torch.randn(shape[1], device=dev, dtype=dt)
        t = timeit(lambda: ops.bias_act(x, b))
        gb = 2 * x.numel() * x.element_size() / t / 1e9
        res.append(dict(kernel="bias_act_fwd", shape=shape, dtype=str(dt), ms=t * 1e3, GBs=gb))
        # gradient mode reads TWO tensors (upstream gradient and the saved forward output): a distinct ``ref`` so that
        # the 3*numel bytes credited are really moved (round 1 aliased ref = x and reported 1.4x the HBM peak)
        ref = torch.randn(*shape, device=dev, dtype=dt)
        t = timeit(lambda: ops.bias_act(x, None, ref, 3, 1))
        res.append(dict(kernel="bias_act_grad", shape=shape, dtype=str(dt), ms=t * 1e3,
                        GBs=3 * x.numel() * x.element_size() / t / 1e9))
        t = timeit(lambda: ops.bias_act_grad_bias(x, ref))
        res.append(dict(kernel="bias_act_grad_bias", shape=shape, dtype=str(dt), ms=t * 1e3,
                        GBs=3 * x.numel() * x.element_size() / t / 1e9))
        del x, ref
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


def bench_igemm(res):
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


if __name__ == "__main__":
    main()
