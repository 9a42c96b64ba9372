#!/usr/bin/env python
"""BASELINE config 3: kernel sweep over every synthesis resolution 4^2..256^2, channel widths of
channel_multiplier 2 and 1 (512 -> 64), batch 1..64.

Per case (CUDA events, median of 7, L2 flushed between iterations, one JSON line each):
  modconv      plain modulated 3x3 conv as the engine runs it (per-sample demod table, noise, bias, lrelu, fused ToRGB sums)
  modconv_up   stride-2 transposed modulated conv: tall-image parity phases + blur_act (demod/noise/bias/lrelu)
  upfirdn2d    the op/ API kernel, fp32 NCHW: Blur pad (1,1) on [B,C,2r+1,2r+1] and ToRGB Upsample (up 2, pad (2,1)) of [B,3,r,r]
  bias_act     fused_bias_act forward, fp32 and bf16, [B,C,r,r]
TFLOP/s use the algorithmic FLOPs of SURVEY 8(d) (transposed conv: 9 taps at the INPUT resolution);
GB/s use algorithmic bytes (input once + output once).
usage: kernel_sweep.py [--batches 1,4,16,32,64] [--quick]"""
import argparse
import json
import os
import sys
import types

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-fm-gan_b200"))
from fm3d import engine, ops  # noqa: E402

dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False            # the checks below are strict fp32
torch.backends.cuda.matmul.allow_tf32 = False
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=7, warm=2):
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


def channels(cm):
    return {4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * cm, 128: 128 * cm, 256: 64 * cm}


def emit(**kw):
    print(json.dumps(kw), flush=True)


def _sel(cout):
    """A few output channels from every 64-wide n-tile: the fp32 check costs 1/16 of the conv."""
    s = set()
    for n0 in range(0, cout, 64):
        hi = min(n0 + 64, cout) - 1
        s.update((n0, (n0 + hi) // 2, hi))
    return torch.tensor(sorted(s), device=dev)


def check(name, got, ref, tol):
    """Every shape is verified ONCE before it is timed: a speed without a correct result is not a measurement."""
    err = float((got.float() - ref.float()).abs().max() / ref.float().abs().max().clamp_min(1e-12))
    if not (err < tol):
        raise SystemExit(f"kernel_sweep: {name}: result differs from the fp32 reference (rel err {err:.3e} >= {tol})")
    return err


def sweep_modconv(B, r, Cin, Cout, cm):
    cs = lambda c: (c + 7) // 8 * 8               # physical channel stride (pruned widths are not multiples of 8)
    x = torch.zeros(B, r, r, cs(Cin), device=dev, dtype=torch.bfloat16)
    x[..., :Cin] = torch.randn(B, r, r, Cin, device=dev).to(torch.bfloat16)
    w = torch.zeros(9, Cout, cs(Cin), device=dev, dtype=torch.bfloat16)
    w[..., :Cin] = (torch.randn(9, Cout, Cin, device=dev) / (Cin * 9) ** 0.5).to(torch.bfloat16)
    tab = torch.zeros(B, Cout, 8, device=dev); tab[..., 0] = 1; tab[..., 2] = 0.2; tab[..., 3] = 1.4; tab[..., 4:7] = 0.01
    out = torch.empty(B, r, r, cs(Cout), device=dev, dtype=torch.bfloat16)
    noise = torch.randn(B, r, r, device=dev)
    nw = torch.ones(1, device=dev)
    rgb = torch.zeros(B, r, r, 4, device=dev)
    run = lambda: ops.conv_igemm(x, w, ops.conv_taps(3, 3, 1), out, tab, B=B, H=r, W=r, Cin=Cin, Cout=Cout, OH=r, OW=r,
                                 tab_per_sample=True, noise=noise, noise_w=nw, rgb=rgb)
    # ---- correctness first: fp32 conv (TF32 off) of the same bf16 operands on a channel sample + the whole epilogue
    run()
    sel = _sel(Cout)
    wf = w[:, sel, :Cin].float().permute(1, 2, 0).reshape(len(sel), Cin, 3, 3)
    nb = min(B, 4)                                                   # first and last samples
    bs = torch.tensor(sorted(set(list(range(nb)) + list(range(B - nb, B)))), device=dev)
    acc = F.conv2d(x[bs][..., :Cin].float().permute(0, 3, 1, 2), wf, padding=1)
    v = acc * tab[bs][:, sel, 0, None, None] + tab[bs][:, sel, 1, None, None] + noise[bs][:, None]
    v = torch.where(v > 0, v, v * tab[bs][:, sel, 2, None, None]) * tab[bs][:, sel, 3, None, None]
    err = check(f"modconv B={B} r={r} {Cin}->{Cout}", out[bs][..., sel].permute(0, 3, 1, 2), v, 2e-2)
    rgb.zero_()
    t = timeit(run)
    fl = 2.0 * B * r * r * Cin * Cout * 9
    emit(kernel="modconv3x3+torgb", cm=cm, B=B, res=r, Cin=Cin, Cout=Cout, us=t * 1e6, TFLOPs=fl / t / 1e12, checked_rel_err=err)


def sweep_modconv_up(B, h, Cin, Cout, cm):
    """h -> 2h."""
    xp = torch.zeros(B, h + 1, h + 1, Cin, device=dev, dtype=torch.bfloat16)
    xp[:, :h, :h] = torch.randn(B, h, h, Cin, device=dev).to(torch.bfloat16)
    w = (torch.randn(9, Cout, Cin, device=dev) / (Cin * 9) ** 0.5).to(torch.bfloat16)
    tp = torch.empty(B, 2 * h + 2, 2 * h + 2, Cout, device=dev, dtype=torch.bfloat16)
    y = torch.empty(B, 2 * h, 2 * h, Cout, device=dev, dtype=torch.bfloat16)
    L = types.SimpleNamespace(cin=Cin, cout=Cout, wq=w, wpair=None, wpair_all=None)
    if Cout <= 128 and Cout % 32 == 0 and h >= 12:
        n0 = len(engine._PAIR_VIEWS[0])
        L.wpair_all = torch.zeros(n0 + len(engine._PAIR_VIEWS[1]), 2 * Cout, Cin, device=dev, dtype=torch.bfloat16)
        L.wpair = {0: L.wpair_all[:n0], 1: L.wpair_all[n0:]}
        for py, views in engine._PAIR_VIEWS.items():
            wp = L.wpair[py]
            for v, (_, t0, t1) in enumerate(views):
                wp[v, :Cout] = w[t0]
                if t1 is not None:
                    wp[v, Cout:] = w[t1]
    k = torch.tensor([1., 3., 3., 1.]); k = (k[None] * k[:, None]); k = (k / k.sum() * 4).to(dev)
    tab = torch.zeros(B, Cout, 8, device=dev); tab[..., 0] = 1; tab[..., 2] = 0.2; tab[..., 3] = 1.4
    noise = torch.randn(B, 1, 2 * h, 2 * h, device=dev)
    nw = torch.ones(1, device=dev)
    fl = 2.0 * B * h * h * Cin * Cout * 9
    # ---- correctness first: transposed conv vs F.conv_transpose2d, blur + epilogue vs the torch composition
    engine.SynthesisPlan._up_conv(None, L, xp, tp, B, h)
    ops.blur_act_nhwc(tp, k, tab, noise, True, nw, Cout, out=y, padded=True)
    sel = _sel(Cout)
    nb = min(B, 2)
    wf = w[:, sel].float().permute(1, 2, 0).reshape(len(sel), Cin, 3, 3)
    tref = F.conv_transpose2d(xp[:nb, :h, :h].float().permute(0, 3, 1, 2), wf.transpose(0, 1), stride=2)
    e1 = check(f"up-conv B={B} h={h} {Cin}->{Cout}", tp[:nb, :2 * h + 1, :2 * h + 1][..., sel].permute(0, 3, 1, 2), tref, 2e-2)
    tin = tp[:nb, :2 * h + 1, :2 * h + 1][..., sel].float().permute(0, 3, 1, 2)
    kk = torch.flip(k, [0, 1])[None, None].expand(len(sel), 1, 4, 4)
    bl = F.conv2d(F.pad(tin, (1, 1, 1, 1)), kk, groups=len(sel))
    v = bl * tab[:nb, sel, 0, None, None] + tab[:nb, sel, 1, None, None] + noise[:nb]
    v = torch.where(v > 0, v, v * 0.2) * 1.4
    e2 = check(f"blur_act B={B} res={2 * h} C={Cout}", y[:nb][..., sel].permute(0, 3, 1, 2), v, 2e-2)
    t_conv = timeit(lambda: engine.SynthesisPlan._up_conv(None, L, xp, tp, B, h))
    t_blur = timeit(lambda: ops.blur_act_nhwc(tp, k, tab, noise, True, nw, Cout, out=y, padded=True))
    emit(kernel="modconv_up(phases)", cm=cm, B=B, res_in=h, Cin=Cin, Cout=Cout, us=t_conv * 1e6, TFLOPs=fl / t_conv / 1e12,
         checked_rel_err=e1)
    emit(kernel="blur_act(after up)", cm=cm, B=B, res_out=2 * h, C=Cout, us=t_blur * 1e6,
         GBs=(B * (2 * h + 1) ** 2 + B * 4 * h * h) * Cout * 2 / t_blur / 1e9, checked_rel_err=e2)


def sweep_ops(B, r, C):
    k = torch.tensor([1., 3., 3., 1.]); k = (k[None] * k[:, None]); k = (k / k.sum() * 4).to(dev)
    # Blur after the up-conv: [B,C,r+1,r+1] -> [B,C,r,r]  (r = output resolution of the layer)
    x = torch.randn(B, C, r + 1, r + 1, device=dev)
    yv = ops.upfirdn2d_planes(x, k, 1, 1, 1, 1, 1, 1, 1, 1)
    n1 = min(B, 2)
    check(f"upfirdn2d blur B={B} r={r} C={C}", yv[:n1],
          F.conv2d(F.pad(x[:n1], (1, 1, 1, 1)), torch.flip(k, [0, 1])[None, None].expand(C, 1, 4, 4), groups=C), 1e-5)
    t = timeit(lambda: ops.upfirdn2d_planes(x, k, 1, 1, 1, 1, 1, 1, 1, 1))
    emit(kernel="upfirdn2d blur pad(1,1) fp32", B=B, res=r, C=C, us=t * 1e6, GBs=(x.numel() + yv.numel()) * 4 / t / 1e9)
    del x, yv
    x = torch.randn(B, 3, r // 2, r // 2, device=dev)
    yv = ops.upfirdn2d_planes(x, k, 2, 2, 1, 1, 2, 1, 2, 1)
    xs = torch.zeros(B, 3, r, r, device=dev); xs[:, :, ::2, ::2] = x          # zero-stuffing, pad (2,1), flipped FIR
    check(f"upfirdn2d up2 B={B} r={r}", yv,
          F.conv2d(F.pad(xs, (2, 1, 2, 1)), torch.flip(k, [0, 1])[None, None].expand(3, 1, 4, 4), groups=3), 1e-5)
    t = timeit(lambda: ops.upfirdn2d_planes(x, k, 2, 2, 1, 1, 2, 1, 2, 1))
    emit(kernel="upfirdn2d up2 pad(2,1) fp32 (ToRGB skip)", B=B, res=r, C=3, us=t * 1e6, GBs=(x.numel() + yv.numel()) * 4 / t / 1e9)
    del x, yv
    for dt in (torch.float32, torch.bfloat16):
        x = torch.randn(B, C, r, r, device=dev, dtype=dt)
        b = torch.randn(C, device=dev, dtype=dt)
        check(f"bias_act {dt} B={B} r={r} C={C}", ops.bias_act(x[:2], b),
              F.leaky_relu(x[:2].float() + b.float()[None, :, None, None], 0.2) * 2 ** 0.5, 1e-5 if dt == torch.float32 else 1e-2)
        t = timeit(lambda: ops.bias_act(x, b))
        emit(kernel="fused_bias_act fwd", dtype=str(dt).split(".")[1], B=B, res=r, C=C, us=t * 1e6, GBs=2 * x.numel() * x.element_size() / t / 1e9)
        ref = torch.randn_like(x)
        t = timeit(lambda: ops.bias_act(x, None, ref, 3, 1))
        emit(kernel="fused_bias_act grad", dtype=str(dt).split(".")[1], B=B, res=r, C=C, us=t * 1e6, GBs=3 * x.numel() * x.element_size() / t / 1e9)
        del x, ref


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="1,4,16,32,64")
    ap.add_argument("--quick", action="store_true", help="channel_multiplier 2 only")
    args = ap.parse_args()
    batches = [int(b) for b in args.batches.split(",")]
    for cm in ((2,) if args.quick else (2, 1)):
        ch = channels(cm)
        for B in batches:
            for r in (4, 8, 16, 32, 64, 128, 256):
                if cm == 1 and r < 64:
                    continue                      # identical to cm = 2 below 64^2
                sweep_modconv(B, r, ch[r], ch[r], cm)
                if r > 4:
                    sweep_modconv_up(B, r // 2, ch[r // 2], ch[r], cm)
                sweep_ops(B, r, ch[r])
    # pruned (odd) widths the reference's network-slimming path produces (Util/network_util.py:87-95)
    for (r, Cin, Cout) in ((64, 370, 200), (32, 200, 370)):
        sweep_modconv(32, r, Cin, Cout, 0)


if __name__ == "__main__":
    main()
