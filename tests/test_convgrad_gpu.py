"""Differentiable convolutions on the tcgen05 kernels (fm3d/convgrad.py: fm_conv_igemm for forward and dgrad,
fm_conv_wgrad for the weight gradient) against ATen autograd in strict fp32: forward, first-order gradients and the
second-order passes R1 / path-length regularisation need.  Operands are rounded to bf16, so the tolerance is 2e-2 of
each tensor's max magnitude (measured ~4e-3)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _strict_fp32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))


def _inputs(cuda, B, I, O, H, k, seed, transpose=False):
    g = torch.Generator(device=cuda).manual_seed(seed)
    x = torch.randn(B, I, H, H, generator=g, device=cuda)
    wshape = (I, O, k, k) if transpose else (O, I, k, k)
    w = torch.randn(*wshape, generator=g, device=cuda) / (I * k * k) ** 0.5
    return x, w, g


CONV_CASES = [
    # B, I, O, H, k, stride, pad
    (2, 64, 64, 16, 3, 1, 1),
    (2, 96, 40, 9, 3, 1, 1),          # ragged channels, odd size
    (2, 64, 128, 17, 3, 2, 0),        # discriminator conv2: blur pad(2,2) then 3x3 stride 2 (stylegan2.py:705-716)
    (2, 64, 128, 15, 1, 2, 0),        # discriminator skip: blur pad(1,1) then 1x1 stride 2 (:748-750)
    (3, 512, 512, 4, 3, 1, 1),        # 4x4 layers (conv1)
    (2, 513, 512, 4, 3, 1, 1),        # final_conv after the minibatch-stddev channel (:816)
    (2, 3, 64, 32, 1, 1, 0),          # discriminator stem (from RGB)
    (2, 128, 3, 32, 1, 1, 0),         # ToRGB
    (4, 128, 128, 64, 3, 1, 1),
    (2, 64, 64, 32, 3, 2, 1),         # encoders' stride-2 3x3 pad 1
    (2, 256, 256, 33, 3, 1, 1),       # K loop over several chunks, 33-wide rows
]


@pytest.mark.parametrize("B,I,O,H,k,s,p", CONV_CASES)
def test_conv2d_forward_backward(cuda, B, I, O, H, k, s, p):
    from fm3d import convgrad
    x, w, g = _inputs(cuda, B, I, O, H, k, seed=H * 7 + I)
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    xn, wn = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, stride=s, padding=p)
    yn = convgrad.conv2d(xn, wn, stride=s, padding=p)
    assert yn.shape == yr.shape
    gy = torch.randn(yr.shape, generator=g, device=cuda)
    gxr, gwr = torch.autograd.grad(yr, [xr, wr], gy)
    gxn, gwn = torch.autograd.grad(yn, [xn, wn], gy)
    e = (_rel(yn, yr), _rel(gxn, gxr), _rel(gwn, gwr))
    print(f"conv2d B{B} {I}->{O} {H}^2 k{k} s{s} p{p}: fwd {e[0]:.4f} dgrad {e[1]:.4f} wgrad {e[2]:.4f}")
    assert max(e) < 2e-2, e


@pytest.mark.parametrize("B,I,O,H,k,s", [(2, 64, 48, 8, 3, 2), (3, 512, 512, 4, 3, 2), (2, 128, 64, 32, 3, 2), (2, 256, 128, 17, 3, 2)])
def test_conv_transpose2d_forward_backward(cuda, B, I, O, H, k, s):
    """The generator's up-conv (stylegan2.py:276): stride-2 transposed 3x3, h -> 2h+1."""
    from fm3d import convgrad
    x, w, g = _inputs(cuda, B, I, O, H, k, seed=H + O, transpose=True)
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    xn, wn = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    yr = F.conv_transpose2d(xr, wr, stride=s)
    yn = convgrad.conv_transpose2d(xn, wn, stride=s)
    assert yn.shape == yr.shape == (B, O, (H - 1) * s + k, (H - 1) * s + k)
    gy = torch.randn(yr.shape, generator=g, device=cuda)
    gxr, gwr = torch.autograd.grad(yr, [xr, wr], gy)
    gxn, gwn = torch.autograd.grad(yn, [xn, wn], gy)
    e = (_rel(yn, yr), _rel(gxn, gxr), _rel(gwn, gwr))
    print(f"conv_transpose2d B{B} {I}->{O} {H}^2: fwd {e[0]:.4f} dgrad {e[1]:.4f} wgrad {e[2]:.4f}")
    assert max(e) < 2e-2, e


@pytest.mark.parametrize("kind,B,I,O,H,k,s,p", [("conv", 2, 32, 48, 12, 3, 1, 1), ("conv", 2, 32, 64, 13, 3, 2, 0),
                                                ("conv", 2, 16, 32, 11, 1, 2, 0), ("convT", 2, 48, 32, 6, 3, 2, 0)])
def test_double_backward(cuda, kind, B, I, O, H, k, s, p):
    """R1 / path-length pattern: differentiate a function of the first-order gradient w.r.t. the weight and the input.
    Every second-order term is again ConvFwd / ConvBwdData / ConvBwdWeight."""
    from fm3d import convgrad
    x, w, g = _inputs(cuda, B, I, O, H, k, seed=H * 3 + O, transpose=(kind == "convT"))

    def run(native):
        xv, wv = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
        if kind == "conv":
            y = convgrad.conv2d(xv, wv, stride=s, padding=p) if native else F.conv2d(xv, wv, stride=s, padding=p)
        else:
            y = convgrad.conv_transpose2d(xv, wv, stride=s) if native else F.conv_transpose2d(xv, wv, stride=s)
        probe = torch.randn(y.shape, generator=torch.Generator(device=cuda).manual_seed(1), device=cuda)
        gx, gw = torch.autograd.grad((y.tanh() * probe).sum(), [xv, wv], create_graph=True)
        loss = gx.pow(2).sum() + gw.pow(2).sum()
        ggx, ggw = torch.autograd.grad(loss, [xv, wv])
        return y.detach(), gx.detach(), gw.detach(), ggx, ggw
    ref, got = run(False), run(True)
    errs = [_rel(a, b) for a, b in zip(got, ref)]
    print(f"{kind} double backward: " + " ".join(f"{e:.4f}" for e in errs))
    assert max(errs[:3]) < 2e-2 and max(errs[3:]) < 4e-2, errs


def test_conv_wgrad_direct(cuda):
    """fm_conv_wgrad on hand-made NHWC operands: ragged channel counts (3, 40, 136), a shifted / strided operand on either
    side, grids that do not fill the 64-pixel K chunk, forced split-K.  Contract: one operand IS the grid (H = GH, W = GW,
    stride 1, no shift) -- its out-of-bounds zero fill masks the tile pixels beyond the grid."""
    from fm3d import convgrad
    g = torch.Generator(device=cuda).manual_seed(3)
    B, GH, GW = 3, 6, 10
    cs = lambda c: (c + 7) // 8 * 8
    for (Cg, Cs, s, swap) in ((136, 40, 1, False), (64, 3, 2, False), (200, 264, 2, True), (24, 130, 1, True)):
        Hs, Ws = GH * s + 2, GW * s + 1
        grid = torch.zeros(B, GH, GW, cs(Cg), device=cuda, dtype=torch.bfloat16)
        other = torch.zeros(B, Hs, Ws, cs(Cs), device=cuda, dtype=torch.bfloat16)
        grid[..., :Cg] = torch.randn(B, GH, GW, Cg, generator=g, device=cuda).to(torch.bfloat16)
        other[..., :Cs] = torch.randn(B, Hs, Ws, Cs, generator=g, device=cuda).to(torch.bfloat16)
        shifts = [(0, 0), (1, 2), (-1, -1), (2, 0)]
        Gm = grid[..., :Cg].float().reshape(-1, Cg)
        for ksplit in (0, 1, 3):
            if swap:      # a = shifted / strided operand, b = grid
                dw = convgrad.conv_wgrad(other, grid, Cs, Cg, B, GH, GW, [(dy, dx, 0, 0) for (dy, dx) in shifts], s, 1, ksplit=ksplit)
            else:
                dw = convgrad.conv_wgrad(grid, other, Cg, Cs, B, GH, GW, [(0, 0, dy, dx) for (dy, dx) in shifts], 1, s, ksplit=ksplit)
            for t, (dy, dx) in enumerate(shifts):
                S = torch.zeros(B, GH, GW, Cs, device=cuda)
                for gy in range(GH):
                    for gx in range(GW):
                        y, x = gy * s + dy, gx * s + dx
                        if 0 <= y < Hs and 0 <= x < Ws:
                            S[:, gy, gx] = other[:, y, x, :Cs].float()
                S = S.reshape(-1, Cs)
                ref = S.t() @ Gm if swap else Gm.t() @ S
                assert _rel(dw[t], ref) < 2e-3, (Cg, Cs, s, swap, ksplit, t, _rel(dw[t], ref))


@pytest.mark.parametrize("shape", [(3, 8, 16, 16), (2, 5, 7, 9), (4, 64, 32, 32), (2, 3, 1, 1)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_channel_scale_and_dot_kernels(cuda, shape, dtype):
    """fm_channel_scale / fm_channel_dot (vector and scalar paths) against the torch broadcast forms."""
    from fm3d import ops
    g = torch.Generator(device=cuda).manual_seed(sum(shape))
    x = torch.randn(*shape, generator=g, device=cuda).to(dtype)
    y = torch.randn(*shape, generator=g, device=cuda).to(dtype)
    s = torch.randn(shape[0], shape[1], generator=g, device=cuda).to(dtype)
    ref = (x.float() * s.float()[:, :, None, None]).to(dtype)
    assert torch.equal(ops.channel_scale(x, s), ref)
    dot = ops.channel_dot(x, y)
    dref = (x.float() * y.float()).sum(dim=(2, 3))
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    torch.testing.assert_close(dot.float(), dref, rtol=tol, atol=tol * max(1.0, float(dref.abs().max())))


def test_channel_scale_autograd_to_second_order(cuda):
    """ChannelScale / ChannelDot (each the other's gradient) against autograd of the broadcast multiply: first-order
    gradients and a double backward of the kind path-length regularisation takes through ModulatedConv2d."""
    from fm3d import convgrad
    g = torch.Generator(device=cuda).manual_seed(5)
    B, C, H = 3, 12, 10
    x0 = torch.randn(B, C, H, H, generator=g, device=cuda)
    s0 = torch.randn(B, C, generator=g, device=cuda)
    w = torch.randn(B, C, H, H, generator=g, device=cuda)
    res = []
    for native in (True, False):
        x, s = x0.clone().requires_grad_(True), s0.clone().requires_grad_(True)
        y = convgrad.channel_scale(x, s) if native else x * s.view(B, C, 1, 1)
        loss = (y * w).sum() + y.pow(2).sum()
        gx, gs = torch.autograd.grad(loss, [x, s], create_graph=True)
        pen = gx.pow(2).sum() + gs.pow(2).sum()
        ggx, ggs = torch.autograd.grad(pen, [x, s])
        res.append((y.detach(), gx.detach(), gs.detach(), ggx, ggs))
    for a, b in zip(*res):
        assert _rel(a, b) < 1e-5


def test_conv2d_bias_add_autograd(cuda):
    """convgrad.conv2d with a bias (BiasAdd on the fused bias-act kernel): output, gradients w.r.t. input / weight / bias
    and a double backward through the input gradient, against ATen."""
    from fm3d import convgrad
    x0, w0, g = _inputs(cuda, 2, 24, 40, 12, 3, 11)
    b0 = torch.randn(40, generator=g, device=cuda)
    res = []
    for native in (True, False):
        x, w, b = (t.clone().requires_grad_(True) for t in (x0, w0, b0))
        xq, wq = (x, w) if native else (x.to(torch.bfloat16).float(), w.to(torch.bfloat16).float())
        y = convgrad.conv2d(xq, wq, b, 1, 1) if native else F.conv2d(xq, wq, b, 1, 1)
        gx, gw, gb = torch.autograd.grad(y.pow(2).sum(), [x, w, b], create_graph=True)
        ggb, = torch.autograd.grad(gx.pow(2).sum(), [b])
        res.append((y.detach(), gx.detach(), gw.detach(), gb.detach(), ggb))
    for a, r in zip(*res):
        assert _rel(a, r) < 2e-2


@pytest.mark.parametrize("shape,shared", [((3, 8, 16, 16), False), ((3, 8, 16, 16), True), ((2, 5, 7, 9), False), ((1, 4, 32, 32), False),
                                          ((2, 3, 5, 5), True)])
def test_plane_add_and_noise_injection(cuda, shape, shared):
    """fm_plane_add (vector and scalar paths, per-sample and shared planes) and NoiseInjection on it: output, gradients
    w.r.t. the image and the noise weight, and a double backward, against the broadcast form."""
    import stylegan2
    from fm3d import ops
    g = torch.Generator(device=cuda).manual_seed(sum(shape))
    B, C, H, W = shape
    x0 = torch.randn(*shape, generator=g, device=cuda)
    n = torch.randn(1 if shared else B, 1, H, W, generator=g, device=cuda)
    assert torch.equal(ops.plane_add(x0, n), x0 + n)
    inj = stylegan2.NoiseInjection().to(cuda)
    with torch.no_grad():
        inj.weight.fill_(0.37)
    res = []
    for native in (True, False):
        x = x0.clone().requires_grad_(True)
        y = inj(x, noise=n) if native else x + inj.weight * n
        gx, gw = torch.autograd.grad(y.pow(3).sum(), [x, inj.weight], create_graph=True)
        ggx, ggw = torch.autograd.grad(gx.pow(2).sum() + gw.pow(2).sum(), [x, inj.weight])
        res.append((y.detach(), gx.detach(), gw.detach(), ggx, ggw))
    for a, r in zip(*res):
        assert _rel(a, r) < 1e-5
