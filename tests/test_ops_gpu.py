"""Parity of the CUDA ops (through the C ABI) with the oracle and the golden vectors."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu

from oracle import fm_oracle as orc  # noqa: E402  (test infrastructure)


def _t(a, dev, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev).to(dtype)


# ----------------------------------------------------------------------------- upfirdn2d
def test_upfirdn2d_golden(cuda):
    from fm3d import ops
    g = load_golden("upfirdn2d.npz")
    for name in g["names"]:
        x, k, cfg, y = g[f"{name}.x"], g[f"{name}.k"], g[f"{name}.cfg"], g[f"{name}.y"]
        out = ops.upfirdn2d_planes(_t(x, cuda), _t(k, cuda), *[int(c) for c in cfg])
        assert tuple(out.shape) == y.shape, name
        # fp32 tolerance: same taps, different summation order / FMA contraction
        np.testing.assert_allclose(out.cpu().numpy(), y, rtol=1e-5, atol=1e-5, err_msg=str(name))


def test_upfirdn2d_api_and_analytic(cuda):
    import op
    g = load_golden("upfirdn2d.npz")
    out = op.upfirdn2d(_t(g["api.x"], cuda), _t(g["api.k"], cuda), up=2, down=1, pad=(2, 1))
    np.testing.assert_allclose(out.cpu().numpy(), g["api.y"], rtol=1e-5, atol=1e-5)
    # analytic KATs: DC gain of the x4 kernel
    k = orc.make_kernel_ref([1, 3, 3, 1]).to(cuda) * 4
    up = op.upfirdn2d(torch.ones(1, 1, 8, 8, device=cuda), k, up=2, pad=(2, 1))
    assert up.shape == (1, 1, 16, 16)
    assert torch.allclose(up[0, 0, 2:-2, 2:-2], torch.ones(12, 12, device=cuda), atol=1e-6)
    assert abs(float(up[0, 0, 0, 0]) - 0.5625) < 1e-6
    bl = op.upfirdn2d(torch.ones(1, 1, 17, 17, device=cuda), k, pad=(1, 1))
    assert bl.shape == (1, 1, 16, 16)
    assert torch.allclose(bl[0, 0, 2:-2, 2:-2], torch.full((12, 12), 4.0, device=cuda), atol=1e-5)


@pytest.mark.parametrize("shape,cfg", [
    ((4, 128, 257, 257), (1, 1, 1, 1, 1, 1, 1, 1)),     # largest G blur (per-sample slice)
    ((8, 3, 128, 128), (2, 2, 1, 1, 2, 1, 2, 1)),       # ToRGB skip upsample
    ((2, 3, 256, 256), (1, 1, 2, 2, 1, 1, 1, 1)),       # bwd of Upsample
    ((2, 7, 33, 130), (1, 1, 1, 1, 2, 2, 2, 2)),        # D blur, ragged tile edge
    ((1, 1, 1, 1), (1, 1, 1, 1, 2, 2, 2, 2)),           # minimum size
    ((2, 2, 5, 300), (2, 2, 1, 1, 2, 1, 2, 1)),
])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_upfirdn2d_vs_oracle(cuda, shape, cfg, dtype):
    from fm3d import ops
    gen = torch.Generator().manual_seed(sum(shape) + cfg[0])
    x = torch.randn(*shape, generator=gen).to(dtype)
    k = orc.make_kernel_ref([1, 3, 3, 1]) * (4 if cfg[0] == 2 or cfg[4] == 1 else 1)
    ref = orc.upfirdn2d_ref(x.float(), k, *cfg)
    out = ops.upfirdn2d_planes(x.to(cuda), k.to(cuda), *cfg)
    assert out.dtype == dtype and tuple(out.shape) == tuple(ref.shape)
    tol = 1e-5 if dtype == torch.float32 else (2e-2 if dtype == torch.bfloat16 else 3e-3)
    torch.testing.assert_close(out.float().cpu(), ref, rtol=tol, atol=tol * 4)


@pytest.mark.parametrize("shape,kshape,pads", [
    ((3, 5, 17, 17), (4, 4), (1, 1, 1, 1)),        # several whole planes per work item
    ((2, 2, 300, 1000), (3, 3), (2, 0, 1, 3)),     # wide rows: few rows per strip, asymmetric pads
    ((1, 3, 131, 67), (2, 4), (0, 3, 2, 1)),       # non-square kernel, odd sizes
    ((5, 1, 4, 4), (1, 1), (0, 0, 0, 0)),          # 1x1 kernel (identity scale)
    ((2, 1, 6, 9), (4, 4), (3, 3, 3, 3)),          # pads as large as the kernel allows
])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_upfirdn2d_stream_cases(cuda, shape, kshape, pads, dtype):
    """up = down = 1 takes the bulk-copy streaming kernel: non-separable random taps, ragged sizes and a
    source tensor that is only element-aligned (the copy widens its byte range to 16-byte granules)."""
    from fm3d import ops
    gen = torch.Generator().manual_seed(sum(shape) * 3 + kshape[0])
    numel = 1
    for d in shape:
        numel *= d
    base = torch.randn(numel + 3, generator=gen).to(dtype)
    k = torch.randn(*kshape, generator=gen)
    cfg = (1, 1, 1, 1) + pads
    for off in (0, 1, 3):
        x = base[off:off + numel].view(*shape)
        ref = orc.upfirdn2d_ref(x.float(), k, *cfg)
        xd = base.to(cuda)[off:off + numel].view(*shape)
        out = ops.upfirdn2d_planes(xd, k.to(cuda), *cfg)
        assert tuple(out.shape) == tuple(ref.shape)
        tol = 1e-5 if dtype == torch.float32 else 2e-2
        torch.testing.assert_close(out.float().cpu(), ref, rtol=tol, atol=tol * 8)


@pytest.mark.parametrize("shape,pads", [
    ((37, 3, 65, 65), (1, 1, 1, 1)),          # G blur after the 32->64 up-conv: whole planes between zero rows, 2-3 per item
    ((5, 7, 129, 129), (1, 1, 1, 1)),         # one plane per item in fp32, two in the 2-byte types
    ((3, 4, 64, 64), (2, 2, 2, 2)),           # D blur before the stride-2 conv (out 65 wide: odd output pitch)
    ((2, 9, 40, 51), (2, 1, 2, 1)),           # asymmetric pads, ragged last column pair
    ((2, 3, 48, 33), (0, 3, 3, 0)),           # out_w = in_w, pads as far as the zero rows reach
    ((1, 2, 200, 257), (1, 1, 1, 1)),         # strip mode, odd width
    ((1, 2, 130, 300), (2, 2, 2, 2)),         # strip mode, even width, odd output pitch
])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("sep", [True, False])
def test_upfirdn2d_stream_zero_row_layouts(cuda, shape, pads, dtype, sep):
    """The streaming kernel's zero-row layouts (strip mode and whole planes copied one by one between rows of zeros),
    the warp-uniform choice of the edge / interior walk with rotated column groups, and the packed column-pair walk of
    the 2-byte types (rank-1 taps) against the oracle, from element-aligned bases so that every parity of the staged
    address occurs."""
    from fm3d import ops
    gen = torch.Generator().manual_seed(sum(shape) * 7 + pads[0])
    numel = 1
    for d in shape:
        numel *= d
    base = torch.randn(numel + 5, generator=gen).to(dtype)
    k = orc.make_kernel_ref([1, 3, 3, 1]) * 4 if sep else torch.randn(4, 4, generator=gen)
    cfg = (1, 1, 1, 1) + pads
    based = base.to(cuda)
    for off in (0, 1, 2, 5):
        x = base[off:off + numel].view(*shape)
        ref = orc.upfirdn2d_ref(x.float(), k, *cfg)
        out = ops.upfirdn2d_planes(based[off:off + numel].view(*shape), k.to(cuda), *cfg)
        assert tuple(out.shape) == tuple(ref.shape) and out.dtype == dtype
        tol = 1e-5 if dtype == torch.float32 else (2e-2 if dtype == torch.bfloat16 else 3e-3)
        torch.testing.assert_close(out.float().cpu(), ref, rtol=tol, atol=tol * 8)


@pytest.mark.parametrize("shape,kshape,pads", [
    ((4, 3, 64, 64), (4, 4), (2, 1, 2, 1)),       # ToRGB skip upsample (stylegan2.py:65-80)
    ((2, 3, 9, 13), (4, 4), (2, 1, 2, 1)),        # output width not a multiple of 4, odd heights
    ((3, 2, 6, 5), (3, 4), (1, 2, 0, 3)),         # every parity of the two leading pads, non-square taps
    ((1, 5, 4, 9), (4, 2), (3, 0, 1, 1)),
    ((2, 2, 6, 6), (4, 4), (-1, 2, 2, -1)),       # negative pads crop
    ((2, 1, 3, 3), (1, 1), (0, 0, 0, 0)),         # pure zero-stuffing
])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_upfirdn2d_up2_polyphase(cuda, shape, kshape, pads, dtype):
    """up = 2, down = 1 takes the polyphase kernel (2 x 2 live taps per output, pad parities as template parameters)."""
    from fm3d import ops
    gen = torch.Generator().manual_seed(sum(shape) * 11 + kshape[0])
    x = torch.randn(*shape, generator=gen).to(dtype)
    k = torch.randn(*kshape, generator=gen)
    cfg = (2, 2, 1, 1) + pads
    ref = orc.upfirdn2d_ref(x.float(), k, *cfg)
    out = ops.upfirdn2d_planes(x.to(cuda), k.to(cuda), *cfg)
    assert tuple(out.shape) == tuple(ref.shape) and out.dtype == dtype
    tol = 1e-5 if dtype == torch.float32 else (2e-2 if dtype == torch.bfloat16 else 3e-3)
    torch.testing.assert_close(out.float().cpu(), ref, rtol=tol, atol=tol * 8)


def test_upfirdn2d_empty(cuda):
    from fm3d import ops
    k = orc.make_kernel_ref([1, 3, 3, 1]).to(cuda)
    out = ops.upfirdn2d_planes(torch.zeros(0, 3, 8, 8, device=cuda), k, 1, 1, 1, 1, 1, 1, 1, 1)
    assert out.shape == (0, 3, 7, 7)
    out = ops.upfirdn2d_planes(torch.zeros(1, 1, 2, 2, device=cuda), k, 1, 1, 1, 1, 0, 0, 0, 0)
    assert out.numel() == 0


def test_upfirdn2d_grad_and_gradgrad(cuda):
    """first and second order through UpFirDn2d vs autograd of the oracle (double backward
    is what R1 / path-length regularisation need, op/upfirdn2d.py:71-94)."""
    import op
    k = orc.make_kernel_ref([1, 3, 3, 1]) * 4
    for up, down, pad, hw in [(1, 1, (1, 1), 9), (2, 1, (2, 1), 6), (1, 2, (1, 1), 8), (1, 1, (2, 2), 8)]:
        x = torch.randn(2, 3, hw, hw, dtype=torch.float32)
        xr = x.clone().requires_grad_(True)
        yr = orc.upfirdn2d_api_ref(xr, k, up, down, pad)
        w = torch.randn_like(yr)
        gr, = torch.autograd.grad((yr * w).sum(), xr, create_graph=True)
        v = torch.randn_like(gr)
        xg = x.to(cuda).requires_grad_(True)
        yg = op.upfirdn2d(xg, k.to(cuda), up=up, down=down, pad=pad)
        wg = w.to(cuda).requires_grad_(True)
        gg, = torch.autograd.grad((yg * wg).sum(), xg, create_graph=True)
        torch.testing.assert_close(yg.detach().cpu(), yr.detach(), rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(gg.detach().cpu(), gr.detach(), rtol=1e-5, atol=1e-5)
        # second order: d/dw (g . v) = upfirdn(v)
        h, = torch.autograd.grad((gg * v.to(cuda)).sum(), wg)
        href = orc.upfirdn2d_api_ref(v, k, up, down, pad)
        torch.testing.assert_close(h.cpu(), href, rtol=1e-5, atol=1e-5)


# ----------------------------------------------------------------------------- bias_act
def test_bias_act_golden(cuda):
    import op
    g = load_golden("bias_act.npz")
    for name in g["names"]:
        x = _t(g[f"{name}.x"], cuda).requires_grad_(True)
        has_b = f"{name}.b" in g
        b = _t(g[f"{name}.b"], cuda).requires_grad_(True) if has_b else None
        y = op.fused_leaky_relu(x, b)
        np.testing.assert_allclose(y.detach().cpu().numpy(), g[f"{name}.y"], rtol=1e-6, atol=1e-6, err_msg=str(name))
        grads = torch.autograd.grad(y, [x] + ([b] if has_b else []), _t(g[f"{name}.gy"], cuda))
        np.testing.assert_allclose(grads[0].cpu().numpy(), g[f"{name}.gx"], rtol=1e-6, atol=1e-6)
        if has_b:
            np.testing.assert_allclose(grads[1].cpu().numpy(), g[f"{name}.gb"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("shape", [(32, 512, 4, 4), (4, 128, 64, 64), (3, 5, 7, 11), (32, 512), (2, 3, 1, 1), (5,),
                                   (2, 6, 24, 24), (3, 16, 40, 64)])     # rows that are / are not whole 128-vector chunks
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_bias_act_modes_vs_oracle(cuda, shape, dtype):
    from fm3d import ops
    gen = torch.Generator().manual_seed(len(shape) * 7 + shape[0])
    x = torch.randn(*shape, generator=gen).to(dtype)
    ch = shape[1] if len(shape) > 1 else shape[0]
    b = torch.randn(ch, generator=gen).to(dtype)
    r = torch.randn(*shape, generator=gen).to(dtype)
    tol = 1e-6 if dtype == torch.float32 else (1.6e-2 if dtype == torch.bfloat16 else 2e-3)
    for act, grad, use_b, use_r in [(3, 0, True, False), (3, 0, False, False), (3, 1, False, True), (3, 1, True, True),
                                    (1, 0, True, False), (1, 1, False, True), (3, 2, False, True), (1, 2, True, False)]:
        xs = x if len(shape) > 1 else x.view(1, -1)
        rs = r if len(shape) > 1 else r.view(1, -1)
        ref = orc.fused_bias_act_ref(xs.float(), b.float() if use_b else None, rs.float() if use_r else None,
                                     act, grad, 0.2, 2 ** 0.5).view(shape)
        out = ops.bias_act(x.to(cuda), b.to(cuda) if use_b else None, r.to(cuda) if use_r else None, act, grad, 0.2, 2 ** 0.5)
        assert out.dtype == dtype
        torch.testing.assert_close(out.float().cpu(), ref, rtol=tol, atol=tol, msg=f"act{act} grad{grad}")


@pytest.mark.parametrize("shape", [(2, 6, 24, 24), (4, 16, 64, 64), (3, 5, 7, 11), (2, 8, 48, 130), (32, 512)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("act", [3, 1])
def test_bias_act_grad_bias_fused(cuda, shape, dtype, act):
    """Gradient + fused bias-gradient reduction (every FusedLeakyReLU.backward, op/fused_act.py:42-48): the 16-byte path
    (rows of whole vectors, two vectors per thread in flight, tail vector) and the scalar fallback against the oracle; the
    bias gradient is the sum of the ROUNDED input gradient, as in the reference."""
    from fm3d import ops
    gen = torch.Generator().manual_seed(sum(shape) + act)
    g = torch.randn(*shape, generator=gen).to(dtype)
    r = torch.randn(*shape, generator=gen).to(dtype)
    ref = orc.fused_bias_act_ref(g.float(), None, r.float(), act, 1, 0.2, 2 ** 0.5).to(dtype)
    gin, gb = ops.bias_act_grad_bias(g.to(cuda), r.to(cuda), act, 0.2, 2 ** 0.5)
    tol = 1e-6 if dtype == torch.float32 else 1.6e-2
    torch.testing.assert_close(gin.float().cpu(), ref.float(), rtol=tol, atol=tol)
    dims = [0] + list(range(2, len(shape)))
    gb_ref = ref.float().sum(dim=dims)
    torch.testing.assert_close(gb.float().cpu(), gb_ref, rtol=2e-2 if dtype == torch.bfloat16 else 1e-4,
                               atol=(2e-2 if dtype == torch.bfloat16 else 1e-4) * max(1.0, float(gb_ref.abs().max())))


def test_bias_act_unaligned_and_empty(cuda):
    from fm3d import ops
    base = torch.randn(4 * 6 * 9 + 1, device=cuda)
    x = base[1:].view(4, 6, 9)              # 4-byte aligned only, inner not a multiple of 4
    b = torch.randn(6, device=cuda)
    ref = orc.fused_leaky_relu_ref(x.cpu(), b.cpu())
    torch.testing.assert_close(ops.bias_act(x, b).cpu(), ref, rtol=1e-6, atol=1e-6)
    assert ops.bias_act(torch.zeros(0, 4, 2, 2, device=cuda), torch.zeros(4, device=cuda)).shape == (0, 4, 2, 2)


def test_fused_leaky_relu_double_backward(cuda):
    """Second-order path (op/fused_act.py:55-62) against autograd of the oracle."""
    import op
    x = torch.randn(3, 6, 5, 5)
    b = torch.randn(6)
    xr, br = x.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = orc.fused_leaky_relu_ref(xr, br)
    gxr, gbr = torch.autograd.grad(yr.pow(2).sum(), [xr, br], create_graph=True)
    lr = gxr.pow(2).sum() + gbr.pow(2).sum()
    ggr = torch.autograd.grad(lr, [xr, br])
    xg, bg = x.to(cuda).requires_grad_(True), b.to(cuda).requires_grad_(True)
    yg = op.fused_leaky_relu(xg, bg)
    gxg, gbg = torch.autograd.grad(yg.pow(2).sum(), [xg, bg], create_graph=True)
    lg = gxg.pow(2).sum() + gbg.pow(2).sum()
    ggg = torch.autograd.grad(lg, [xg, bg])
    torch.testing.assert_close(gxg.detach().cpu(), gxr.detach(), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(gbg.detach().cpu(), gbr.detach(), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(ggg[0].cpu(), ggr[0], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(ggg[1].cpu(), ggr[1], rtol=1e-4, atol=1e-3)


def test_cpu_tensor_is_rejected():
    import op
    with pytest.raises(RuntimeError):
        op.fused_leaky_relu(torch.zeros(1, 2, 3, 3))
    with pytest.raises(RuntimeError):
        op.upfirdn2d(torch.zeros(1, 2, 3, 3), torch.ones(2, 2))
