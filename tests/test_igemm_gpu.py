"""tcgen05 implicit-GEMM convolution vs a CPU fp32 convolution on the same bf16-rounded data."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ident_tab(cout, dev, scale=1.0, shift=None, slope=1.0, post=1.0, B=None):
    rows = cout if B is None else B * cout
    t = torch.zeros(rows, 8, device=dev)
    t[:, 0] = scale
    if shift is not None:
        t[:, 1] = shift.to(dev).repeat(1 if B is None else B)
    t[:, 2] = slope
    t[:, 3] = post
    return t


def _run_conv(cuda, B, H, W, Cin, Cout, k, stride=1, pad=None, block_n=0, tile=(0, 0), seed=0, out_nchw=False):
    from fm3d import ops
    pad = k // 2 if pad is None else pad
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(B, Cin, H, W, generator=gen)
    w = torch.randn(Cout, Cin, k, k, generator=gen) / (Cin * k * k) ** 0.5
    xq = x.to(torch.bfloat16).float()
    wq = w.to(torch.bfloat16).float()
    ref = F.conv2d(xq, wq, stride=stride, padding=pad)
    OH, OW = ref.shape[2], ref.shape[3]
    xg = ops.nchw_to_nhwc_bf16(x.to(cuda))
    wg, _ = ops.prep_weight(w.to(cuda), 1.0)
    tab = _ident_tab(Cout, cuda)
    if out_nchw:
        out = torch.zeros(B, Cout, OH, OW, device=cuda)
    else:
        out = torch.zeros(B, OH, OW, (Cout + 7) // 8 * 8, device=cuda, dtype=torch.bfloat16)
    ops.conv_igemm(xg, wg, ops.conv_taps(k, k, pad), out, tab, B=B, H=H, W=W, Cin=Cin, Cout=Cout, OH=OH, OW=OW,
                   stride=stride, out_nchw_f32=out_nchw, block_n=block_n, tile_w=tile[0], tile_h=tile[1])
    torch.cuda.synchronize()
    got = out.cpu() if out_nchw else out[..., :Cout].float().permute(0, 3, 1, 2).cpu()
    return got, ref


def test_nhwc_roundtrip(cuda):
    from fm3d import ops
    x = torch.randn(3, 13, 5, 7)
    s = torch.rand(3, 13) + 0.5
    y = ops.nchw_to_nhwc_bf16(x.to(cuda), s.to(cuda))
    assert y.shape == (3, 5, 7, 16)
    ref = (x * s[:, :, None, None]).to(torch.bfloat16)
    assert torch.equal(y[..., :13].permute(0, 3, 1, 2).cpu(), ref)
    assert torch.count_nonzero(y[..., 13:]) == 0
    back = ops.nhwc_bf16_to_nchw(y, 13)
    assert torch.equal(back.cpu(), ref.float())


def test_gemm_1x1_minimal(cuda):
    """Pure GEMM (1x1 conv): M=128 pixels, K=64, N=64 -- the smallest tcgen05 problem."""
    got, ref = _run_conv(cuda, 1, 8, 16, 64, 64, 1, out_nchw=True)
    torch.testing.assert_close(got, ref, rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("B,H,W,Cin,Cout,k", [
    (1, 8, 16, 128, 64, 1),      # K loop over 2 chunks
    (2, 16, 16, 64, 128, 3),     # 3x3 taps, zero padding via TMA OOB fill
    (1, 32, 32, 256, 256, 3),    # BLOCK_N 256 path
    (4, 4, 4, 512, 512, 3),      # multi-sample tiles (tile_b = 8 > B)
    (3, 8, 8, 96, 40, 3),        # ragged Cin (not /64) and Cout (not /8 of tile)
    (2, 64, 64, 128, 128, 3),    # several persistent tiles per CTA
    (1, 128, 128, 64, 64, 3),    # tile_w = 128 rows
    (2, 5, 7, 72, 24, 3),        # odd spatial sizes
])
def test_conv_stride1(cuda, B, H, W, Cin, Cout, k):
    got, ref = _run_conv(cuda, B, H, W, Cin, Cout, k, seed=H + Cin)
    # operands identical (bf16-rounded); fp32 accumulate order + bf16 output rounding
    torch.testing.assert_close(got, ref, rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("bn", [64, 128, 256])
def test_conv_block_n(cuda, bn):
    got, ref = _run_conv(cuda, 2, 16, 16, 128, 256, 3, block_n=bn, seed=bn, out_nchw=True)
    torch.testing.assert_close(got, ref, rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,pad", [
    (2, 32, 32, 64, 64, 3, 1),
    (1, 64, 64, 128, 256, 3, 1),
    (2, 16, 16, 64, 128, 1, 0),
    (3, 8, 8, 512, 512, 3, 1),
    (2, 2, 2, 512, 512, 3, 1),   # pSp style-head tail 2x2 -> 1x1
])
def test_conv_stride2(cuda, B, H, W, Cin, Cout, k, pad):
    got, ref = _run_conv(cuda, B, H, W, Cin, Cout, k, stride=2, pad=pad, seed=W + Cout, out_nchw=True)
    torch.testing.assert_close(got, ref, rtol=2e-3, atol=2e-3)


def test_conv_epilogue_tables_noise_residual_rgb(cuda):
    """Full epilogue: per-sample scale, bias, noise, residual, leaky-ReLU, post-scale, fused RGB."""
    from fm3d import ops
    B, H, W, Cin, Cout = 3, 16, 16, 64, 320          # Cout -> 2 N-tiles: RGB goes through atomics
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(B, Cin, H, W, generator=gen)
    w = torch.randn(Cout, Cin, 3, 3, generator=gen) / (Cin * 9) ** 0.5
    d = torch.rand(B, Cout, generator=gen) + 0.5
    bias = torch.randn(Cout, generator=gen) * 0.2
    post = torch.randn(B, Cout, generator=gen)
    wr = torch.randn(B, Cout, 3, generator=gen) * 0.1
    noise = torch.randn(B, H, W, generator=gen)
    nw = torch.tensor([0.3])
    res = torch.randn(B, Cout, H, W, generator=gen).to(torch.bfloat16)
    acc = F.conv2d(x.to(torch.bfloat16).float(), w.to(torch.bfloat16).float(), padding=1)
    v = acc * d[:, :, None, None] + bias[None, :, None, None] + 0.3 * noise[:, None] + res.float()
    v = torch.where(v > 0, v, 0.2 * v)
    ref_rgb = torch.einsum("bchw,bcj->bjhw", v, wr)
    ref = v * post[:, :, None, None]
    tab = torch.zeros(B, Cout, 8)
    tab[..., 0] = d; tab[..., 1] = bias; tab[..., 2] = 0.2; tab[..., 3] = post; tab[..., 4:7] = wr
    out = torch.zeros(B, H, W, Cout, device=cuda, dtype=torch.bfloat16)
    rgb = torch.zeros(B, H, W, 4, device=cuda)
    resg = res.permute(0, 2, 3, 1).contiguous().to(cuda)
    wg, _ = ops.prep_weight(w.to(cuda), 1.0)
    ops.conv_igemm(ops.nchw_to_nhwc_bf16(x.to(cuda)), wg, ops.conv_taps(3, 3, 1), out, tab.to(cuda),
                   B=B, H=H, W=W, Cin=Cin, Cout=Cout, OH=H, OW=W, tab_per_sample=True,
                   noise=noise.to(cuda), noise_per_sample=True, noise_w=nw.to(cuda), residual=resg, rgb=rgb)
    torch.cuda.synchronize()
    torch.testing.assert_close(out.float().permute(0, 3, 1, 2).cpu(), ref, rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(rgb[..., :3].permute(0, 3, 1, 2).cpu(), ref_rgb, rtol=2e-3, atol=2e-3)
