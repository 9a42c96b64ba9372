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


# ------------------------------------------------------------------ generalisations used by the encoders
def _tab(cout, dev, groups=1, shift=None, slope=1.0):
    t = torch.zeros(groups, cout, 8, device=dev)
    t[..., 0] = 1.0; t[..., 2] = slope; t[..., 3] = 1.0
    if shift is not None:
        t[..., 1] = shift.to(dev)
    return t


@pytest.mark.parametrize("Bg,H,stride", [(4, 8, 2), (2, 2, 2), (32, 1, 1), (3, 16, 1)])
def test_conv_grouped(cuda, Bg, H, stride):
    """groups: per-group weights and tables over a group-major batch (pSp map2style heads)."""
    from fm3d import ops
    G, C = 3, 64
    gen = torch.Generator().manual_seed(Bg * 10 + H)
    k = 3 if H > 1 else 1
    pad = k // 2
    x = torch.randn(G * Bg, C, H, H, generator=gen)
    w = torch.randn(G, C, C, k, k, generator=gen) / (C * k * k) ** 0.5
    bias = torch.randn(G, C, generator=gen)
    ref = torch.cat([F.leaky_relu(F.conv2d(x[g * Bg:(g + 1) * Bg].to(torch.bfloat16).float(), w[g].to(torch.bfloat16).float(),
                                           bias[g], stride=stride, padding=pad), 0.01) for g in range(G)], 0)
    OH = ref.shape[2]
    wq = torch.cat([ops.prep_weight(w[g].to(cuda), 1.0, want_wsq=False)[0] for g in range(G)], 0).contiguous()
    out = torch.zeros(G * Bg, C, OH, OH, device=cuda)
    ops.conv_igemm(ops.nchw_to_nhwc_bf16(x.to(cuda)), wq, ops.conv_taps(k, k, pad), out, _tab(C, cuda, G, bias, 0.01),
                   B=G * Bg, H=H, W=H, Cin=C, Cout=C, OH=OH, OW=OH, stride=stride, groups=G, w_rows=C, out_nchw_f32=True)
    torch.cuda.synchronize()
    torch.testing.assert_close(out.cpu(), ref, rtol=2e-3, atol=2e-3)


def test_conv_concatenated_output(cuda):
    """out_cgroup: one launch computes n heads that share an input, written head-major."""
    from fm3d import ops
    B, C, H, n = 2, 64, 8, 3
    gen = torch.Generator().manual_seed(77)
    x = torch.randn(B, C, H, H, generator=gen)
    w = torch.randn(n * C, C, 3, 3, generator=gen) / (C * 9) ** 0.5
    ref = F.conv2d(x.to(torch.bfloat16).float(), w.to(torch.bfloat16).float(), stride=2, padding=1)     # [B, n*C, 4, 4]
    wq, _ = ops.prep_weight(w.to(cuda), 1.0, want_wsq=False)
    out = torch.zeros(n * B, 4, 4, C, device=cuda, dtype=torch.bfloat16)
    ops.conv_igemm(ops.nchw_to_nhwc_bf16(x.to(cuda)), wq, ops.conv_taps(3, 3, 1), out, _tab(n * C, cuda), B=B, H=H, W=H,
                   Cin=C, Cout=n * C, OH=4, OW=4, stride=2, out_cgroup=C, out_gstride=B * 16 * C, out_cstride=C)
    torch.cuda.synchronize()
    got = out.float().view(n, B, 4, 4, C).permute(1, 0, 4, 2, 3).reshape(B, n * C, 4, 4).cpu()
    torch.testing.assert_close(got, ref, rtol=1e-2, atol=1e-2)


def test_conv_input_bn_border_table(cuda):
    """BatchNorm in front of a zero-padded 3x3 conv folded into weights + 9-class border table."""
    from fm3d import ops
    from fm3d.encoder_engine import _border_table, _table
    B, C, O, H = 2, 64, 128, 12
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(B, C, H, H, generator=gen)
    a = torch.rand(C, generator=gen) + 0.5
    b = torch.randn(C, generator=gen) * 0.5
    w = torch.randn(O, C, 3, 3, generator=gen) / (C * 9) ** 0.5
    xq = x.to(torch.bfloat16).float()
    wf = (w * a.view(1, -1, 1, 1)).to(torch.bfloat16).float()
    # reference: conv(zero_pad(a*x + b)) with the same rounded operands
    ref = F.conv2d(xq, wf, padding=1) + F.conv2d(torch.ones(B, C, H, H) * b.view(1, -1, 1, 1), w, padding=1)
    full, corr = _border_table(w.to(cuda), b.to(cuda))
    wq, _ = ops.prep_weight((w * a.view(1, -1, 1, 1)).to(cuda), 1.0, want_wsq=False)
    out = torch.zeros(B, O, H, H, device=cuda)
    ops.conv_igemm(ops.nchw_to_nhwc_bf16(x.to(cuda)), wq, ops.conv_taps(3, 3, 1), out, _table(O, cuda, None, full, 1.0),
                   B=B, H=H, W=H, Cin=C, Cout=O, OH=H, OW=H, border_tab=corr, out_nchw_f32=True)
    torch.cuda.synchronize()
    torch.testing.assert_close(out.cpu(), ref, rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("k,stride,pad,S", [(7, 2, 3, 64), (3, 1, 1, 32)])
def test_conv_stem_window(cuda, k, stride, pad, S):
    """3-channel stems: 8 pixels x 8 channels of a padded row as one 64-wide K chunk."""
    from fm3d import ops
    from fm3d.encoder_engine import _stem_weight, _table
    B, O = 2, 64
    gen = torch.Generator().manual_seed(k)
    x = torch.randn(B, 3, S, S, generator=gen)
    w = torch.randn(O, 3, k, k, generator=gen) / (3 * k * k) ** 0.5
    ref = F.conv2d(x.to(torch.bfloat16).float(), w.to(torch.bfloat16).float(), stride=stride, padding=pad)
    oh = ref.shape[2]
    Hp, Wp = S + 2 * pad, max(S + 2 * pad, stride * (oh - 1) + 8)
    packed = ops.image_to_nhwc8_padded(x.to(cuda), pad, pad, Hp, Wp)
    out = torch.zeros(B, O, oh, oh, device=cuda)
    ops.conv_igemm(packed, _stem_weight(w.to(cuda)), [(ky, 0, ky) for ky in range(k)], out, _table(O, cuda), B=B, H=Hp, W=oh,
                   Cin=64, Cout=O, OH=oh, OW=oh, stride_x=1, stride_y=stride, x_pixstride=8 * stride, x_rowstride=Wp * 8,
                   x_imgstride=Hp * Wp * 8, out_nchw_f32=True)
    torch.cuda.synchronize()
    torch.testing.assert_close(out.cpu(), ref, rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("B,C,H,W,pad_t,pad_l,Hp,Wp", [(2, 3, 64, 64, 3, 3, 70, 72), (3, 3, 32, 32, 1, 1, 34, 34), (1, 5, 7, 9, 0, 2, 9, 13),
                                                    (2, 8, 5, 6, 2, 0, 8, 7), (32, 3, 256, 256, 1, 1, 258, 258)])
def test_image_pack_padded(cuda, B, C, H, W, pad_t, pad_l, Hp, Wp):
    """fp32 NCHW image -> zero-padded bf16 [B, Hp, Wp, 8] (the stems' operand), four output pixels per thread: row tails
    that are not a multiple of 4, every channel count up to 8, borders and spare channels zero."""
    from fm3d import ops
    gen = torch.Generator().manual_seed(B * 100 + W)
    x = torch.randn(B, C, H, W, generator=gen)
    ref = torch.zeros(B, Hp, Wp, 8)
    ref[:, pad_t:pad_t + H, pad_l:pad_l + W, :C] = x.permute(0, 2, 3, 1)
    out = torch.full((B, Hp, Wp, 8), float("nan"), device=cuda, dtype=torch.bfloat16)
    ops.image_to_nhwc8_padded(x.to(cuda), pad_t, pad_l, Hp, Wp, out=out)
    assert torch.equal(out.cpu(), ref.to(torch.bfloat16))


def test_encoder_helper_kernels(cuda):
    from fm3d import ops
    gen = torch.Generator().manual_seed(9)
    B, C, H = 3, 64, 10
    x = torch.randn(B, C, H, H, generator=gen).to(torch.bfloat16)
    xg = x.permute(0, 2, 3, 1).contiguous().to(cuda)
    # max pool
    out = torch.empty(B, 5, 5, C, device=cuda, dtype=torch.bfloat16)
    ops.maxpool3x3s2_nhwc(xg, out)
    ref = F.max_pool2d(x.float(), 3, 2, 1)
    assert torch.equal(out.float().permute(0, 3, 1, 2).cpu(), ref)
    # avg pool -> NCHW fp32
    torch.testing.assert_close(ops.avgpool_nhwc_to_nchw(xg, C, 2, 2).cpu(), F.avg_pool2d(x.float(), 2, 2), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(ops.avgpool_nhwc_to_nchw(xg, C, H, H).cpu(), x.float().mean((2, 3), keepdim=True), rtol=1e-5, atol=1e-5)
    # bilinear align_corners
    up = torch.empty(B, 20, 20, C, device=cuda, dtype=torch.bfloat16)
    ops.bilinear_up_nhwc(xg, 20, 20, up)
    refu = F.interpolate(x.float(), size=(20, 20), mode="bilinear", align_corners=True)
    torch.testing.assert_close(up.float().permute(0, 3, 1, 2).cpu(), refu, rtol=1e-2, atol=1e-2)
    # SE block with a strided identity shortcut
    w1 = torch.randn(C // 16, C, generator=gen) * 0.3
    w2 = torch.randn(C, C // 16, generator=gen) * 0.3
    sc = torch.randn(B, C, 2 * H, 2 * H, generator=gen).to(torch.bfloat16)
    gate = torch.sigmoid(F.linear(F.relu(F.linear(x.float().mean((2, 3)), w1)), w2))
    refse = x.float() * gate[:, :, None, None] + sc.float()[:, :, ::2, ::2]
    sums = torch.zeros(B, 512, device=cuda); gbuf = torch.empty(B, 512, device=cuda)
    o = torch.empty_like(xg)
    ops.se_block_nhwc(xg, C, w1.to(cuda), w2.to(cuda), sc.permute(0, 2, 3, 1).contiguous().to(cuda), 2, sums, gbuf, o)
    torch.testing.assert_close(o.float().permute(0, 3, 1, 2).cpu(), refse, rtol=1e-2, atol=2e-2)
    assert float(sums.abs().sum()) == 0.0          # re-zeroed for the next block


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 6, 128, 64, 64), (1, 9, 256, 128, 128), (2, 5, 200, 72, 40), (1, 3, 130, 192, 128)])
def test_conv_row_patch_mode(cuda, B, H, W, Cin, Cout):
    """3x3/s1/p1 with 128-pixel row tiles: one 130-pixel patch feeds the three horizontal taps
    through shifted UMMA descriptors."""
    from fm3d import ops
    gen = torch.Generator().manual_seed(W + Cin)
    x = torch.randn(B, Cin, H, W, generator=gen)
    w = torch.randn(Cout, Cin, 3, 3, generator=gen) / (Cin * 9) ** 0.5
    ref = F.conv2d(x.to(torch.bfloat16).float(), w.to(torch.bfloat16).float(), padding=1)
    wq, _ = ops.prep_weight(w.to(cuda), 1.0, want_wsq=False)
    out = torch.zeros(B, Cout, H, W, device=cuda)
    ops.conv_igemm(ops.nchw_to_nhwc_bf16(x.to(cuda)), wq, ops.conv_taps(3, 3, 1), out, _tab(Cout, cuda), B=B, H=H, W=W,
                   Cin=Cin, Cout=Cout, OH=H, OW=W, out_nchw_f32=True)
    torch.cuda.synchronize()
    torch.testing.assert_close(out.cpu(), ref, rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("B,h,Cin,Cout", [(2, 8, 64, 64), (1, 16, 128, 128), (3, 4, 512, 512), (2, 32, 256, 128), (1, 7, 72, 40)])
def test_conv_fused_upsample_mode(cuda, B, h, Cin, Cout):
    """upmode: stride-2 transposed 3x3 conv (F.conv_transpose2d, stylegan2.py:276) as four output
    parities accumulated from four shared shifted input views in one launch."""
    from fm3d import ops
    gen = torch.Generator().manual_seed(h + Cin)
    x = torch.randn(B, Cin, h, h, generator=gen)
    w = torch.randn(Cout, Cin, 3, 3, generator=gen) / (Cin * 9) ** 0.5
    ref = F.conv_transpose2d(x.to(torch.bfloat16).float(), w.to(torch.bfloat16).float().transpose(0, 1), stride=2)
    assert ref.shape[-1] == 2 * h + 1
    wq, _ = ops.prep_weight(w.to(cuda), 1.0, want_wsq=False)
    cs = (Cout + 7) // 8 * 8
    out = torch.zeros(B, 2 * h + 1, 2 * h + 1, cs, device=cuda, dtype=torch.bfloat16)
    tw = min(16, 1 << (h).bit_length())
    ops.conv_igemm(ops.nchw_to_nhwc_bf16(x.to(cuda)), wq, ops.conv_taps(3, 3, 1), out, _tab(Cout, cuda), B=B, H=h, W=h,
                   Cin=Cin, Cout=Cout, OH=h + 1, OW=h + 1, out_H=2 * h + 1, out_W=2 * h + 1, out_ys=2, out_xs=2,
                   tile_w=tw, tile_h=max(1, min(8, 128 // tw)), upmode=True)
    torch.cuda.synchronize()
    torch.testing.assert_close(out[..., :Cout].float().permute(0, 3, 1, 2).cpu(), ref, rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 32, 32, 64, 64), (2, 16, 64, 64, 128), (1, 48, 40, 128, 64), (4, 64, 64, 64, 64)])
def test_conv_halo_patch_resident_weights(cuda, B, H, W, Cin, Cout):
    """Halo-patch mode with the whole weight set resident in shared memory (single n-tile, small Cin*Cout):
    8x16-pixel tiles, one (18 x 10)-pixel input patch per channel chunk, taps as shifted descriptors; N = 128
    additionally runs as a CTA pair (cta_group::2)."""
    got, ref = _run_conv(cuda, B, H, W, Cin, Cout, 3, seed=H * 3 + Cout, out_nchw=True)
    torch.testing.assert_close(got, ref, rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("B,h,Cin,Cout", [(2, 16, 64, 128), (1, 24, 128, 64), (3, 12, 72, 32)])
def test_conv_paired_parity_upsample(cuda, B, h, Cin, Cout):
    """Stride-2 transposed 3x3 conv as TWO launches (one per output-row parity) whose N dimension concatenates
    the even- and odd-column parities (zero weight blocks where the odd parity has no tap): the engine's
    Cout <= 128 up-conv path, against F.conv_transpose2d (stylegan2.py:276)."""
    from fm3d import engine, ops
    gen = torch.Generator().manual_seed(h * 7 + Cin)
    x = torch.randn(B, Cin, h, h, generator=gen)
    w = torch.randn(Cout, Cin, 3, 3, generator=gen) / (Cin * 9) ** 0.5
    ref = F.conv_transpose2d(x.to(torch.bfloat16).float(), w.to(torch.bfloat16).float().transpose(0, 1), stride=2)
    wq, _ = ops.prep_weight(w.to(cuda), 1.0, want_wsq=False)          # [9][Cout][cin_stride]
    cs = (Cout + 7) // 8 * 8
    out = torch.full((B, 2 * h + 1, 2 * h + 1, cs), 7.0, device=cuda, dtype=torch.bfloat16)
    xg = ops.nchw_to_nhwc_bf16(x.to(cuda))
    for py, views in engine._PAIR_VIEWS.items():
        wp = torch.zeros(len(views), 2 * Cout, wq.shape[2], device=cuda, dtype=torch.bfloat16)
        for v, (_, t0, t1) in enumerate(views):
            wp[v, :Cout] = wq[t0, :Cout]
            if t1 is not None:
                wp[v, Cout:] = wq[t1, :Cout]
        taps = [(dy, dx, v) for v, ((dy, dx), _, _) in enumerate(views)]
        ops.conv_igemm(xg, wp, taps, out, None, B=B, H=h, W=h, Cin=Cin, Cout=2 * Cout, OH=h + 1 - py, OW=h + 1,
                       out_H=2 * h + 1, out_W=2 * h + 1, out_y0=py, out_x0=0, out_ys=2, out_xs=2, tab_per_sample=False,
                       out_cgroup=Cout, out_gstride=cs, out_cstride=cs, out_cgroup_ow_shrink=1)
    torch.cuda.synchronize()
    torch.testing.assert_close(out[..., :Cout].float().permute(0, 3, 1, 2).cpu(), ref, rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("B,h,Cin,Cout", [(2, 16, 128, 512), (1, 8, 64, 64), (3, 12, 64, 128), (2, 32, 128, 320), (5, 20, 72, 96)])
def test_conv_upsampled_residual(cuda, B, h, Cin, Cout):
    """1x1 lateral conv + bilinear (align_corners=True) upsampling of a half-resolution map sampled in the
    epilogue: the FPN top-down add of the pSp encoder (psp_encoders.py:81-98) without the intermediate tensor."""
    from fm3d import ops
    gen = torch.Generator().manual_seed(h + Cout)
    x = torch.randn(B, Cin, 2 * h, 2 * h, generator=gen)
    low = torch.randn(B, Cout, h, h, generator=gen)
    w = torch.randn(Cout, Cin, 1, 1, generator=gen) / Cin ** 0.5
    ref = F.conv2d(x.to(torch.bfloat16).float(), w.to(torch.bfloat16).float()) + \
        F.interpolate(low.to(torch.bfloat16).float(), size=(2 * h, 2 * h), mode="bilinear", align_corners=True)
    wq, _ = ops.prep_weight(w.to(cuda), 1.0, want_wsq=False)
    out = torch.zeros(B, 2 * h, 2 * h, Cout, device=cuda, dtype=torch.bfloat16)
    ops.conv_igemm(ops.nchw_to_nhwc_bf16(x.to(cuda)), wq, [(0, 0, 0)], out, _tab(Cout, cuda), B=B, H=2 * h, W=2 * h, Cin=Cin,
                   Cout=Cout, OH=2 * h, OW=2 * h, residual_up=ops.nchw_to_nhwc_bf16(low.to(cuda)))
    torch.cuda.synchronize()
    torch.testing.assert_close(out.float().permute(0, 3, 1, 2).cpu(), ref, rtol=1e-2, atol=2e-2)


@pytest.mark.parametrize("B,h,Cin,Cout,pair", [(2, 16, 64, 128, True), (3, 12, 72, 32, True), (2, 16, 128, 256, False),
                                                (5, 4, 64, 64, False), (1, 33, 64, 64, False), (4, 8, 512, 512, False)])
def test_conv_tall_image_upsample(cuda, B, h, Cin, Cout, pair):
    """The engine's stride-2 transposed conv: the batch as ONE tall image [B*(h+1), h+1] whose zero separator row /
    column per image is the conv's padding, phases written into [B, 2h+2, 2h+2] (spare row / column = exact zeros),
    against F.conv_transpose2d (stylegan2.py:276)."""
    import types
    from fm3d import engine, ops
    gen = torch.Generator().manual_seed(h * 11 + Cin + B)
    x = torch.randn(B, Cin, h, h, generator=gen)
    w = torch.randn(Cout, Cin, 3, 3, generator=gen) / (Cin * 9) ** 0.5
    ref = F.conv_transpose2d(x.to(torch.bfloat16).float(), w.to(torch.bfloat16).float().transpose(0, 1), stride=2)
    wq, _ = ops.prep_weight(w.to(cuda), 1.0, want_wsq=False)
    L = types.SimpleNamespace(cin=Cin, cout=Cout, wq=wq, wpair=None, wpair_all=None)
    if pair:
        n0 = len(engine._PAIR_VIEWS[0])
        L.wpair_all = torch.zeros(n0 + len(engine._PAIR_VIEWS[1]), 2 * Cout, wq.shape[2], device=cuda, dtype=torch.bfloat16)
        L.wpair = {0: L.wpair_all[:n0], 1: L.wpair_all[n0:]}
        for py, views in engine._PAIR_VIEWS.items():
            wp = L.wpair[py]
            for v, (_, t0, t1) in enumerate(views):
                wp[v, :Cout] = wq[t0, :Cout]
                if t1 is not None:
                    wp[v, Cout:] = wq[t1, :Cout]
    cs_in, cs = (Cin + 7) // 8 * 8, (Cout + 7) // 8 * 8
    xp = torch.zeros(B, h + 1, h + 1, cs_in, device=cuda, dtype=torch.bfloat16)
    xp[:, :h, :h] = ops.nchw_to_nhwc_bf16(x.to(cuda))
    t = torch.full((B, 2 * h + 2, 2 * h + 2, cs), 7.0, device=cuda, dtype=torch.bfloat16)
    engine.SynthesisPlan._up_conv(None, L, xp, t, B, h)
    torch.cuda.synchronize()
    torch.testing.assert_close(t[:, :2 * h + 1, :2 * h + 1, :Cout].float().permute(0, 3, 1, 2).cpu(), ref, rtol=1e-2, atol=1e-2)
    assert float(t[:, 2 * h + 1].abs().max()) == 0.0 and float(t[:, :, 2 * h + 1].abs().max()) == 0.0


def test_conv_output_pitch(cuda):
    """out_pitch_h / out_pitch_w: the NHWC output lives in a larger zero-initialised tensor (one spare row / column
    per image) while noise and the fused RGB sums stay on the logical grid."""
    from fm3d import ops
    B, H, Cin, Cout = 3, 16, 64, 128
    gen = torch.Generator().manual_seed(77)
    x = torch.randn(B, Cin, H, H, generator=gen)
    w = torch.randn(Cout, Cin, 3, 3, generator=gen) / (Cin * 9) ** 0.5
    noise = torch.randn(B, H, H, generator=gen)
    tab = torch.zeros(B, Cout, 8)
    tab[..., 0] = 1.0; tab[..., 2] = 0.2; tab[..., 3] = 1.0; tab[..., 4:7] = torch.randn(B, Cout, 3, generator=gen) * 0.1
    wg, _ = ops.prep_weight(w.to(cuda), 1.0)
    xg = ops.nchw_to_nhwc_bf16(x.to(cuda))
    outs, rgbs = [], []
    for pitch in (0, H + 1):
        out = torch.zeros(B, H + (1 if pitch else 0), H + (1 if pitch else 0), Cout, device=cuda, dtype=torch.bfloat16)
        rgb = torch.zeros(B, H, H, 4, device=cuda)
        ops.conv_igemm(xg, wg, ops.conv_taps(3, 3, 1), out, tab.to(cuda), B=B, H=H, W=H, Cin=Cin, Cout=Cout, OH=H, OW=H,
                       tab_per_sample=True, noise=noise.to(cuda), noise_per_sample=True, rgb=rgb,
                       out_pitch_h=pitch, out_pitch_w=pitch)
        torch.cuda.synchronize()
        outs.append(out); rgbs.append(rgb)
    assert torch.equal(outs[1][:, :H, :H], outs[0])
    assert float(outs[1][:, H].abs().max()) == 0.0 and float(outs[1][:, :, H].abs().max()) == 0.0
    torch.testing.assert_close(rgbs[1], rgbs[0], rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------ dynamic tile schedule / SM partitions
@pytest.mark.parametrize("B,H,Cin,Cout,k,stride,cap", [
    (4, 64, 64, 64, 3, 1, 0),       # halo patch, resident weights (hpw)
    (4, 64, 64, 64, 3, 1, 5),       # the same confined to 5 SMs
    (2, 64, 128, 256, 3, 1, 0),     # CTA pairs + halo patch
    (2, 64, 128, 256, 3, 1, 7),     # CTA pairs, odd cap (rounded down to whole clusters)
    (2, 128, 64, 128, 3, 1, 12),    # row-patch mode
    (2, 64, 128, 512, 1, 1, 0),     # 1x1, several n-tiles
    (3, 32, 128, 256, 3, 2, 9),     # stride 2, generic ring
    (8, 16, 256, 256, 3, 1, 0),     # small M (split-K eligible when a workspace is given)
    (2, 64, 8, 64, 3, 1, 3),        # single K chunk: many more stages than a tile uses (queue depth)
])
def test_dynamic_tile_schedule_matches_static(cuda, B, H, Cin, Cout, k, stride, cap):
    """fm_conv_desc.tile_counter / max_ctas change which CTA computes a tile, never the tile: outputs are identical to
    the static round robin, and the counter pair is back at zero after every launch."""
    from fm3d import ops
    gen = torch.Generator().manual_seed(H + Cin + cap)
    x = torch.randn(B, H, H, (Cin + 7) // 8 * 8, generator=gen).to(torch.bfloat16).to(cuda)
    w = (torch.randn(k * k, Cout, (Cin + 7) // 8 * 8, generator=gen) / (Cin * k * k) ** 0.5).to(torch.bfloat16).to(cuda)
    tab = torch.zeros(1, Cout, 8, device=cuda)
    tab[..., 0] = torch.rand(Cout, generator=gen).to(cuda) + 0.5
    tab[..., 1] = 0.1
    tab[..., 2] = 0.2
    tab[..., 3] = 1.0
    OH = (H + 2 * (k // 2) - k) // stride + 1
    outs = []
    ctr = torch.zeros(2, device=cuda, dtype=torch.int32)
    for c in (None, ctr, ctr):                           # static, dynamic, dynamic again on the re-armed counters
        out = torch.zeros(B, OH, OH, Cout, device=cuda, dtype=torch.bfloat16)
        ops.conv_igemm(x, w, ops.conv_taps(k, k, k // 2), out, tab, B=B, H=H, W=H, Cin=Cin, Cout=Cout, OH=OH, OW=OH,
                       stride=stride, tile_counter=c, max_ctas=cap if c is not None else 0, ksplit=1)
        torch.cuda.synchronize()
        assert ctr.tolist() == [0, 0]
        outs.append(out)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    ref = F.conv2d(x[..., :Cin].float().permute(0, 3, 1, 2).cpu(), w[..., :Cin].float().view(k, k, Cout, Cin).permute(2, 3, 0, 1).cpu(),
                   stride=stride, padding=k // 2)
    ref = ref * tab[0, :, 0].cpu().view(1, -1, 1, 1) + 0.1
    ref = torch.where(ref > 0, ref, 0.2 * ref)
    torch.testing.assert_close(outs[1].float().permute(0, 3, 1, 2).cpu(), ref, rtol=2e-2, atol=2e-2)


def test_dynamic_tile_schedule_up_phases(cuda):
    """Several output phases in one launch (phase = tile coordinate) under the dynamic schedule and an SM cap."""
    from fm3d import ops
    B, h, Cin, Cout = 2, 32, 128, 128
    gen = torch.Generator().manual_seed(3)
    x = torch.zeros(B, h + 1, h + 1, Cin, dtype=torch.bfloat16)
    x[:, :h, :h] = torch.randn(B, h, h, Cin, generator=gen).to(torch.bfloat16)
    x = x.to(cuda)
    w = (torch.randn(9, Cout, Cin, generator=gen) / (Cin * 9) ** 0.5).to(torch.bfloat16).to(cuda)
    from fm3d.engine import _up_phase_taps
    taps, ph = [], []
    for py in (0, 1):
        for px in (0, 1):
            tl = _up_phase_taps(py, px)
            taps += tl
            ph.append((len(tl), py, px))
    outs = []
    ctr = torch.zeros(2, device=cuda, dtype=torch.int32)
    # (static, all SMs: phase-major), (dynamic), (dynamic, 10 SMs), (static, 4 SMs: >= 6 tiles per cluster, so a cluster
    # takes whole tiles with a rotated phase order)
    for c, cap in ((None, 0), (ctr, 0), (ctr, 10), (None, 4)):
        out = torch.zeros(1, B * (2 * h + 2), 2 * h + 2, Cout, device=cuda, dtype=torch.bfloat16)
        ops.conv_igemm(x.view(1, B * (h + 1), h + 1, Cin), w, taps, out, None, B=1, H=B * (h + 1), W=h + 1, Cin=Cin, Cout=Cout,
                       OH=B * (h + 1), OW=h + 1, out_H=B * (2 * h + 2), out_W=2 * h + 2, out_ys=2, out_xs=2, phases=ph,
                       tile_counter=c, max_ctas=cap)
        torch.cuda.synchronize()
        assert ctr.tolist() == [0, 0]
        outs.append(out)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2]) and torch.equal(outs[0], outs[3])
    assert float(outs[0].float().abs().max()) > 0.1


@pytest.mark.parametrize("B,H,Cin,Cout,stride,bn", [
    (3, 32, 64, 64, 1, 0),       # halo patch, resident weights
    (2, 64, 64, 128, 2, 0),      # stride 2
    (2, 128, 64, 64, 1, 0),      # row-patch mode (R accumulators per tile)
    (2, 16, 128, 320, 1, 0),     # ragged Cout, several n-tiles, CTA pairs
    (5, 16, 256, 512, 2, 128),   # 8x8 outputs: half-empty tiles would straddle images -> tile_b must stay 1
])
def test_conv_fused_channel_sums(cuda, B, H, Cin, Cout, stride, bn):
    """fm_conv_desc.colsum: per-(sample, channel) sums of the conv's output accumulated in the epilogue (the SE squeeze)
    against a sum over the stored tensor."""
    from fm3d import ops
    gen = torch.Generator().manual_seed(H + Cout)
    x = torch.randn(B, H, H, Cin, generator=gen).to(torch.bfloat16).to(cuda)
    w = (torch.randn(9, Cout, Cin, generator=gen) / (Cin * 9) ** 0.5).to(torch.bfloat16).to(cuda)
    tab = torch.zeros(1, Cout, 8, device=cuda)
    tab[..., 0] = torch.rand(Cout, generator=gen).to(cuda) + 0.5
    tab[..., 1] = torch.randn(Cout, generator=gen).to(cuda) * 0.3
    tab[..., 2] = 0.25
    tab[..., 3] = 1.0
    OH = (H + 2 - 3) // stride + 1
    out = torch.zeros(B, OH, OH, (Cout + 7) // 8 * 8, device=cuda, dtype=torch.bfloat16)
    sums = torch.zeros(B, Cout, device=cuda)
    if OH * OH < 128:
        with pytest.raises(RuntimeError, match="colsum"):
            ops.conv_igemm(x, w, ops.conv_taps(3, 3, 1), out, tab, B=B, H=H, W=H, Cin=Cin, Cout=Cout, OH=OH, OW=OH,
                           stride=stride, colsum=sums, ksplit=1, block_n=bn)
        return
    for _ in range(2):                 # accumulates: two launches = twice the sums
        ops.conv_igemm(x, w, ops.conv_taps(3, 3, 1), out, tab, B=B, H=H, W=H, Cin=Cin, Cout=Cout, OH=OH, OW=OH,
                       stride=stride, colsum=sums, ksplit=1, block_n=bn)
    torch.cuda.synchronize()
    ref = 2.0 * out[..., :Cout].float().sum((1, 2))
    # out is bf16-rounded, the fused sums are taken before rounding: relative to the sum of magnitudes
    mag = 2.0 * out[..., :Cout].float().abs().sum((1, 2)) + 1e-3
    assert float(((sums - ref).abs() / mag).max()) < 2e-3
    assert float(sums.abs().max()) > 1.0
