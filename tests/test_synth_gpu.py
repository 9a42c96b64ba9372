"""Bandwidth kernels of the synthesis engine (csrc/synth.cu) against the oracle / plain torch fp32
on the same bf16-rounded inputs: the fused blur + noise + bias + leaky-ReLU pass that follows the
stride-2 transposed conv (stylegan2.py:279,312,371) and the ToRGB tail (stylegan2.py:394-399)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import fm_oracle as orc  # noqa: E402  (test infrastructure)


@pytest.mark.parametrize("B,OH,OW,C", [(2, 8, 8, 64), (1, 64, 64, 128), (2, 33, 70, 72), (1, 5, 3, 8), (3, 16, 16, 512)])
@pytest.mark.parametrize("rank1", [True, False])
@pytest.mark.parametrize("per_sample_noise", [True, False])
def test_blur_act_nhwc(cuda, B, OH, OW, C, rank1, per_sample_noise):
    from fm3d import ops
    gen = torch.Generator().manual_seed(B * 1000 + OH * 10 + C)
    cs = (C + 7) // 8 * 8
    t = torch.randn(B, OH + 1, OW + 1, cs, generator=gen).to(torch.bfloat16)
    k = orc.make_kernel_ref([1, 3, 3, 1]) * 4 if rank1 else torch.randn(4, 4, generator=gen) * 0.3
    tab = torch.zeros(B, C, 8)
    tab[..., 0] = torch.rand(B, C, generator=gen) + 0.5        # demodulation
    tab[..., 1] = torch.randn(B, C, generator=gen) * 0.2       # bias
    tab[..., 2] = 0.2                                          # leaky slope
    tab[..., 3] = torch.rand(B, C, generator=gen) + 0.5        # sqrt2 * next style
    noise = torch.randn(B if per_sample_noise else 1, 1, OH, OW, generator=gen)
    nw = torch.tensor([0.37])
    # reference: true convolution with the flipped kernel, pad (1,1)  (op/upfirdn2d.py:168-209)
    x = t[..., :C].float().permute(0, 3, 1, 2)
    blur = orc.upfirdn2d_api_ref(x, k, 1, 1, (1, 1))
    v = blur * tab[..., 0][:, :, None, None] + tab[..., 1][:, :, None, None] + noise * nw
    v = torch.where(v > 0, v, v * 0.2) * tab[..., 3][:, :, None, None]
    out = ops.blur_act_nhwc(t.to(cuda), k.to(cuda), tab.to(cuda), noise.to(cuda), per_sample_noise, nw.to(cuda), C)
    assert out.shape == (B, OH, OW, cs)
    got = out[..., :C].float().permute(0, 3, 1, 2).cpu()
    # bf16 output rounding (2^-8 relative) on fp32-accumulated values
    torch.testing.assert_close(got, v, rtol=1e-2, atol=1e-2)
    # the engine's layout: one spare row / column per image (pitch only, poisoned here: they must never be read)
    tp = torch.full((B, OH + 2, OW + 2, cs), float("nan")).to(torch.bfloat16)
    tp[:, :OH + 1, :OW + 1] = t
    out_p = ops.blur_act_nhwc(tp.to(cuda), k.to(cuda), tab.to(cuda), noise.to(cuda), per_sample_noise, nw.to(cuda), C,
                              padded=True)
    assert torch.equal(out_p, out)


@pytest.mark.parametrize("B,H", [(2, 4), (3, 8), (2, 64), (1, 256)])
@pytest.mark.parametrize("with_skip", [True, False])
def test_rgb_finalize(cuda, B, H, with_skip):
    from fm3d import ops
    gen = torch.Generator().manual_seed(B * 100 + H)
    acc = torch.randn(B, H, H, 4, generator=gen)
    bias = torch.randn(3, generator=gen)
    k = orc.make_kernel_ref([1, 3, 3, 1]) * 4
    skip = torch.randn(B, 3, H // 2, H // 2, generator=gen) if with_skip else None
    ref = acc[..., :3].permute(0, 3, 1, 2) + bias[None, :, None, None]
    if with_skip:
        ref = ref + orc.upfirdn2d_api_ref(skip, k, 2, 1, (2, 1))     # Upsample (stylegan2.py:52-63)
    acc_d = acc.to(cuda)
    out = ops.rgb_finalize(acc_d, bias.to(cuda), skip.to(cuda) if with_skip else None, k.to(cuda) if with_skip else None)
    torch.testing.assert_close(out.cpu(), ref, rtol=1e-5, atol=1e-5)
    assert torch.count_nonzero(acc_d) == 0          # the accumulator is handed back zeroed


@pytest.mark.parametrize("B,H,W", [(1, 4, 4), (3, 64, 64), (2, 256, 256), (2, 6, 10)])
def test_tensor2im_bit_exact(cuda, B, H, W):
    """Device tensor2im vs the reference's numpy recipe (Evaluation/visual_eval.py:24-38): integer output, so the
    bar is bit-exact -- including values outside [-1,1] and values that land exactly on an integer."""
    import numpy as np
    from Evaluation.visual_eval import tensor2im, tensor2im_batch
    gen = torch.Generator().manual_seed(B * 7 + H)
    img = torch.randn(B, 3, H, W, generator=gen) * 0.8
    n = img.numel()
    img.view(-1)[:min(n, 64)] = torch.linspace(-1.25, 1.25, 64)[:min(n, 64)]                 # clipped on both sides
    if n >= 320:
        img.view(-1)[64:320] = (torch.arange(256, dtype=torch.float32) / 127.5) - 1.0    # exact integer preimages
    ref = np.clip(img.numpy(), -1, 1)
    ref = ((np.transpose(ref, (0, 2, 3, 1)) + 1.) * (255. / 2.)).astype(np.uint8)
    out = tensor2im_batch(img.to(cuda))
    assert out.dtype == torch.uint8 and tuple(out.shape) == (B, H, W, 3)
    assert np.array_equal(out.cpu().numpy(), ref)
    one = tensor2im(img.to(cuda))
    assert one.dtype == np.uint8 and np.array_equal(one, ref[0])
