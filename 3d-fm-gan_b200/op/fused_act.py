"""Bias + leaky-ReLU + gain with first- and second-order autograd.

API mirror of the reference's op/fused_act.py (class and function names, argument order,
defaults); the arithmetic runs in ``fm_bias_act`` / ``fm_bias_act_grad_bias``.

Differences from the reference, all deliberate:
  * the bias gradient is reduced inside the gradient kernel (the reference launches a
    separate ``sum`` over the whole tensor, op/fused_act.py:42-48);
  * ``negative_slope`` is honoured (the reference's CPU branch hard-codes 0.2, :119,:125);
  * bf16 is supported; indexing is 64-bit.
"""
import torch
from torch import nn
from torch.autograd import Function

from fm3d import ops

_LRELU, _FWD, _GRAD = 3, 0, 1


class FusedLeakyReLUFunctionBackward(Function):
    """d/dx of lrelu(x+b)*scale, itself differentiable (R1 / path-length double backward)."""

    @staticmethod
    def forward(ctx, grad_output, out, bias, negative_slope, scale):
        ctx.save_for_backward(out)
        ctx.slope, ctx.gain = negative_slope, scale
        if bias:
            grad_input, grad_bias = ops.bias_act_grad_bias(grad_output, out, _LRELU, negative_slope, scale)
        else:
            grad_input = ops.bias_act(grad_output, None, out, _LRELU, _GRAD, negative_slope, scale)
            grad_bias = grad_output.new_empty(0)
        return grad_input, grad_bias

    @staticmethod
    def backward(ctx, gradgrad_input, gradgrad_bias):
        (out,) = ctx.saved_tensors
        # second order: the same gating applied to (gg_input + gg_bias) -- leaky-ReLU has no
        # curvature (the reference's unused grad=2 branch, op/fused_bias_act_kernel.cu:40,44)
        gg_bias = gradgrad_bias if (gradgrad_bias is not None and gradgrad_bias.numel()) else None
        gradgrad_out = ops.bias_act(gradgrad_input, gg_bias, out, _LRELU, _GRAD, ctx.slope, ctx.gain)
        return gradgrad_out, None, None, None, None


class FusedLeakyReLUFunction(Function):
    @staticmethod
    def forward(ctx, input, bias, negative_slope, scale):
        ctx.has_bias = bias is not None
        out = ops.bias_act(input, bias, None, _LRELU, _FWD, negative_slope, scale)
        ctx.save_for_backward(out)
        ctx.slope, ctx.gain = negative_slope, scale
        return out

    @staticmethod
    def backward(ctx, grad_output):
        (out,) = ctx.saved_tensors
        grad_input, grad_bias = FusedLeakyReLUFunctionBackward.apply(
            grad_output, out, ctx.has_bias, ctx.slope, ctx.gain)
        return grad_input, (grad_bias if ctx.has_bias else None), None, None


class FusedLeakyReLU(nn.Module):
    """Holds ``bias[channel]`` (zeros) -- state-dict key ``<name>.bias`` as in the reference."""

    def __init__(self, channel, bias=True, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel)) if bias else None
        self.negative_slope = negative_slope
        self.scale = scale

    def forward(self, input):
        return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)


def fused_leaky_relu(input, bias=None, negative_slope=0.2, scale=2 ** 0.5):
    """leaky_relu(input + bias[None, :, None, ...], negative_slope) * scale on the GPU."""
    if not input.is_cuda:
        raise RuntimeError("fused_leaky_relu: CUDA tensor required (the B200 path has no CPU fallback)")
    return FusedLeakyReLUFunction.apply(input, bias, negative_slope, scale)
