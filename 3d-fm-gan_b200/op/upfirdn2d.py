"""upfirdn2d with first- and second-order autograd on ``fm_upfirdn2d``.

API mirror of the reference's op/upfirdn2d.py: ``upfirdn2d(input, kernel, up, down, pad)``
with scalar ``up``/``down`` and a 2-tuple ``pad`` applied to both axes (:154-165), and the
autograd classes ``UpFirDn2d`` / ``UpFirDn2dBackward``.

The adjoint of (up, down, pad, k) is (down, up, g_pad, flip(k)) (reference :40-51,:120-123);
the adjoint of the adjoint is the forward op again, which is what makes R1 and path-length
regularisation (double backward) work without dedicated kernels.
"""
import torch
from torch.autograd import Function

from fm3d import ops


def _out_size(in_size, up, down, pad0, pad1, k):
    return (in_size * up + pad0 + pad1 - k) // down + 1


class UpFirDn2dBackward(Function):
    @staticmethod
    def forward(ctx, grad_output, kernel, grad_kernel, up, down, pad, g_pad, in_size, out_size):
        (ux, uy), (dx, dy) = up, down
        gx0, gx1, gy0, gy1 = g_pad
        g = grad_output.reshape(-1, out_size[0], out_size[1])
        grad_input = ops.upfirdn2d_planes(g, grad_kernel, dx, dy, ux, uy, gx0, gx1, gy0, gy1)
        ctx.save_for_backward(kernel)
        ctx.cfg = (ux, uy, dx, dy) + tuple(pad)
        ctx.in_size, ctx.out_size = tuple(in_size), tuple(out_size)
        return grad_input.view(in_size[0], in_size[1], in_size[2], in_size[3])

    @staticmethod
    def backward(ctx, gradgrad_input):
        (kernel,) = ctx.saved_tensors
        gg = gradgrad_input.reshape(-1, ctx.in_size[2], ctx.in_size[3])
        gg_out = ops.upfirdn2d_planes(gg, kernel, *ctx.cfg)
        gg_out = gg_out.view(ctx.in_size[0], ctx.in_size[1], ctx.out_size[0], ctx.out_size[1])
        return gg_out, None, None, None, None, None, None, None, None


class UpFirDn2d(Function):
    @staticmethod
    def forward(ctx, input, kernel, up, down, pad):
        (ux, uy), (dx, dy) = up, down
        px0, px1, py0, py1 = pad
        kh, kw = kernel.shape
        n, c, in_h, in_w = input.shape
        out_h = _out_size(in_h, uy, dy, py0, py1, kh)
        out_w = _out_size(in_w, ux, dx, px0, px1, kw)
        ctx.in_size, ctx.out_size = tuple(input.shape), (out_h, out_w)
        ctx.up, ctx.down, ctx.pad = (ux, uy), (dx, dy), (px0, px1, py0, py1)
        # padding of the adjoint filter
        ctx.g_pad = (kw - px0 - 1, in_w * ux - out_w * dx + px0 - ux + 1,
                     kh - py0 - 1, in_h * uy - out_h * dy + py0 - uy + 1)
        ctx.save_for_backward(kernel, torch.flip(kernel, [0, 1]))
        out = ops.upfirdn2d_planes(input.reshape(-1, in_h, in_w), kernel, ux, uy, dx, dy, px0, px1, py0, py1)
        return out.view(-1, c, out_h, out_w)

    @staticmethod
    def backward(ctx, grad_output):
        kernel, grad_kernel = ctx.saved_tensors
        grad_input = UpFirDn2dBackward.apply(grad_output, kernel, grad_kernel, ctx.up, ctx.down, ctx.pad,
                                             ctx.g_pad, ctx.in_size, ctx.out_size)
        return grad_input, None, None, None, None


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    if not input.is_cuda:
        raise RuntimeError("upfirdn2d: CUDA tensor required (the B200 path has no CPU fallback)")
    return UpFirDn2d.apply(input, kernel, (up, up), (down, down), (pad[0], pad[1], pad[0], pad[1]))
