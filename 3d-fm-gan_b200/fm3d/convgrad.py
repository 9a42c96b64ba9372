"""Differentiable convolutions on the tcgen05 kernels: the training path of BASELINE config 4.

The reference differentiates ``F.conv2d`` / ``F.conv_transpose2d`` through ATen (cuDNN dgrad / wgrad:
stylegan2.py:129,276,285,291 under ``loss.backward()``, train_3_encoder.py:476,492,557,593).  Here the three maps a
convolution is made of are ``autograd.Function``s over ``libfm3d.so``:

    ConvFwd(x, w)        y[b,o,oy,ox]  = sum w[o,i,ky,kx] * x[b,i,oy*s+ky-p,ox*s+kx-p]     fm_conv_igemm
    ConvBwdData(g, w)    gx[b,i,y,x]   = sum w[o,i,ky,kx] * g[b,o,oy,ox]  (y = oy*s+ky-p)   fm_conv_igemm, one launch per
                                                                                           output phase (y+p mod s)
    ConvBwdWeight(x, g)  gw[o,i,ky,kx] = sum x[b,i,oy*s+ky-p,ox*s+kx-p] * g[b,o,oy,ox]     fm_conv_wgrad (GEMM over pixels)

Each is bilinear and the derivative of each is again one of the three ("gradfix" pattern), so
second-order passes -- R1 through the discriminator (Util/training_util.py:46-52) and path-length regularisation
through the generator (stylegan2.py:683-688) -- fall out of autograd with no further kernels:

    d ConvFwd       : dx = ConvBwdData(gy, w)      dw = ConvBwdWeight(x, gy)
    d ConvBwdData   : dg = ConvFwd(ggx, w)         dw = ConvBwdWeight(ggx, g)
    d ConvBwdWeight : dx = ConvBwdData(g, ggw)     dg = ConvFwd(x, ggw)

``conv_transpose2d(x, w)`` (the generator's stride-2 up-conv, stylegan2.py:276) IS ConvBwdData(x, w).

Tensors at this boundary are NCHW fp32 (the autograd graph of the mirrored modules); operands are rounded to bf16,
products accumulate in fp32 in TMEM (the precision of the whole B200 path).  CUDA only.
"""
import ctypes as C

import torch
from torch.autograd import Function

from . import _lib, ops
from ._lib import WgradDesc

_IDENT_TABS = {}


def _ident_tab(cout, device):
    key = (cout, device)
    t = _IDENT_TABS.get(key)
    if t is None:
        t = torch.zeros(1, cout, 8, device=device, dtype=torch.float32)
        t[..., 0] = 1.0
        t[..., 2] = 1.0
        t[..., 3] = 1.0
        _IDENT_TABS[key] = t
    return t


def _out_size(h, k, s, p):
    return (h + 2 * p - k) // s + 1


# ---------------------------------------------------------------------------------------------- operand memo
# Every map converts its NCHW fp32 operands to the kernels' layout (NHWC bf16 activations, K-major bf16 weights).  The same
# tensor is needed more than once: the upstream gradient by dgrad AND wgrad of one layer, the layer input by forward AND
# wgrad, a weight by several passes.  In a training iteration the conversion pass was the most-launched kernel (21 % of the
# GPU time), so converted operands are memoised: keyed on the tensor's storage address, version counter and geometry, with a
# strong reference to the source tensor so that the address cannot be recycled while the entry lives.  Small and bounded
# (activations: the last few; weights: the last few dozen), cleared by ``clear_memo()``.
class _Memo:
    def __init__(self, size):
        self.size, self.items = size, {}

    @staticmethod
    def key(t, tag):
        return (t.data_ptr(), t._version, tuple(t.shape), tuple(t.stride()), t.device.index, tag)

    def get(self, t, tag):
        e = self.items.get(self.key(t, tag))
        return None if e is None else e[1]

    def put(self, t, tag, val):
        if len(self.items) >= self.size:
            self.items.pop(next(iter(self.items)))          # oldest entry
        self.items[self.key(t, tag)] = (t, val)
        return val


_ACT_MEMO = _Memo(4)
_W_MEMO = _Memo(96)


def clear_memo():
    _ACT_MEMO.items.clear()
    _W_MEMO.items.clear()


def _nhwc(t):
    """fp32 NCHW -> bf16 NHWC (memoised)."""
    v = _ACT_MEMO.get(t, "nhwc")
    return v if v is not None else _ACT_MEMO.put(t, "nhwc", ops.nchw_to_nhwc_bf16(t))


def _wq(w, transposed):
    """fp32 [O,I,k,k] -> bf16 [k*k][O][I] (or [k*k][I][O] for the data gradient), memoised."""
    tag = "wT" if transposed else "w"
    v = _W_MEMO.get(w, tag)
    if v is None:
        v = _W_MEMO.put(w, tag, ops.prep_weight(w.transpose(0, 1) if transposed else w, 1.0, want_wsq=False)[0])
    return v


# ---------------------------------------------------------------------------------------------- kernels
def conv_forward(x, w, stride, pad):
    """x [B,I,H,W] fp32, w [O,I,k,k] fp32 -> [B,O,OH,OW] fp32."""
    B, I, H, W = x.shape
    O, _, k, _ = w.shape
    OH, OW = _out_size(H, k, stride, pad), _out_size(W, k, stride, pad)
    out = torch.empty(B, O, OH, OW, device=x.device, dtype=torch.float32)
    if out.numel() == 0:
        return out
    xq = _nhwc(x)
    wq = _wq(w, False)
    ops.conv_igemm(xq, wq, ops.conv_taps(k, k, pad), out, _ident_tab(O, x.device), B=B, H=H, W=W, Cin=I, Cout=O,
                   OH=OH, OW=OW, stride=stride, out_nchw_f32=True)
    return out


def conv_backward_data(g, w, stride, pad, in_hw):
    """g [B,O,OH,OW], w [O,I,k,k] -> gx [B,I,H,W] (H,W = in_hw): the adjoint of conv_forward w.r.t. x, i.e. the
    transposed convolution.  Phase r = (y + pad) mod stride of the output takes the taps ky = r (mod stride); each
    phase is one plain conv over the grid q (y = stride*q' + y0) written with output stride ``stride``."""
    B, O, OH, OW = g.shape
    _, I, k, _ = w.shape
    H, W = in_hw
    s = stride
    multi = s > 1
    out = (torch.zeros if multi else torch.empty)(B, I, H, W, device=g.device, dtype=torch.float32)
    if out.numel() == 0 or g.numel() == 0:
        return out.zero_()
    gq = _nhwc(g)
    wq = _wq(w, True)                                                     # [k*k][I rows][O]: K dimension = g's channels
    tab = _ident_tab(I, g.device)

    def phase_1d(r, n_in, n_out):
        """-> (y0, rows, [(d, kidx)]) for residue r: outputs y = s*q' + y0, q' in [0, rows); tap j reads g[q' + d]."""
        y0 = (r - pad) % s
        qshift = (y0 - (r - pad)) // s
        rows = (n_out - y0 + s - 1) // s if n_out > y0 else 0
        taps = [(qshift - j, r + s * j) for j in range((k - r + s - 1) // s)] if r < k else []
        return y0, rows, taps
    for ry in range(s):
        y0, rows, ty = phase_1d(ry, OH, H)
        for rx in range(s):
            x0, cols, tx = phase_1d(rx, OW, W)
            if rows <= 0 or cols <= 0:
                continue
            taps = [(dy, dx, ky * k + kx) for (dy, ky) in ty for (dx, kx) in tx]
            if not taps:
                continue                                  # this phase receives no contribution: stays zero
            ops.conv_igemm(gq, wq, taps, out, tab, B=B, H=OH, W=OW, Cin=O, Cout=I, OH=rows, OW=cols, out_H=H, out_W=W,
                           out_y0=y0, out_x0=x0, out_ys=s, out_xs=s, out_nchw_f32=True,
                           algo_flops=2.0 * B * OH * OW * O * I * len(taps))
    return out


def conv_wgrad(a, b, Ca, Cb, B, GH, GW, taps, stride_a=1, stride_b=1, ksplit=0):
    """fm_conv_wgrad: a, b bf16 NHWC [B,H,W,cs]; taps = [(dy_a, dx_a, dy_b, dx_b)] -> dw fp32 [ntaps, Ca, Cb]."""
    dw = torch.zeros(len(taps), Ca, (Cb + 3) // 4 * 4, device=a.device, dtype=torch.float32)
    d = WgradDesc()
    for o, t, c, s in ((d.a, a, Ca, stride_a), (d.b, b, Cb, stride_b)):
        o.ptr, o.C, o.cstride, o.H, o.W, o.stride = t.data_ptr(), c, t.shape[3], t.shape[1], t.shape[2], s
    d.B, d.GH, d.GW, d.ntaps = B, GH, GW, len(taps)
    for i, (dya, dxa, dyb, dxb) in enumerate(taps):
        d.tap_dy_a[i], d.tap_dx_a[i], d.tap_dy_b[i], d.tap_dx_b[i] = dya, dxa, dyb, dxb
    d.dw, d.dw_tap_stride, d.dw_row_stride, d.ksplit = dw.data_ptr(), dw.shape[1] * dw.shape[2], dw.shape[2], ksplit
    with torch.cuda.device(a.device):
        _lib.check(_lib.lib().fm_conv_wgrad(C.byref(d), torch.cuda.current_stream().cuda_stream), "fm_conv_wgrad")
    return dw[..., :Cb]


def conv_backward_weight(x, g, stride, pad, k):
    """x [B,I,H,W], g [B,O,OH,OW] -> gw [O,I,k,k]: one GEMM over the B*OH*OW output pixels per tap (csrc/wgrad.cu).  Both
    operands are read as NHWC bf16 -- the layout the forward / dgrad kernels use; tap (ky,kx) shifts x by (ky-p, kx-p)."""
    B, I, H, W = x.shape
    _, O, OH, OW = g.shape
    xq = _nhwc(x)
    gq = _nhwc(g)
    shifts = [(ky - pad, kx - pad) for ky in range(k) for kx in range(k)]
    if O >= I:          # the larger channel count takes the 128-row M side, the smaller one the N side (>= 16 wide)
        dw = conv_wgrad(gq, xq, O, I, B, OH, OW, [(0, 0, dy, dx) for (dy, dx) in shifts], 1, stride)      # [t][o][i]
        return dw.permute(1, 2, 0).reshape(O, I, k, k).contiguous()
    dw = conv_wgrad(xq, gq, I, O, B, OH, OW, [(dy, dx, 0, 0) for (dy, dx) in shifts], stride, 1)          # [t][i][o]
    return dw.permute(2, 1, 0).reshape(O, I, k, k).contiguous()


# ---------------------------------------------------------------------------------------------- autograd
class ConvFwd(Function):
    @staticmethod
    def forward(ctx, x, w, stride, pad):
        ctx.save_for_backward(x, w)
        ctx.cfg = (stride, pad)
        xd = x.detach()
        y = conv_forward(xd, w.detach(), stride, pad)
        # the bf16 NHWC copy of the input is what the weight gradient reads: keep it with the graph instead of converting
        # the saved fp32 tensor again in backward (+50 % of this layer's saved activation bytes, one conversion pass less)
        ctx.xq = _ACT_MEMO.get(xd, "nhwc") if w.requires_grad else None
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        stride, pad = ctx.cfg
        gx = gw = None
        if ctx.needs_input_grad[0]:
            gx = ConvBwdData.apply(gy, w, stride, pad, (x.shape[2], x.shape[3]))
        if ctx.needs_input_grad[1]:
            if ctx.xq is not None and _ACT_MEMO.get(x.detach(), "nhwc") is None:
                _ACT_MEMO.put(x.detach(), "nhwc", ctx.xq)
            gw = ConvBwdWeight.apply(x, gy, stride, pad, w.shape[2])
        return gx, gw, None, None


class ConvBwdData(Function):
    @staticmethod
    def forward(ctx, g, w, stride, pad, in_hw):
        ctx.save_for_backward(g, w)
        ctx.cfg = (stride, pad)
        return conv_backward_data(g.detach(), w.detach(), stride, pad, tuple(in_hw))

    @staticmethod
    def backward(ctx, ggx):
        g, w = ctx.saved_tensors
        stride, pad = ctx.cfg
        gg = gw = None
        if ctx.needs_input_grad[0]:
            gg = ConvFwd.apply(ggx, w, stride, pad)
            if gg.shape != g.shape:
                # rows / columns of g the strided window never reached (H not = (OH-1)*s + k - 2p): zero gradient there
                full = gg.new_zeros(g.shape)
                full[:, :, :gg.shape[2], :gg.shape[3]] = gg[:, :, :g.shape[2], :g.shape[3]]
                gg = full
        if ctx.needs_input_grad[1]:
            gw = ConvBwdWeight.apply(ggx, g, stride, pad, w.shape[2])
        return gg, gw, None, None, None


class ConvBwdWeight(Function):
    @staticmethod
    def forward(ctx, x, g, stride, pad, k):
        ctx.save_for_backward(x, g)
        ctx.cfg = (stride, pad)
        return conv_backward_weight(x.detach(), g.detach(), stride, pad, k)

    @staticmethod
    def backward(ctx, ggw):
        x, g = ctx.saved_tensors
        stride, pad = ctx.cfg
        gx = gg = None
        if ctx.needs_input_grad[0]:
            gx = ConvBwdData.apply(g, ggw, stride, pad, (x.shape[2], x.shape[3]))
        if ctx.needs_input_grad[1]:
            gg = ConvFwd.apply(x, ggw, stride, pad)
            if gg.shape != g.shape:
                full = gg.new_zeros(g.shape)
                full[:, :, :gg.shape[2], :gg.shape[3]] = gg[:, :, :g.shape[2], :g.shape[3]]
                gg = full
        return gx, gg, None, None, None


class ChannelScale(Function):
    """y = x * s[:, :, None, None] -- the modulation (input channels times style) and demodulation (output channels times
    d) of ModulatedConv2d's shared-weight composition (stylegan2.py:250-298).  As a broadcast ``x * s.view(B, C, 1, 1)`` it
    ran on ATen's non-vectorised elementwise kernel and its gradient on a broadcast multiply plus a reduction: 201 + ~280
    launches, ~30 ms of an iteration.  Together with ChannelDot it is closed under differentiation (each one's gradient is
    the other), so R1 / path-length double backward go through the same two kernels."""

    @staticmethod
    def forward(ctx, x, s):
        ctx.save_for_backward(x, s)
        return ops.channel_scale(x.detach(), s.detach())

    @staticmethod
    def backward(ctx, g):
        x, s = ctx.saved_tensors
        gx = ChannelScale.apply(g, s) if ctx.needs_input_grad[0] else None
        gs = ChannelDot.apply(g, x).reshape(s.shape) if ctx.needs_input_grad[1] else None
        return gx, gs


class ChannelDot(Function):
    """r[b, c] = sum over the trailing dims of a * b."""

    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a, b)
        return ops.channel_dot(a.detach(), b.detach())

    @staticmethod
    def backward(ctx, gr):
        a, b = ctx.saved_tensors
        ga = ChannelScale.apply(b, gr) if ctx.needs_input_grad[0] else None
        gb = ChannelScale.apply(a, gr) if ctx.needs_input_grad[1] else None
        return ga, gb


class BiasAdd(Function):
    """y + bias[None, :, None, None] on the fused bias-act kernel in its linear mode (``y + bias.view(1, -1, 1, 1)`` runs on
    ATen's non-vectorised broadcast kernel; the VGG16 convolutions of the LPIPS term add their biases to 256^2 x 64 maps).
    The gradient is the identity and a sum: torch ops, differentiable again."""

    @staticmethod
    def forward(ctx, y, bias):
        return ops.bias_act(y.detach(), bias.detach(), None, 1, 0, 0.0, 1.0)

    @staticmethod
    def backward(ctx, g):
        return (g if ctx.needs_input_grad[0] else None), (g.sum(dim=(0, 2, 3)) if ctx.needs_input_grad[1] else None)


class PlaneAdd(Function):
    """image [B, C, H, W] + plane [B or 1, 1, H, W] (NoiseInjection, stylegan2.py:312): the gradient is the identity and a
    sum over the channels (and the batch for a shared plane) -- torch ops, differentiable again."""

    @staticmethod
    def forward(ctx, x, n):
        ctx.n_shape = n.shape
        return ops.plane_add(x.detach(), n.detach())

    @staticmethod
    def backward(ctx, g):
        gn = None
        if ctx.needs_input_grad[1]:
            gn = g.sum(dim=1, keepdim=True)
            if ctx.n_shape[0] == 1 and g.shape[0] > 1:
                gn = gn.sum(dim=0, keepdim=True)
        return (g if ctx.needs_input_grad[0] else None), gn


def plane_add(x, n):
    return PlaneAdd.apply(x, n)


def channel_scale(x, s):
    """x [B, C, H, W] * s [B, C] (any shape with B * C elements), differentiable to any order on the native kernels."""
    return ChannelScale.apply(x, s.reshape(x.shape[0], x.shape[1]))


# ---------------------------------------------------------------------------------------------- public
def _check(x, w):
    if not (x.is_cuda and w.is_cuda):
        raise RuntimeError("fm3d.convgrad: CUDA tensors required (the B200 path has no CPU fallback)")
    if x.ndim != 4 or w.ndim != 4 or w.shape[2] != w.shape[3]:
        raise RuntimeError(f"fm3d.convgrad: expected x [B,I,H,W] and a square kernel, got {tuple(x.shape)} / {tuple(w.shape)}")


def conv2d(x, w, bias=None, stride=1, padding=0):
    """``F.conv2d(x, w, bias, stride, padding)`` (square kernels, symmetric stride / padding, groups = 1)."""
    _check(x, w)
    if x.shape[1] != w.shape[1]:
        raise RuntimeError(f"conv2d: input has {x.shape[1]} channels, weight expects {w.shape[1]}")
    y = ConvFwd.apply(x.float(), w.float(), int(stride), int(padding))
    if bias is not None:
        y = BiasAdd.apply(y, bias.float())
    return y


def conv_transpose2d(x, w, stride=1, padding=0):
    """``F.conv_transpose2d(x, w, stride=stride, padding=padding)`` with w [I,O,k,k]."""
    _check(x, w)
    if x.shape[1] != w.shape[0]:
        raise RuntimeError(f"conv_transpose2d: input has {x.shape[1]} channels, weight expects {w.shape[0]}")
    k, s, p = w.shape[2], int(stride), int(padding)
    out_hw = ((x.shape[2] - 1) * s + k - 2 * p, (x.shape[3] - 1) * s + k - 2 * p)
    return ConvBwdData.apply(x.float(), w.float(), s, p, out_hw)


# ---------------------------------------------------------------------------------------------- nn.Conv2d on the native kernels
class NativeConv2d(torch.nn.Conv2d):
    """``nn.Conv2d`` whose forward runs on the tcgen05 kernels (forward, dgrad, wgrad; closed under double backward).
    Instances are made by re-classing existing modules (``use_native_convs``): parameters, buffers, state-dict keys and
    ``isinstance(m, nn.Conv2d)`` are unchanged, and copies / pickles of the module keep working."""

    def forward(self, x):
        import os
        if x.is_cuda and x.dtype == torch.float32 and x.ndim == 4 and os.environ.get("FM3D_NATIVE_GRAD", "1") != "0":
            return conv2d(x, self.weight, self.bias, self.stride[0], self.padding[0])
        return super().forward(x)


def use_native_convs(module):
    """Route every eligible ``nn.Conv2d`` inside ``module`` (square kernel up to 7x7, symmetric stride 1 / 2 and zero
    padding, no dilation, groups = 1) through ``convgrad.conv2d``.  Used for the encoders in training
    (resnet_encoder.py:36,42,193; helpers.py:80-131; psp_encoders.py:27-31) and for frozen loss networks on the gradient
    path of the image (LPIPS-VGG16 lpips/networks_basic.py:36-101, ArcFace Util/arcface_pytorch: SURVEY 8f rank 2).
    Returns the number of convolutions switched."""
    n = 0
    for m in module.modules():
        if type(m) is torch.nn.Conv2d:
            ok = (m.groups == 1 and m.dilation == (1, 1) and m.kernel_size[0] == m.kernel_size[1] and
                  m.stride[0] == m.stride[1] and m.stride[0] in (1, 2) and isinstance(m.padding, tuple) and
                  m.padding[0] == m.padding[1] and m.padding_mode == "zeros" and m.kernel_size[0] <= 7)
            if ok:
                m.__class__ = NativeConv2d
                n += 1
    return n
