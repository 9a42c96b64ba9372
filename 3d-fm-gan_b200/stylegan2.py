"""StyleGAN2 generator / discriminator for 3D-FM GAN on B200.

Drop-in for the reference's ``stylegan2.py``: identical class names, constructor
signatures, attribute names and state-dict keys/shapes (checkpoints load either way),
identical ``Generator.forward`` keyword arguments and return structures.

What is different underneath:

* every module runs on the hand-written sm_100a kernels of ``libfm3d.so``
  (``op.fused_leaky_relu``, ``op.upfirdn2d``, and for the whole synthesis network the
  tcgen05 implicit-GEMM engine in ``fm3d.engine``);
* the modulated convolution never materialises per-sample weights.  It uses the
  equivalent form  y = d[b,o] * conv(x * s[b,i], W/sqrt(fan_in))  with
  d = rsqrt(sum_i s^2 * sum_k (W/sqrt(fan_in))^2 + 1e-8)  (reference math:
  stylegan2.py:257-262; equality up to fp32 rounding), so the batch is a
  plain GEMM M dimension and the weight gradient is one ordinary conv-wgrad;
* ``Generator.forward`` dispatches gradient-free CUDA calls to the fused bf16 engine
  (``FM3D_ENGINE=0`` disables it); anything that needs autograd (training, path-length
  regularisation, R1) takes the differentiable fp32 composition below -- still CUDA only.
"""
import math
import os
import random

import torch
from torch import autograd, nn
from torch.nn import functional as F

from op import FusedLeakyReLU, fused_leaky_relu, upfirdn2d

_SQRT2 = math.sqrt(2.0)


def _native_grad(x):
    """Differentiable convolutions run on the tcgen05 kernels (fm3d/convgrad.py: forward, dgrad and wgrad, all closed
    under double backward).  ``FM3D_NATIVE_GRAD=0`` keeps ATen's convolutions instead -- the fp32 cross-check of the
    tests, not a product path."""
    return x.is_cuda and os.environ.get("FM3D_NATIVE_GRAD", "1") != "0"


def _convgrad():
    from fm3d import convgrad
    return convgrad


def _conv2d(x, w, bias=None, stride=1, padding=0):
    if _native_grad(x):
        from fm3d import convgrad
        return convgrad.conv2d(x, w, bias, stride, padding)
    return F.conv2d(x, w, bias=bias, stride=stride, padding=padding)


def _conv_transpose2d(x, w, stride):
    if _native_grad(x):
        from fm3d import convgrad
        return convgrad.conv_transpose2d(x, w, stride=stride)
    return F.conv_transpose2d(x, w, padding=0, stride=stride)


# --------------------------------------------------------------------------- small layers
class PixelNorm(nn.Module):
    """x / sqrt(mean_c(x^2) + 1e-8) over dim 1 (vectors [N,D] or maps [N,D,H,W])."""

    def forward(self, input):
        return input * torch.rsqrt(input.square().mean(dim=1, keepdim=True) + 1e-8)


def make_kernel(k):
    """1-D taps -> separable 2-D FIR, normalised to unit DC gain."""
    k = torch.tensor(k, dtype=torch.float32)
    if k.ndim == 1:
        k = torch.outer(k, k)
    return k / k.sum()


def _resample_pad(taps, factor, extra):
    """(pad0, pad1) for a length-``taps`` FIR around a x``factor`` resampler."""
    p = taps - factor + extra
    return p


class Upsample(nn.Module):
    def __init__(self, kernel, factor=2):
        super().__init__()
        self.factor = factor
        self.register_buffer('kernel', make_kernel(kernel) * (factor ** 2))
        p = self.kernel.shape[0] - factor
        self.pad = ((p + 1) // 2 + factor - 1, p // 2)

    def forward(self, input):
        return upfirdn2d(input, self.kernel, up=self.factor, down=1, pad=self.pad)


class Downsample(nn.Module):
    def __init__(self, kernel, factor=2):
        super().__init__()
        self.factor = factor
        self.register_buffer('kernel', make_kernel(kernel))
        p = self.kernel.shape[0] - factor
        self.pad = ((p + 1) // 2, p // 2)

    def forward(self, input):
        return upfirdn2d(input, self.kernel, up=1, down=self.factor, pad=self.pad)


class Blur(nn.Module):
    def __init__(self, kernel, pad, upsample_factor=1):
        super().__init__()
        taps = make_kernel(kernel)
        if upsample_factor > 1:
            taps = taps * (upsample_factor ** 2)
        self.register_buffer('kernel', taps)
        self.pad = pad

    def forward(self, input):
        return upfirdn2d(input, self.kernel, pad=self.pad)


class EqualConv2d(nn.Module):
    """Conv2d with the equalised-learning-rate runtime scale 1/sqrt(fan_in)."""

    def __init__(self, in_channel, out_channel, kernel_size, stride=1, padding=0, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_channel, in_channel, kernel_size, kernel_size))
        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)
        self.stride = stride
        self.padding = padding
        self.bias = nn.Parameter(torch.zeros(out_channel)) if bias else None

    def forward(self, input):
        return _conv2d(input, self.weight * self.scale, bias=self.bias, stride=self.stride, padding=self.padding)

    def __repr__(self):
        o, i, k, _ = self.weight.shape
        return f'{self.__class__.__name__}({i}, {o}, {k}, stride={self.stride}, padding={self.padding})'


class EqualLinear(nn.Module):
    def __init__(self, in_dim, out_dim, bias=True, bias_init=0, lr_mul=1, activation=None):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_dim, in_dim).div_(lr_mul))
        self.bias = nn.Parameter(torch.zeros(out_dim).fill_(bias_init)) if bias else None
        self.activation = activation
        self.scale = (1 / math.sqrt(in_dim)) * lr_mul
        self.lr_mul = lr_mul

    def forward(self, input):
        w = self.weight * self.scale
        if self.activation:
            return fused_leaky_relu(F.linear(input, w), self.bias * self.lr_mul)
        return F.linear(input, w, bias=self.bias * self.lr_mul)

    def __repr__(self):
        return f'{self.__class__.__name__}({self.weight.shape[1]}, {self.weight.shape[0]})'


class ScaledLeakyReLU(nn.Module):
    def __init__(self, negative_slope=0.2):
        super().__init__()
        self.negative_slope = negative_slope

    def forward(self, input):
        return F.leaky_relu(input, negative_slope=self.negative_slope) * _SQRT2


# --------------------------------------------------------------------------- modulated conv
class ModulatedConv2d(nn.Module):
    def __init__(self, in_channel, out_channel, kernel_size, style_dim, demodulate=True,
                 upsample=False, downsample=False, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        self.eps = 1e-8
        self.kernel_size = kernel_size
        self.in_channel = in_channel
        self.out_channel = out_channel
        self.upsample = upsample
        self.downsample = downsample

        if upsample:
            p = (len(blur_kernel) - 2) - (kernel_size - 1)
            self.blur = Blur(blur_kernel, pad=((p + 1) // 2 + 1, p // 2 + 1), upsample_factor=2)
        if downsample:
            p = (len(blur_kernel) - 2) + (kernel_size - 1)
            self.blur = Blur(blur_kernel, pad=((p + 1) // 2, p // 2))

        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)
        self.padding = kernel_size // 2
        self.weight = nn.Parameter(torch.randn(1, out_channel, in_channel, kernel_size, kernel_size))
        self.modulation = EqualLinear(style_dim, in_channel, bias_init=1)
        self.demodulate = demodulate

    def __repr__(self):
        return (f'{self.__class__.__name__}({self.in_channel}, {self.out_channel}, {self.kernel_size}, '
                f'upsample={self.upsample}, downsample={self.downsample})')

    def forward(self, input, style, return_style_scalars=False):
        """input [B,I,H,W], style [B,style_dim] -> [B,O,H',W'] (differentiable composition).

        Shared-weight form of the reference's per-sample grouped conv: scale the input
        channels by s, convolve once for the whole batch, scale the output channels by d."""
        batch = input.shape[0]
        s = self.modulation(style)                                        # [B, I]
        w = self.weight[0] * self.scale                                   # [O, I, k, k]
        native = _native_grad(input) and input.dtype == torch.float32 and s.dtype == torch.float32
        x = _convgrad().channel_scale(input, s) if native else input * s.view(batch, self.in_channel, 1, 1)

        if self.upsample:
            out = _conv_transpose2d(x, w.transpose(0, 1), stride=2)
            out = self.blur(out)
        elif self.downsample:
            out = _conv2d(self.blur(x), w, padding=0, stride=2)
        else:
            out = _conv2d(x, w, padding=self.padding)

        if self.demodulate:
            wsq = w.square().sum(dim=(2, 3))                              # [O, I]
            d = torch.rsqrt(F.linear(s.square(), wsq) + 1e-8)             # literal 1e-8, not self.eps
            out = _convgrad().channel_scale(out, d) if native else out * d.view(batch, self.out_channel, 1, 1)

        if return_style_scalars:
            return out, s.view(batch, 1, self.in_channel, 1, 1)
        return out


class NoiseInjection(nn.Module):
    def __init__(self):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(1))

    def forward(self, image, noise=None):
        if noise is None:
            n, _, h, w = image.shape
            noise = image.new_empty(n, 1, h, w).normal_()
        if (_native_grad(image) and image.dtype == torch.float32 and image.ndim == 4 and noise.ndim == 4 and noise.shape[1] == 1
                and noise.shape[2:] == image.shape[2:] and noise.shape[0] in (1, image.shape[0])):
            return _convgrad().plane_add(image, self.weight * noise)
        return image + self.weight * noise


class ConstantInput(nn.Module):
    """Learned 4x4 start tensor, repeated to the batch size of the latent it is given."""

    def __init__(self, channel, size=4):
        super().__init__()
        self.input = nn.Parameter(torch.randn(1, channel, size, size))

    def forward(self, input):
        return self.input.repeat(input.shape[0], 1, 1, 1)


class StyledConv(nn.Module):
    def __init__(self, in_channel, out_channel, kernel_size, style_dim, upsample=False,
                 blur_kernel=[1, 3, 3, 1], demodulate=True):
        super().__init__()
        self.conv = ModulatedConv2d(in_channel, out_channel, kernel_size, style_dim, upsample=upsample,
                                    blur_kernel=blur_kernel, demodulate=demodulate)
        self.noise = NoiseInjection()
        self.activate = FusedLeakyReLU(out_channel)

    def forward(self, input, style, return_style_scalars=False, noise=None):
        res = self.conv(input, style, return_style_scalars)
        out, styles = res if return_style_scalars else (res, None)
        out = self.activate(self.noise(out, noise=noise))
        return (out, styles) if return_style_scalars else out


class ToRGB(nn.Module):
    def __init__(self, in_channel, style_dim, upsample=True, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        if upsample:
            self.upsample = Upsample(blur_kernel)
        self.conv = ModulatedConv2d(in_channel, 3, 1, style_dim, demodulate=False)
        self.bias = nn.Parameter(torch.zeros(1, 3, 1, 1))

    def forward(self, input, style, skip=None, return_style_scalars=False):
        res = self.conv(input, style, return_style_scalars)
        out, styles = res if return_style_scalars else (res, None)
        out = out + self.bias
        if skip is not None:
            out = out + self.upsample(skip)
        return (out, styles) if return_style_scalars else out


# --------------------------------------------------------------------------- generator
def _channel_table(channel_multiplier):
    return {4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * channel_multiplier, 128: 128 * channel_multiplier,
            256: 64 * channel_multiplier, 512: 32 * channel_multiplier, 1024: 16 * channel_multiplier}


class Generator(nn.Module):
    def __init__(self, size, style_dim, n_mlp, channel_multiplier=2, blur_kernel=[1, 3, 3, 1], lr_mlp=0.01,
                 generator_net_shape=None):
        """``generator_net_shape``: optional per-layer channel list of a pruned generator,
        [input, conv1, (up_i, conv_i) ...]; ToRGB widths follow the odd entries."""
        super().__init__()
        self.size = size
        self.style_dim = style_dim

        mlp = [PixelNorm()]
        mlp += [EqualLinear(style_dim, style_dim, lr_mul=lr_mlp, activation='fused_lrelu') for _ in range(n_mlp)]
        self.style = nn.Sequential(*mlp)

        self.channels = _channel_table(channel_multiplier)
        self.log_size = int(math.log(size, 2))
        self.num_layers = (self.log_size - 2) * 2 + 1

        if generator_net_shape is None:
            widths = [self.channels[4], self.channels[4]]
            for i in range(3, self.log_size + 1):
                widths += [self.channels[2 ** i]] * 2
        else:
            widths = list(generator_net_shape)

        self.input = ConstantInput(widths[0])
        self.conv1 = StyledConv(widths[0], widths[1], 3, style_dim, blur_kernel=blur_kernel)
        self.to_rgb1 = ToRGB(widths[1], style_dim, upsample=False)

        self.convs = nn.ModuleList()
        self.upsamples = nn.ModuleList()
        self.to_rgbs = nn.ModuleList()
        self.noises = nn.Module()
        for layer_idx in range(self.num_layers):
            res = 2 ** ((layer_idx + 5) // 2)
            self.noises.register_buffer(f'noise_{layer_idx}', torch.randn(1, 1, res, res))

        for j in range(1, len(widths) // 2):
            c_in, c_mid, c_out = widths[2 * j - 1], widths[2 * j], widths[2 * j + 1]
            self.convs.append(StyledConv(c_in, c_mid, 3, style_dim, upsample=True, blur_kernel=blur_kernel))
            self.convs.append(StyledConv(c_mid, c_out, 3, style_dim, blur_kernel=blur_kernel))
            self.to_rgbs.append(ToRGB(c_out, style_dim))

        self.n_latent = self.log_size * 2 - 2
        self._engine_epoch = 0

    # -- engine cache control (fm3d/engine.py: SynthesisPlan._refresh_weights)
    def invalidate_engine(self):
        """Tell the fused engine that parameter VALUES may have changed behind autograd's back (writes through
        ``.data`` / ``.detach()`` views do not move the version counters the engine watches).  Cheap: the derived bf16
        weights are recomputed in place on the next gradient-free forward; captured CUDA graphs stay valid."""
        self._engine_epoch += 1

    def named_parameters(self, *args, **kwargs):
        # whoever asks for parameter handles may write through ``.data`` (the reference's EMA does, every iteration:
        # ``accumulate``, train_3_encoder.py:195-200): treat the derived weights as stale from here on
        self._engine_epoch += 1
        return super().named_parameters(*args, **kwargs)

    # -- helpers also present on the reference class
    def make_noise(self):
        device = self.input.input.device
        noises = [torch.randn(1, 1, 4, 4, device=device)]
        for i in range(3, self.log_size + 1):
            noises += [torch.randn(1, 1, 2 ** i, 2 ** i, device=device) for _ in range(2)]
        return noises

    def mean_latent(self, n_latent):
        z = torch.randn(n_latent, self.style_dim, device=self.input.input.device)
        return self.style(z).mean(0, keepdim=True)

    def get_latent(self, input):
        return self.style(input)

    # -- forward pieces
    def _assemble_latent(self, styles, inject_index):
        """One W (or W+) -> [B, n_latent, D]; two Ws -> style mixing at ``inject_index``."""
        if len(styles) < 2:
            w = styles[0]
            return w.unsqueeze(1).repeat(1, self.n_latent, 1) if w.ndim < 3 else w
        if inject_index is None:
            inject_index = random.randint(1, self.n_latent - 1)
        head = styles[0].unsqueeze(1).repeat(1, inject_index, 1)
        tail = styles[1].unsqueeze(1).repeat(1, self.n_latent - inject_index, 1)
        return torch.cat([head, tail], 1)

    def _engine_ok(self, latent, start, noise, return_style_scalars, PPL_regularize):
        if os.environ.get("FM3D_ENGINE", "1") == "0" or return_style_scalars or PPL_regularize:
            return False
        if not (latent.is_cuda and latent.dtype == torch.float32):
            return False
        if torch.is_grad_enabled() and (latent.requires_grad or start.requires_grad or
                                        any(p.requires_grad for _, p in nn.Module.named_parameters(self))):
            return False
        if noise is not None and any(n is not None and n.requires_grad for n in noise):
            return False
        from fm3d.engine import check_inputs
        return check_inputs(self, latent, start, noise)      # raises on shapes the reference would reject too

    def forward(self, noise_z, return_latents=False, inject_index=None, truncation=1, truncation_latent=None,
                latent_styles=None, input_is_latent=False, noise=None, randomize_noise=True,
                use_external_input_tensor=False, external_input_tensor=None, PPL_regularize=False,
                return_rgb_list=False, return_style_scalars=False):
        """Same contract as the reference (stylegan2.py:554-688).

        input_is_latent / latent_styles: skip the mapping MLP and take W (or W+) directly;
        use_external_input_tensor: replace the learned 4x4 constant (3-encoder path);
        PPL_regularize: return (image, path_lengths); return_rgb_list: per-resolution RGBs;
        return_style_scalars: also return the per-layer modulation scalars."""
        styles = latent_styles if input_is_latent else [self.style(z) for z in noise_z]

        if noise is None:
            if randomize_noise:
                noise = [None] * self.num_layers
            else:
                noise = [getattr(self.noises, f'noise_{i}') for i in range(self.num_layers)]

        if truncation < 1:
            styles = [truncation_latent + truncation * (w - truncation_latent) for w in styles]

        latent = self._assemble_latent(styles, inject_index)

        if use_external_input_tensor:
            assert external_input_tensor is not None
            start = external_input_tensor
        else:
            start = self.input(latent)

        if self._engine_ok(latent, start, noise, return_style_scalars, PPL_regularize):
            from fm3d.engine import run_synthesis
            rgbs = run_synthesis(self, latent, start, noise)
            return rgbs if return_rgb_list else rgbs[-1]

        # ---- differentiable per-module composition
        styles_list = []

        def call(mod, *a, **k):
            if return_style_scalars:
                out, sc = mod(*a, return_style_scalars=True, **k)
                styles_list.append(sc)
                return out
            return mod(*a, **k)

        out = call(self.conv1, start, latent[:, 0], noise=noise[0])
        skip = self.to_rgb1(out, latent[:, 1])
        rgb_img_list = [skip]
        i = 1
        for j, to_rgb in enumerate(self.to_rgbs):
            out = call(self.convs[2 * j], out, latent[:, i], noise=noise[2 * j + 1])
            out = call(self.convs[2 * j + 1], out, latent[:, i + 1], noise=noise[2 * j + 2])
            if return_style_scalars and (i + 3) == latent.shape[1]:      # only the last ToRGB reports
                skip = call(to_rgb, out, latent[:, i + 2], skip)
            else:
                skip = to_rgb(out, latent[:, i + 2], skip)
            rgb_img_list.append(skip)
            i += 2
        image = skip

        if PPL_regularize:
            pl_noise = torch.randn_like(image) / math.sqrt(image.shape[2] * image.shape[3])
            grad, = autograd.grad(outputs=(image * pl_noise).sum(), inputs=latent, create_graph=True)
            path_lengths = torch.sqrt(grad.pow(2).sum(2).mean(1))
            return image, path_lengths

        returns = rgb_img_list if return_rgb_list else image
        if return_style_scalars:
            returns = returns, styles_list
        return returns


# --------------------------------------------------------------------------- discriminator
class ConvLayer(nn.Sequential):
    """[Blur ->] EqualConv2d [-> FusedLeakyReLU | ScaledLeakyReLU]."""

    def __init__(self, in_channel, out_channel, kernel_size, downsample=False, blur_kernel=[1, 3, 3, 1],
                 bias=True, activate=True):
        layers = []
        if downsample:
            p = (len(blur_kernel) - 2) + (kernel_size - 1)
            layers.append(Blur(blur_kernel, pad=((p + 1) // 2, p // 2)))
            stride, self.padding = 2, 0
        else:
            stride, self.padding = 1, kernel_size // 2
        layers.append(EqualConv2d(in_channel, out_channel, kernel_size, padding=self.padding, stride=stride,
                                  bias=bias and not activate))
        if activate:
            layers.append(FusedLeakyReLU(out_channel) if bias else ScaledLeakyReLU(0.2))
        super().__init__(*layers)


class ResBlock(nn.Module):
    def __init__(self, in_channel, out_channel, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        self.conv1 = ConvLayer(in_channel, in_channel, 3)
        self.conv2 = ConvLayer(in_channel, out_channel, 3, downsample=True)
        self.skip = ConvLayer(in_channel, out_channel, 1, downsample=True, activate=False, bias=False)

    def forward(self, input):
        return (self.conv2(self.conv1(input)) + self.skip(input)) / _SQRT2


class Discriminator(nn.Module):
    def __init__(self, size, channel_multiplier=2, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        channels = _channel_table(channel_multiplier)
        log_size = int(math.log(size, 2))
        in_channel = channels[size]
        convs = [ConvLayer(3, in_channel, 1)]
        for i in range(log_size, 2, -1):
            out_channel = channels[2 ** (i - 1)]
            convs.append(ResBlock(in_channel, out_channel, blur_kernel))
            in_channel = out_channel
        self.convs = nn.Sequential(*convs)

        self.stddev_group = 4
        self.stddev_feat = 1
        self.final_conv = ConvLayer(in_channel + 1, channels[4], 3)
        self.final_linear = nn.Sequential(
            EqualLinear(channels[4] * 4 * 4, channels[4], activation='fused_lrelu'),
            EqualLinear(channels[4], 1),
        )

    def forward(self, input):
        out = self.convs(input)
        batch, channel, height, width = out.shape
        # minibatch standard deviation over groups of (at most) 4 samples
        group = min(batch, self.stddev_group)
        sd = out.view(group, -1, self.stddev_feat, channel // self.stddev_feat, height, width)
        sd = torch.sqrt(sd.var(0, unbiased=False) + 1e-8)
        sd = sd.mean([2, 3, 4], keepdims=True).squeeze(2).repeat(group, 1, height, width)
        out = self.final_conv(torch.cat([out, sd], 1))
        return self.final_linear(out.view(batch, -1))
