"""pSp GradualStyleEncoder (E_W_Plus): photo -> [B, n_styles, 512] W+ codes.
Mirrors the reference's psp_encoder_model/encoders/psp_encoders.py (:20-132): IR-SE body,
three pyramid levels (coarse/middle/fine) and one strided-conv "map2style" head per style."""
import os

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn
from torch.nn import BatchNorm2d, Conv2d, Module, PReLU, Sequential

from .helpers import bottleneck_IR, bottleneck_IR_SE, get_blocks
from stylegan2 import EqualLinear


class GradualStyleBlock(Module):
    """log2(spatial) stride-2 3x3 convs (each + LeakyReLU(0.01)) down to 1x1, then EqualLinear."""

    def __init__(self, in_c, out_c, spatial):
        super().__init__()
        self.out_c = out_c
        self.spatial = spatial
        steps = int(np.log2(spatial))
        mods = []
        for i in range(steps):
            mods += [Conv2d(in_c if i == 0 else out_c, out_c, kernel_size=3, stride=2, padding=1), nn.LeakyReLU()]
        self.convs = nn.Sequential(*mods)
        self.linear = EqualLinear(out_c, out_c, lr_mul=1)

    def forward(self, x):
        return self.linear(self.convs(x).view(-1, self.out_c))


_INTERP = {}


def _interp_matrix(n_in, n_out, device):
    """[n_out, n_in] weights of 1-D linear interpolation with align_corners=True (F.interpolate semantics)."""
    key = (n_in, n_out, device)
    m = _INTERP.get(key)
    if m is None:
        src = torch.arange(n_out, dtype=torch.float64) * ((n_in - 1) / (n_out - 1) if n_out > 1 else 0.0)
        i0 = src.floor().clamp(max=n_in - 1).long()
        i1 = (i0 + 1).clamp(max=n_in - 1)
        f = (src - i0.double()).float()
        m = torch.zeros(n_out, n_in)
        m[torch.arange(n_out), i0] += 1.0 - f
        m[torch.arange(n_out), i1] += f
        m = _INTERP[key] = m.to(device)
    return m


class GradualStyleEncoder(Module):
    def __init__(self, num_layers, mode='ir', opts=None):
        super().__init__()
        assert num_layers in [18, 50, 100, 152], 'num_layers should be 18, 50, 100, or 152'
        assert mode in ['ir', 'ir_se'], 'mode should be ir or ir_se'
        self.num_layers = num_layers
        unit = bottleneck_IR if mode == 'ir' else bottleneck_IR_SE
        self.input_layer = Sequential(Conv2d(opts.input_nc, 64, (3, 3), 1, 1, bias=False), BatchNorm2d(64), PReLU(64))
        self.body = Sequential(*[unit(b.in_channel, b.depth, b.stride) for blk in get_blocks(num_layers) for b in blk])
        self.styles = nn.ModuleList()
        self.style_count = opts.n_styles
        self.coarse_ind = 3
        self.middle_ind = 7
        for i in range(self.style_count):
            spatial = 16 if i < self.coarse_ind else (32 if i < self.middle_ind else 64)
            self.styles.append(GradualStyleBlock(512, 512, spatial))
        self.latlayer1 = nn.Conv2d(256, 512, kernel_size=1, stride=1, padding=0)
        self.latlayer2 = nn.Conv2d(128, 512, kernel_size=1, stride=1, padding=0)
        if os.environ.get("FM3D_NATIVE_ENC", "1") != "0":
            # under autograd (training) the convolutions differentiate through the tcgen05 kernels, like G's and D's
            from fm3d.convgrad import use_native_convs
            use_native_convs(self)

    def _upsample_add(self, x, y):
        """Bilinear (align_corners) upsample of the coarser map to y's size, plus y (reference psp_encoders.py:81-98).
        Bilinear interpolation is separable and linear: out = A_h x A_w^T with two small interpolation matrices, i.e. two
        batched GEMMs that autograd differentiates for free.  (ATen's NCHW bilinear kernel took 8 ms per call at
        [32,512,32,32] -> 64x64, 4 % of a training iteration.)"""
        if not x.is_cuda:
            return F.interpolate(x, size=y.shape[2:], mode='bilinear', align_corners=True) + y
        ah = _interp_matrix(x.shape[2], y.shape[2], x.device)
        aw = _interp_matrix(x.shape[3], y.shape[3], x.device)
        return torch.einsum('oh,bchw,pw->bcop', ah, x, aw) + y

    def _taps(self):
        return {50: (6, 20, 23), 18: (3, 5, 7)}[self.num_layers]

    def _engine_ok(self, x):
        return (x.is_cuda and not self.training and self.num_layers == 18 and
                os.environ.get("FM3D_ENGINE", "1") != "0" and
                not (torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))))

    def forward(self, x):
        if self._engine_ok(x):
            from fm3d.encoder_engine import run_psp
            out = run_psp(self, x)
            if out is not None:
                return out
        x = self.input_layer(x)
        t1, t2, t3 = self._taps()
        feats = {}
        for i, unit in enumerate(self.body):
            x = unit(x)
            feats[i] = x
        c1, c2, c3 = feats[t1], feats[t2], feats[t3]
        codes = [self.styles[j](c3) for j in range(self.coarse_ind)]
        p2 = self._upsample_add(c3, self.latlayer1(c2))
        codes += [self.styles[j](p2) for j in range(self.coarse_ind, self.middle_ind)]
        p1 = self._upsample_add(p2, self.latlayer2(c1))
        codes += [self.styles[j](p1) for j in range(self.middle_ind, self.style_count)]
        return torch.stack(codes, dim=1)
