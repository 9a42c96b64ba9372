"""Encoder inference on the tcgen05 implicit-GEMM kernel (eval mode, frozen BatchNorm).

ResNet-18 (E_Tsr / E_W, resnet_encoder.py:258-280 of the reference) and the pSp
GradualStyleEncoder(18, 'ir_se') (E_W_Plus, psp_encoders.py:100-132) run NHWC bf16 end to end:

  * every Conv2d is one ``fm_conv_igemm`` launch with BatchNorm folded into the epilogue table
    (scale/shift), ReLU / PReLU / LeakyReLU as the table's slope, and the residual add fused;
  * a BatchNorm *in front of* a zero-padded 3x3 conv (IR block) is folded into the weights; the
    shift term that zero padding removes at the image border is restored by a 9-class border table;
  * the 3-channel stems pack 8 pixels x 8 channels of a zero-padded row into one 64-wide K chunk
    (overlapping TMA windows) instead of padding 3 channels to 64;
  * the 14 map2style heads are grouped convolutions: heads that share an input are one launch with
    concatenated output channels, later levels are one launch per level over all heads.

Derived tensors are caches keyed on the parameters' version counters, never state.
"""
import math

import os

import weakref

import torch

from . import ops
from .engine import epoch_of, plans_of
from .graphs import GraphRunner


def _versions(module):
    """Cache key of everything derived from ``module``: version counter and address of every parameter / buffer, plus
    the module's engine epoch (``fm3d.engine.invalidate(module)`` for value changes the counters do not show)."""
    return [epoch_of(module)] + [(t._version, t.data_ptr()) for t in module.state_dict(keep_vars=True).values()]


def _fold_bn(bn):
    a = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    return a, bn.bias.detach().float() - bn.running_mean.detach().float() * a


def _table(cout, device, scale=None, shift=None, slope=1.0, groups=1):
    """[groups, cout, 8] fp32 epilogue table (scale, shift, slope, post=1, 0...)."""
    t = torch.zeros(groups, cout, 8, device=device, dtype=torch.float32)
    t[..., 0] = 1.0 if scale is None else scale
    if shift is not None:
        t[..., 1] = shift
    t[..., 2] = slope
    t[..., 3] = 1.0
    return t


def _cs(c):
    return (c + 7) // 8 * 8


def _stem_weight(w, kpix=8):
    """[O, C<=8, kh, kw<=kpix] -> bf16 [kh, O, kpix*8]: K index = kx*8 + c (one padded-row window)."""
    O, C, kh, kw = w.shape
    out = torch.zeros(kh, O, kpix, 8, device=w.device, dtype=torch.float32)
    out[:, :, :kw, :C] = w.detach().float().permute(2, 0, 3, 1)
    return out.reshape(kh, O, kpix * 8).to(torch.bfloat16).contiguous()


class _Conv:
    """One prepared convolution: bf16 weights + epilogue table (+ optional border table)."""

    def __init__(self, weight, stride, pad, tab, border=None, scale=1.0):
        self.k = weight.shape[-1]
        self.cout, self.cin = weight.shape[0], weight.shape[1]
        self.stride, self.pad = stride, pad
        self.wq, _ = ops.prep_weight(weight.detach(), scale, want_wsq=False)
        self.taps = ops.conv_taps(self.k, self.k, pad)
        self.tab = tab
        self.border = border

    def out_size(self, h):
        return (h + 2 * self.pad - self.k) // self.stride + 1

    def run(self, x, out, B, H, W, residual=None, **kw):
        OH, OW = self.out_size(H), self.out_size(W)
        return ops.conv_igemm(x, self.wq, self.taps, out, self.tab, B=B, H=H, W=W, Cin=self.cin, Cout=self.cout,
                              OH=OH, OW=OW, stride=self.stride, residual=residual, border_tab=self.border, **kw)


# ======================================================================================
# ResNet-18
# ======================================================================================
_FUSE_UPSAMPLE = os.environ.get("FM3D_FUSE_UPSAMPLE", "1") != "0"
_FUSE_SE_SUM = os.environ.get("FM3D_FUSE_SE_SUM", "1") != "0"


class ResNetPlan:
    """One ResNet-18, or SEVERAL of identical architecture on the same input batch as the groups of grouped launches
    (E_Tsr and E_W both encode the render under ``tsr_encode='Render Image'``, Util/network_util.py:310-314: at B = 32 a
    ResNet-18 layer is a 20-40 us latency-bound launch, so running the two networks as one sequence with twice the GEMM M
    dimension per launch halves the launches).  Activations are [G*B, h, w, c] group-major, weights [G*taps, Cout, Cin]."""
    models = property(lambda self: [m() for m in self._models])      # weak: the plan cache is keyed weakly on the module

    def __init__(self, models, B, H, W, device):
        self._models, self.B, self.H, self.W, self.device = [weakref.ref(m) for m in models], B, H, W, device
        self.G = len(models)
        self.versions = None
        bf = dict(device=device, dtype=torch.bfloat16)
        GB = self.G * B
        # stem geometry: 7x7 stride 2 pad 3; a window of 8 padded pixels x 8 channels per (ky, out x)
        self.oh, self.ow = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
        self.Hp, self.Wp = H + 6, max(W + 6, 2 * (self.ow - 1) + 8)
        self.packed = torch.zeros(GB, self.Hp, self.Wp, 8, **bf)
        self.stem_out = torch.empty(GB, self.oh, self.ow, 64, **bf)
        ph, pw = (self.oh - 1) // 2 + 1, (self.ow - 1) // 2 + 1
        self.pool_out = torch.empty(GB, ph, pw, 64, **bf)
        self.bufs = {}
        self.runner = GraphRunner(self._run)
        self.refresh()

    def _buf(self, key, n, h, w, c):
        t = self.bufs.get(key)
        if t is None or t.shape != (n, h, w, _cs(c)):
            t = torch.empty(n, h, w, _cs(c), device=self.device, dtype=torch.bfloat16)
            self.bufs[key] = t
        return t

    def refresh(self):
        models = self.models
        v = [_versions(m) for m in models]
        if v == self.versions:
            return
        self.versions = v
        self.runner.invalidate()
        dev, G = self.device, self.G

        def conv(mods, bns, stride, pad, slope):
            """Grouped conv: the same layer of every model (weights [G*taps, O, I], table [G, O, 8])."""
            tabs, ws = [], []
            for cm, bn in zip(mods, bns):
                a, b = _fold_bn(bn)
                tabs.append(_table(cm.out_channels, dev, a, b, slope)[0])
                ws.append(ops.prep_weight(cm.weight.detach(), 1.0, want_wsq=False)[0])
            c = _Conv.__new__(_Conv)
            c.k, c.cout, c.cin = mods[0].weight.shape[-1], mods[0].weight.shape[0], mods[0].weight.shape[1]
            c.stride, c.pad = stride, pad
            c.wq = torch.cat(ws, 0).contiguous()
            c.taps = ops.conv_taps(c.k, c.k, pad)
            c.tab = torch.stack(tabs, 0).contiguous()
            c.border = None
            c.groups = G
            return c
        self.stem_w = torch.cat([_stem_weight(m.conv1.weight) for m in models], 0).contiguous()
        self.stem_tab = torch.stack([_table(64, dev, *_fold_bn(m.bn1), slope=0.0)[0] for m in models], 0).contiguous()
        self.blocks = []
        for li in range(4):
            layers = [getattr(m, f"layer{li + 1}") for m in models]
            for bi in range(len(layers[0])):
                blks = [l[bi] for l in layers]
                c1 = conv([b.conv1 for b in blks], [b.bn1 for b in blks], blks[0].stride, 1, 0.0)
                c2 = conv([b.conv2 for b in blks], [b.bn2 for b in blks], 1, 1, 0.0)
                ds = None
                if blks[0].downsample is not None:
                    ds = conv([b.downsample[0] for b in blks], [b.downsample[1] for b in blks], blks[0].downsample[0].stride[0], 0, 1.0)
                self.blocks.append((c1, c2, ds))

    def run(self, x):
        self.refresh()
        prev, ops.PROFILE_TAG = ops.PROFILE_TAG, "resnet"
        try:
            return self.runner(x.contiguous().float())
        finally:
            ops.PROFILE_TAG = prev

    def _conv(self, c, x, out, n, H, W, residual=None):
        OH, OW = c.out_size(H), c.out_size(W)
        kw = dict(groups=self.G, w_rows=c.cout) if self.G > 1 else {}
        return ops.conv_igemm(x, c.wq, c.taps, out, c.tab, B=n, H=H, W=W, Cin=c.cin, Cout=c.cout, OH=OH, OW=OW,
                              stride=c.stride, residual=residual, **kw)

    def _run(self, x):
        B, G = self.B, self.G
        GB = G * B
        for g in range(G):                       # the same image batch for every group
            ops.image_to_nhwc8_padded(x, 3, 3, self.Hp, self.Wp, out=self.packed[g * B:(g + 1) * B])
        kw = dict(groups=G, w_rows=64) if G > 1 else {}
        # algorithmic FLOPs of the 7x7 conv on 3 channels (the K dimension is padded to 7 x 64 for the tensor core)
        ops.conv_igemm(self.packed, self.stem_w, [(ky, 0, ky) for ky in range(7)], self.stem_out, self.stem_tab,
                       B=GB, H=self.Hp, W=self.ow, Cin=64, Cout=64, OH=self.oh, OW=self.ow, stride_x=1, stride_y=2,
                       x_pixstride=16, x_rowstride=self.Wp * 8, x_imgstride=self.Hp * self.Wp * 8,
                       algo_flops=2.0 * GB * self.oh * self.ow * 3 * 64 * 49, **kw)
        ops.maxpool3x3s2_nhwc(self.stem_out, self.pool_out)
        cur = self.pool_out
        h, w = cur.shape[1], cur.shape[2]
        for i, (c1, c2, ds) in enumerate(self.blocks):
            oh, ow = c1.out_size(h), c1.out_size(w)
            identity = cur
            if ds is not None:
                identity = self._conv(ds, cur, self._buf(("ds", i), GB, oh, ow, ds.cout), GB, h, w)
            y = self._conv(c1, cur, self._buf(("a", i), GB, oh, ow, c1.cout), GB, h, w)
            cur = self._conv(c2, y, self._buf(("b", i), GB, oh, ow, c2.cout), GB, oh, ow, residual=identity)
            h, w = oh, ow
        outs = []
        for g, m in enumerate(self.models):
            part = cur[g * B:(g + 1) * B]
            if m.tensor_encoding:
                outs.append(ops.avgpool_nhwc_to_nchw(part, 512, 2, 2))                     # nn.AvgPool2d(2, 2)
            else:
                outs.append(ops.avgpool_nhwc_to_nchw(part, 512, h, w).flatten(1))         # AdaptiveAvgPool2d((1,1)) + flatten
        return outs


def _resnet_ok(model, x):
    if x.ndim != 4 or x.shape[1] != 3 or x.shape[2] % 32 or x.shape[3] % 32:
        return False
    return [len(l) for l in (model.layer1, model.layer2, model.layer3, model.layer4)] == [2, 2, 2, 2] and \
        type(model.layer1[0]).__name__ == "BasicBlock"


def run_resnet(model, x):
    if not _resnet_ok(model, x):
        return None
    plans = plans_of(model)
    key = (tuple(x.shape), x.device.index, ops.current_slot(), ops.current_cap())      # captured graphs bake the grid sizes in
    plan = plans.get(key)
    if plan is None:
        plan = plans[key] = ResNetPlan([model], x.shape[0], x.shape[2], x.shape[3], x.device)
    return plan.run(x)[0]


def run_resnet_pair(model_a, model_b, x):
    """Both ResNet-18 encoders on the same input as one grouped launch sequence -> (out_a, out_b), or None."""
    if model_a is model_b or not (_resnet_ok(model_a, x) and _resnet_ok(model_b, x)):
        return None
    plans = plans_of(model_a)
    key = ("pair", id(model_b), tuple(x.shape), x.device.index, ops.current_slot(), ops.current_cap())
    plan = plans.get(key)
    if plan is None or plan.models[1] is not model_b:
        plan = plans[key] = ResNetPlan([model_a, model_b], x.shape[0], x.shape[2], x.shape[3], x.device)
    return plan.run(x)


# ======================================================================================
# pSp GradualStyleEncoder (ir_se, 18 layers)
# ======================================================================================
def _border_table(w, b_in):
    """w [O,I,3,3] (un-folded), b_in [I] shift of the input BatchNorm.  Returns (full shift [O],
    border correction [9,O]) for a zero-padded 3x3 stride-1 conv applied to BN(x)."""
    sk = torch.einsum("oikl,i->okl", w.float(), b_in)                 # [O,3,3]
    full = sk.sum((1, 2))
    corr = torch.zeros(9, w.shape[0], device=w.device, dtype=torch.float32)
    for cy in range(3):
        for cx in range(3):
            miss = torch.zeros(3, 3, dtype=torch.bool)
            if cy == 1: miss[0, :] = True
            if cy == 2: miss[2, :] = True
            if cx == 1: miss[:, 0] = True
            if cx == 2: miss[:, 2] = True
            corr[cy * 3 + cx] = -(sk * miss.to(sk.device)).sum((1, 2))
    return full, corr.contiguous()


class _Unit:
    pass


class PspPlan:
    model = property(lambda self: self._model())

    def __init__(self, model, B, device):
        self._model, self.B, self.device = weakref.ref(model), B, device
        self.versions = None
        bf = dict(device=device, dtype=torch.bfloat16)
        S = 256
        self.S = S
        self.Hp, self.Wp = S + 2, S + 8                 # 3x3 pad 1; 8-pixel windows stay in the padded row
        self.packed = torch.zeros(B, self.Hp, self.Wp, 8, **bf)
        self.bufs = {}
        self.sum_buf = torch.zeros(B, 512, device=device, dtype=torch.float32)
        self.gate_buf = torch.empty(B, 512, device=device, dtype=torch.float32)
        self.runner = GraphRunner(self._run)
        self.refresh()

    def _buf(self, key, n, h, w, c):
        t = self.bufs.get(key)
        if t is None:
            t = torch.empty(n, h, w, _cs(c), device=self.device, dtype=torch.bfloat16)
            self.bufs[key] = t
        return t

    def refresh(self):
        v = _versions(self.model)
        if v == self.versions:
            return
        self.versions = v
        self.runner.invalidate()
        m, dev = self.model, self.device
        conv0, bn0, prelu0 = m.input_layer[0], m.input_layer[1], m.input_layer[2]
        a, b = _fold_bn(bn0)
        self.in_w = _stem_weight(conv0.weight)
        self.in_tab = _table(64, dev, a, b, slope=prelu0.weight.detach().float())
        self.units = []
        for unit in m.body:
            u = _Unit()
            res = unit.res_layer
            bn1, conv1, prelu, conv2, bn2, se = res[0], res[1], res[2], res[3], res[4], res[5]
            a1, b1 = _fold_bn(bn1)
            w1 = conv1.weight.detach().float()
            full, corr = _border_table(w1, b1)
            u.c1 = _Conv(w1 * a1.view(1, -1, 1, 1), 1, 1,
                         _table(w1.shape[0], dev, None, full, slope=prelu.weight.detach().float()), border=corr)
            a2, b2 = _fold_bn(bn2)
            u.c2 = _Conv(conv2.weight, conv2.stride[0], 1, _table(conv2.out_channels, dev, a2, b2, 1.0))
            u.stride = conv2.stride[0]
            u.sc = None
            if not isinstance(unit.shortcut_layer, torch.nn.MaxPool2d):
                sconv, sbn = unit.shortcut_layer[0], unit.shortcut_layer[1]
                asc, bsc = _fold_bn(sbn)
                u.sc = _Conv(sconv.weight, sconv.stride[0], 0, _table(sconv.out_channels, dev, asc, bsc, 1.0))
            u.se_w1 = se.fc1.weight.detach().float().reshape(se.fc1.out_channels, -1).contiguous()
            u.se_w2 = se.fc2.weight.detach().float().reshape(se.fc2.out_channels, -1).contiguous()
            u.depth = conv2.out_channels
            self.units.append(u)
        self.lat1 = _Conv(m.latlayer1.weight, 1, 0, _table(512, dev, None, m.latlayer1.bias.detach().float(), 1.0))
        self.lat2 = _Conv(m.latlayer2.weight, 1, 0, _table(512, dev, None, m.latlayer2.bias.detach().float(), 1.0))

        # ---- map2style heads, grouped by level.  depth[j] = number of stride-2 convs of head j
        heads = list(m.styles)
        self.n_heads = len(heads)
        convs = [[mod for mod in h.convs if isinstance(mod, torch.nn.Conv2d)] for h in heads]
        self.head_depth = [len(c) for c in convs]

        def stack(idx_pairs):
            """Grouped weights / tables for [(head, conv index)]: bf16 [G*9, 512, 512], tab [G,512,8]."""
            ws, tabs = [], []
            for (j, k) in idx_pairs:
                c = convs[j][k]
                wq, _ = ops.prep_weight(c.weight.detach(), 1.0, want_wsq=False)
                ws.append(wq)
                tabs.append(_table(512, dev, None, c.bias.detach().float(), 0.01)[0])
            return torch.cat(ws, 0).contiguous(), torch.stack(tabs, 0).contiguous()
        self.stack = stack
        self.coarse = [j for j in range(self.n_heads) if j < m.coarse_ind]
        self.middle = [j for j in range(self.n_heads) if m.coarse_ind <= j < m.middle_ind]
        self.fine = [j for j in range(self.n_heads) if j >= m.middle_ind]
        # first convs share their input: concatenate output channels
        self.first = {}
        for name, hs in (("coarse", self.coarse), ("middle", self.middle), ("fine", self.fine)):
            if hs:
                wq = torch.cat([ops.prep_weight(convs[j][0].weight.detach(), 1.0, want_wsq=False)[0] for j in hs], 1)
                tab = torch.cat([_table(512, dev, None, convs[j][0].bias.detach().float(), 0.01) for j in hs], 1)
                self.first[name] = (wq.contiguous(), tab.contiguous())
        self.fine_l1 = stack([(j, 1) for j in self.fine]) if self.fine else None      # 32 -> 16
        self.fine_l2 = stack([(j, 2) for j in self.fine]) if self.fine else None      # 16 -> 8
        self.mid_l1 = stack([(j, 1) for j in self.middle]) if self.middle else None   # 16 -> 8
        # common tail over all heads: 8->4, 4->2, 2->1  (conv index counted from the end)
        self.tail = [stack([(j, self.head_depth[j] - 3 + t) for j in range(self.n_heads)]) for t in range(3)]
        lw = torch.stack([ops.prep_weight(h.linear.weight.detach().view(512, 512, 1, 1), h.linear.scale,
                                          want_wsq=False)[0][0] for h in heads], 0)          # [G,512,512]
        ltab = torch.stack([_table(512, dev, None, h.linear.bias.detach().float() * h.linear.lr_mul, 1.0)[0]
                            for h in heads], 0)
        self.linear = (lw.contiguous(), ltab.contiguous())

    def run(self, x):
        self.refresh()
        prev, ops.PROFILE_TAG = ops.PROFILE_TAG, "psp"
        try:
            return self.runner(x.contiguous().float())
        finally:
            ops.PROFILE_TAG = prev

    def _run(self, x):
        B, S = self.B, self.S
        m = self.model
        ops.image_to_nhwc8_padded(x, 1, 1, self.Hp, self.Wp, out=self.packed)
        cur = self._buf("in", B, S, S, 64)
        # Cin = 3 pixels x 8 channels of the 8-pixel window carry weights (the rest multiplies zeros): K = 32 per kernel row
        ops.conv_igemm(self.packed, self.in_w, [(ky, 0, ky) for ky in range(3)], cur, self.in_tab,
                       B=B, H=self.Hp, W=S, Cin=24, Cout=64, OH=S, OW=S, stride_x=1, stride_y=1,
                       x_pixstride=8, x_rowstride=self.Wp * 8, x_imgstride=self.Hp * self.Wp * 8,
                       algo_flops=2.0 * B * S * S * 3 * 64 * 9)          # 3x3 conv on 3 channels, not the padded K
        h = S
        feats = {}
        for i, u in enumerate(self.units):
            oh = u.c2.out_size(h)
            r1 = u.c1.run(cur, self._buf(("r1", i), B, h, h, u.c1.cout), B, h, h)
            # the SE squeeze (per-channel mean of r2) rides on this conv's epilogue when a tile lies inside one image
            fused_sum = _FUSE_SE_SUM and oh * oh >= 128 and (oh * oh) % 128 == 0
            r2 = u.c2.run(r1, self._buf(("r2", i), B, oh, oh, u.depth), B, h, h, **(dict(colsum=self.sum_buf, ksplit=1) if fused_sum else {}))
            if u.sc is not None:
                sc, ss = u.sc.run(cur, self._buf(("sc", i), B, oh, oh, u.depth), B, h, h), 1
            else:
                sc, ss = cur, u.stride
            cur = ops.se_block_nhwc(r2, u.depth, u.se_w1, u.se_w2, sc, ss, self.sum_buf, self.gate_buf,
                                    self._buf(("out", i), B, oh, oh, u.depth), summed=fused_sum)
            h = oh
            feats[i] = (cur, h)
        (c1, h1), (c2, h2), (c3, h3) = feats[3], feats[5], feats[7]
        G = self.n_heads
        taps = ops.conv_taps(3, 3, 1)
        lvl8 = self._buf("lvl8", G * B, h3 // 2, h3 // 2, 512)          # all heads at 8x8, head-major

        def first(name, heads, src, hs, dst, dst_head0):
            wq, tab = self.first[name]
            n = len(heads)
            o = hs // 2
            ops.conv_igemm(src, wq, taps, dst[dst_head0 * B:], tab, B=B, H=hs, W=hs, Cin=512, Cout=512 * n, OH=o, OW=o,
                           stride=2, out_cgroup=512, out_gstride=B * o * o * 512, out_cstride=512)

        def grouped(wt, src, n, hs, dst):
            wq, tab = wt
            o = hs // 2
            # 2x2 -> 1x1: the taps in kernel row / column 0 only ever read the zero padding -- 5 of the 9 weight slabs of
            # every head (37 of 66 MB for 14 heads) need not be streamed (slab indices stay those of the full 3x3 layout)
            tp = [t for t in taps if t[0] >= 0 and t[1] >= 0] if hs == 2 else taps
            ops.conv_igemm(src, wq, tp, dst, tab, B=n * B, H=hs, W=hs, Cin=512, Cout=512, OH=o, OW=o, stride=2,
                           groups=n, w_rows=512)

        nc, nm, nf = len(self.coarse), len(self.middle), len(self.fine)
        first("coarse", self.coarse, c3, h3, lvl8, 0)                                    # 16 -> 8
        # FPN top-down path (psp_encoders.py:81-98,123,127): the bilinear upsampling of the coarser map is sampled
        # inside the lateral 1x1 conv's epilogue instead of being written out and read back
        if _FUSE_UPSAMPLE:
            p2 = self.lat1.run(c2, self._buf("p2", B, h2, h2, 512), B, h2, h2, residual_up=c3)
        else:
            up2 = ops.bilinear_up_nhwc(c3, h2, h2, self._buf("up2", B, h2, h2, 512))
            p2 = self.lat1.run(c2, self._buf("p2", B, h2, h2, 512), B, h2, h2, residual=up2)
        mid16 = self._buf("mid16", nm * B, h2 // 2, h2 // 2, 512)
        first("middle", self.middle, p2, h2, mid16, 0)                                   # 32 -> 16
        grouped(self.mid_l1, mid16, nm, h2 // 2, lvl8[nc * B:])                          # 16 -> 8
        if _FUSE_UPSAMPLE:
            p1 = self.lat2.run(c1, self._buf("p1", B, h1, h1, 512), B, h1, h1, residual_up=p2)
        else:
            up1 = ops.bilinear_up_nhwc(p2, h1, h1, self._buf("up1", B, h1, h1, 512))
            p1 = self.lat2.run(c1, self._buf("p1", B, h1, h1, 512), B, h1, h1, residual=up1)
        fine32 = self._buf("fine32", nf * B, h1 // 2, h1 // 2, 512)
        first("fine", self.fine, p1, h1, fine32, 0)                                      # 64 -> 32
        fine16 = self._buf("fine16", nf * B, h1 // 4, h1 // 4, 512)
        grouped(self.fine_l1, fine32, nf, h1 // 2, fine16)                               # 32 -> 16
        grouped(self.fine_l2, fine16, nf, h1 // 4, lvl8[(nc + nm) * B:])                 # 16 -> 8
        cur, hs = lvl8, h3 // 2
        for t in range(3):                                                               # 8 -> 4 -> 2 -> 1
            nxt = self._buf(("tail", t), G * B, hs // 2, hs // 2, 512)
            grouped(self.tail[t], cur, G, hs, nxt)
            cur, hs = nxt, hs // 2
        lw, ltab = self.linear
        codes = torch.empty(G * B, 512, device=self.device, dtype=torch.float32)
        ops.conv_igemm(cur, lw, [(0, 0, 0)], codes, ltab, B=G * B, H=1, W=1, Cin=512, Cout=512, OH=1, OW=1,
                       groups=G, w_rows=512, out_nchw_f32=True)
        return codes.view(G, B, 512).permute(1, 0, 2).contiguous()


def run_psp(model, x):
    m = model
    if x.ndim != 4 or tuple(x.shape[1:]) != (3, 256, 256) or type(m.body[0]).__name__ != "bottleneck_IR_SE":
        return None
    if m.style_count < m.middle_ind + 1 or m.coarse_ind != 3 or m.middle_ind != 7:
        return None
    plans = plans_of(m)
    key = (x.shape[0], x.device.index, ops.current_slot(), ops.current_cap())
    plan = plans.get(key)
    if plan is None:
        plan = plans[key] = PspPlan(m, x.shape[0], x.device)
    return plan.run(x)
