"""Fused bf16 synthesis engine: runs a whole ``stylegan2.Generator`` synthesis network
(stylegan2.py:627-668 of the reference) on the tcgen05 implicit-GEMM kernel.

Dataflow per forward (B samples, activations NHWC bf16, never NCHW fp32 in between):

  style_affine   all 20 modulation EqualLinears in ONE launch        -> s[layer][B,Cin]
  build_tables   demod d[b,o], bias, lrelu slope, gain*s_next, fused ToRGB weights (ONE launch)
  nchw->nhwc     external 4x4 tensor * s_conv1                        -> bf16 NHWC
  conv1          igemm 3x3 (+noise +bias +lrelu*sqrt2, *s_next, RGB)  -> x, rgb_acc
  per resolution:
     up-conv     stride-2 transposed conv as 4 output-parity phases (4/2/2/1 taps: exactly the
                 algorithmic FLOPs, no zero-stuffing work) -> T[(2h+1)^2] bf16.  Its input is stored
                 with one zero row + column per image, so the batch is ONE tall [B*(h+1), h+1] image
                 whose separator rows are the conv's zero padding: the (h+1)^2 phase grids tile
                 without per-image tile tails (65 = 64+1 rows cost 9 tiles of 8 per image otherwise)
     blur_act    4x4 FIR + demod + noise + bias + lrelu, *s_next      -> x
     conv        igemm 3x3 with fused epilogue and fused ToRGB        -> x, rgb_acc
     rgb_finalize  rgb_acc + bias + Upsample(skip)                    -> skip (fp32 NCHW)

Style modulation is applied to the *activations* by the producing layer's epilogue, so all
convolutions use weights shared across the batch (see stylegan2.ModulatedConv2d).  Derived
tensors (bf16 K-major weights, sum-of-squares tables, descriptor arrays) are caches keyed on
the parameters' version counters -- they are never part of the state dict.
"""
import ctypes as C
import math
import weakref

import torch

from . import _lib, ops
from .graphs import GraphRunner
from ._lib import StyleLayer, TableLayer


import os

# Cout <= 128 up-convs: the two column parities of an output-row parity share one launch as a 2*Cout-wide GEMM
# (N = 256 tiles instead of N = 128: the tensor pipe is no longer starved by the per-SM L2 fill rate)
_PAIR_UP = os.environ.get("FM3D_UPPAIR", "1") != "0"
# all output-parity phases of an up-conv in ONE launch (the phase is a tile coordinate of the persistent grid): the
# weight-bound 1- and 2-tap tiles interleave with the tensor-bound 4-tap tiles instead of running in launches of their own
_MERGE_UP = os.environ.get("FM3D_UPMERGE", "1") != "0"
_MERGE_UP_PAIR = os.environ.get("FM3D_UPMERGE", "1") == "2"

# views (dy, dx) of the input a stride-2 transposed 3x3 conv reads, and for each output-row parity py the
# (view, tap of column parity 0, tap of column parity 1 or None) triples -- see _up_phase_taps
_PAIR_VIEWS = {
    0: [((0, 0), 0, 1), ((0, -1), 2, None), ((-1, 0), 6, 7), ((-1, -1), 8, None)],
    1: [((0, 0), 3, 4), ((0, -1), 5, None)],
}


def _up_phase_taps(py, px):
    """Taps of output parity (py,px) of a stride-2 transposed 3x3 conv:
    T[2y+py, 2x+px] = sum_{ky=py (mod 2), kx=px (mod 2)} x[y + (py-ky)/2, x + (px-kx)/2] * W[ky,kx]."""
    taps = []
    for ky in range(3):
        if (ky - py) % 2:
            continue
        for kx in range(3):
            if (kx - px) % 2:
                continue
            taps.append(((py - ky) // 2, (px - kx) // 2, ky * 3 + kx))
    return taps


class _ConvLayer:
    """One modulated 3x3 conv of the plan."""
    __slots__ = ("mod", "act_bias", "noise_w", "cin", "cout", "up", "res_in", "res_out", "latent_idx",
                 "wq", "wsq", "s", "tab", "rgb_mod", "next_idx", "kernel", "name", "wpair", "wpair_all")


class SynthesisPlan:
    def __init__(self, gen, batch, device):
        self._gen = weakref.ref(gen)        # the plan cache is keyed weakly on the generator: no strong reference back
        self.B = batch
        self.device = device
        self.style_dim = gen.style_dim
        lib = _lib.lib()  # fail early if the library is missing
        assert lib is not None

        # ---- walk the generator: conv1, then (up, conv) per resolution; ToRGB after each plain conv
        convs = []

        def add(styled, latent_idx, res_in, up, name):
            L = _ConvLayer()
            L.mod = styled.conv
            L.act_bias = styled.activate.bias
            L.noise_w = styled.noise.weight
            L.cin, L.cout = styled.conv.in_channel, styled.conv.out_channel
            L.up = up
            L.res_in = res_in
            L.res_out = res_in * 2 if up else res_in
            L.latent_idx = latent_idx
            L.rgb_mod = None
            L.kernel = styled.conv.blur.kernel if up else None
            L.name = name
            convs.append(L)
            return L

        first = add(gen.conv1, 0, 4, False, "conv1")
        self.rgbs = [(gen.to_rgb1, 1, first)]
        res, li = 4, 1
        for j, to_rgb in enumerate(gen.to_rgbs):
            add(gen.convs[2 * j], li, res, True, f"convs.{2 * j}")
            res *= 2
            last = add(gen.convs[2 * j + 1], li + 1, res, False, f"convs.{2 * j + 1}")
            self.rgbs.append((to_rgb, li + 2, last))
            li += 2
        self.convs = convs
        for (to_rgb, lidx, L) in self.rgbs:
            L.rgb_mod = (to_rgb, lidx)

        B = batch
        f32 = dict(device=device, dtype=torch.float32)
        bf16 = dict(device=device, dtype=torch.bfloat16)
        cs = lambda c: (c + 7) // 8 * 8

        # ---- per-layer buffers
        for L in convs:
            L.s = torch.empty(B, L.cin, **f32)
            L.tab = torch.empty(B, L.cout, 8, **f32)
        self.rgb_s = [torch.empty(B, L.cout, **f32) for (_, _, L) in self.rgbs]
        # the last layer's activations feed nothing but its fused ToRGB: they are never stored
        # a plain conv's output feeds the next up-conv: [B, h+1, h+1, cs] with a zero last row / column per image
        # (written once here, never by the conv); an up-conv's (blurred) output is dense
        self.acts = [None if i + 1 == len(convs) else
                     (torch.empty(B, L.res_out, L.res_out, cs(L.cout), **bf16) if L.up else
                      torch.zeros(B, L.res_out + 1, L.res_out + 1, cs(L.cout), **bf16))
                     for i, L in enumerate(convs)]
        # interleaved phase outputs, one spare row / column per image (the phases write exact zeros there)
        self.tbuf = {i: torch.empty(B, L.res_out + 2, L.res_out + 2, cs(L.cout), **bf16)
                     for i, L in enumerate(convs) if L.up}
        self.x0 = torch.empty(B, 4, 4, cs(convs[0].cin), **bf16)
        self.rgb_acc = [torch.zeros(B, L.res_out, L.res_out, 4, **f32) for (_, _, L) in self.rgbs]
        self._ptrs = None
        self._versions = None
        self._epoch = None
        self._desc_keepalive = None
        self._runner = GraphRunner(self._run_flat)
        self._refresh_weights()

    # ------------------------------------------------------------------ derived caches
    # Two levels of staleness:
    #   * a parameter tensor was REPLACED (``.to()``, ``load_state_dict(assign=True)``, a new device): the descriptor
    #     arrays hold raw pointers -> rebuild everything and drop the captured graphs;
    #   * a parameter's VALUES changed: its version counter moved (optimizer step, ``copy_`` under ``no_grad``), or the
    #     generator's engine epoch moved -- ``Generator.named_parameters()`` / ``parameters()`` bump it, because whoever
    #     holds parameter handles may write through ``.data`` without touching the version counter (the reference's EMA
    #     does exactly that every iteration: ``accumulate``, train_3_encoder.py:195-200).  The derived tensors are then
    #     re-computed IN PLACE, so captured graphs stay valid.
    # ``Generator.invalidate_engine()`` forces the second kind for writers this cannot see.
    def _param_ptrs(self):
        v = []
        for L in self.convs:
            v += [L.mod.weight.data_ptr(), L.mod.modulation.weight.data_ptr(), L.mod.modulation.bias.data_ptr(),
                  L.act_bias.data_ptr(), L.noise_w.data_ptr()]
        for (to_rgb, _, _) in self.rgbs:
            v += [to_rgb.conv.weight.data_ptr(), to_rgb.bias.data_ptr(), to_rgb.conv.modulation.weight.data_ptr(),
                  to_rgb.conv.modulation.bias.data_ptr()]
        return v

    def _param_versions(self):
        return [L.mod.weight._version for L in self.convs] + [t[0].conv.weight._version for t in self.rgbs]

    def _derive(self, rebuild):
        """(Re-)compute bf16 K-major conv weights, sum-of-squares tables, paired-parity up-conv weights and the fp32
        ToRGB weights from the live parameters; ``rebuild=False`` writes into the existing buffers.  (Modulation
        weights, biases, noise weights, blur kernels and ToRGB biases are read live through their pointers.)"""
        dev = self.device
        for L in self.convs:
            w = L.mod.weight.detach()[0]
            if rebuild:
                L.wq, L.wsq = ops.prep_weight(w, L.mod.scale, want_wsq=True)
                L.wpair = L.wpair_all = None
            else:
                ops.prep_weight(w, L.mod.scale, want_wsq=True, out=(L.wq, L.wsq))
            if L.up and _PAIR_UP and L.cout % 32 == 0 and L.cout <= 128 and L.res_in >= 12:
                # [view][P_x0 rows | P_x1 rows][cin]: zero block where the odd column parity has no tap for the view
                if rebuild:
                    # one tensor behind both row parities (views 0-3 | 4-5), so that one launch can address either
                    n0 = len(_PAIR_VIEWS[0])
                    L.wpair_all = torch.zeros(n0 + len(_PAIR_VIEWS[1]), 2 * L.cout, L.wq.shape[2], device=dev, dtype=torch.bfloat16)
                    L.wpair = {0: L.wpair_all[:n0], 1: L.wpair_all[n0:]}
                for py, views in _PAIR_VIEWS.items():
                    wp = L.wpair[py]
                    for v, (_, t0, t1) in enumerate(views):
                        wp[v, :L.cout].copy_(L.wq[t0, :L.cout])
                        if t1 is not None:
                            wp[v, L.cout:].copy_(L.wq[t1, :L.cout])
        if rebuild:
            self.rgb_w = [to_rgb.conv.weight.detach().reshape(3, -1).contiguous().float() for (to_rgb, _, _) in self.rgbs]
        else:
            for dst, (to_rgb, _, _) in zip(self.rgb_w, self.rgbs):
                dst.copy_(to_rgb.conv.weight.detach().reshape(3, -1))

    def _refresh_weights(self):
        gen = self._gen()
        epoch = gen._engine_epoch if gen is not None else 0
        ptrs = self._param_ptrs()
        if ptrs == self._ptrs:
            vers = self._param_versions()
            if vers != self._versions or epoch != self._epoch:
                self._versions, self._epoch = vers, epoch
                self._derive(rebuild=False)
            return
        self._ptrs, self._versions, self._epoch = ptrs, self._param_versions(), epoch
        self._runner.invalidate()          # derived buffers are re-created: captured graphs are stale
        dev = self.device
        self._derive(rebuild=True)
        # descriptor arrays (device resident)
        n_style = len(self.convs) + len(self.rgbs)
        sl = (StyleLayer * n_style)()
        k = 0
        for L in self.convs:
            m = L.mod.modulation
            sl[k].wmod, sl[k].bmod, sl[k].s = m.weight.data_ptr(), m.bias.data_ptr(), L.s.data_ptr()
            sl[k].cin, sl[k].latent_idx = L.cin, L.latent_idx
            k += 1
        for (to_rgb, lidx, L), s in zip(self.rgbs, self.rgb_s):
            m = to_rgb.conv.modulation
            sl[k].wmod, sl[k].bmod, sl[k].s = m.weight.data_ptr(), m.bias.data_ptr(), s.data_ptr()
            sl[k].cin, sl[k].latent_idx = L.cout, lidx
            k += 1
        self.n_style = n_style
        self.max_cin = max(max(L.cin for L in self.convs), max(L.cout for (_, _, L) in self.rgbs))
        tl = (TableLayer * len(self.convs))()
        rgb_of = {id(L): (w, s) for (_, _, L), w, s in zip(self.rgbs, self.rgb_w, self.rgb_s)}
        for i, L in enumerate(self.convs):
            nxt = self.convs[i + 1] if i + 1 < len(self.convs) else None
            tl[i].s, tl[i].wsq = L.s.data_ptr(), L.wsq.data_ptr()
            tl[i].act_bias = L.act_bias.data_ptr()
            tl[i].s_next = nxt.s.data_ptr() if nxt is not None else None
            if id(L) in rgb_of:
                tl[i].wrgb, tl[i].s_rgb = rgb_of[id(L)][0].data_ptr(), rgb_of[id(L)][1].data_ptr()
            tl[i].tab = L.tab.data_ptr()
            tl[i].cin, tl[i].cout = L.cin, L.cout
            tl[i].slope, tl[i].gain = 0.2, math.sqrt(2.0)
        self.max_cout = max(L.cout for L in self.convs)

        def to_dev(arr):
            raw = bytes(arr)
            t = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
            return t
        self.style_desc = to_dev(sl)
        self.table_desc = to_dev(tl)
        self._desc_keepalive = (sl, tl)

    # ------------------------------------------------------------------ forward
    def run(self, latent, start, noise):
        """latent [B,n_latent,D] fp32, start [B,C0,4,4] fp32 NCHW, noise: list (None = fresh
        N(0,1) drawn in the reference's order) -> list of fp32 NCHW RGB images per resolution."""
        self._refresh_weights()
        noise = [None if n is None else n.contiguous().float() for n in noise]
        prev, ops.PROFILE_TAG = ops.PROFILE_TAG, "generator"
        try:
            return self._runner(latent.contiguous().float(), start.contiguous().float(), *noise)
        finally:
            ops.PROFILE_TAG = prev

    def _up_conv(self, L, x, t, B, h):
        """Stride-2 transposed 3x3 conv (no epilogue) of the tall image x [B*(h+1), h+1] (zero separator row / column
        per image) into the interleaved tensor t [B*(2h+2), 2h+2]: phase (py,px) of input pixel (y,x) lands on
        (2y+py, 2x+px); the separator inputs produce the last row / column of T and exact zeros in the spare ones."""
        Ht, Wt = B * (h + 1), h + 1
        up_flops = 2.0 * B * h * h * L.cin * L.cout           # per tap, at the input resolution
        # generic-mode tile (1-tap phase, narrow images): the width that wastes the fewest columns
        tw_ = min((8, 16, 32, 64, 128), key=lambda w: ((Wt + w - 1) // w * w, -w))
        th_ = 128 // tw_
        kw = dict(B=1, H=Ht, W=Wt, Cin=L.cin, OH=Ht, OW=Wt, out_H=B * (2 * h + 2), out_W=2 * h + 2, out_ys=2, out_xs=2,
                  tab_per_sample=False, tile_w=tw_, tile_h=th_)
        # halo-patch eligibility of the tall image (igemm.cu): only then can the phase be a tile coordinate
        # (measured, B=32: 8->16 98 -> 85 us, 16->32 120 -> 78, 32->64 229 -> 199, 64->128 346 -> 344; the paired-parity
        # 128->256 layer is better off with its two launches: 405 vs 414 us, FM3D_UPMERGE=2 merges it too)
        merged = _MERGE_UP and Wt >= 8 and Ht >= 12 and (L.wpair is None or _MERGE_UP_PAIR)
        if merged and L.wpair is not None:
            cs_t = t.shape[-1]
            taps, phases, v0 = [], [], 0
            for py, views in _PAIR_VIEWS.items():
                taps += [(dy, dx, v0 + v) for v, ((dy, dx), _, _) in enumerate(views)]
                phases.append((len(views), py, 0))
                v0 += len(views)
            ops.conv_igemm(x, L.wpair_all, taps, t, None, Cout=2 * L.cout, out_cgroup=L.cout, out_gstride=cs_t,
                           out_cstride=cs_t, algo_flops=up_flops * 9, phases=phases, **kw)
        elif merged:
            taps, phases = [], []
            for py in (0, 1):
                for px in (0, 1):
                    tp = _up_phase_taps(py, px)
                    taps += tp
                    phases.append((len(tp), py, px))
            ops.conv_igemm(x, L.wq, taps, t, None, Cout=L.cout, algo_flops=up_flops * 9, phases=phases, **kw)
        elif L.wpair is not None:
            # one launch per output-row parity: N = [even columns | odd columns] (2*Cout wide)
            cs_t = t.shape[-1]
            for py, views in _PAIR_VIEWS.items():
                taps = [(dy, dx, v) for v, ((dy, dx), _, _) in enumerate(views)]
                ntap = sum(1 + (t1 is not None) for (_, _, t1) in views)
                ops.conv_igemm(x, L.wpair[py], taps, t, None, Cout=2 * L.cout, out_y0=py, out_x0=0,
                               out_cgroup=L.cout, out_gstride=cs_t, out_cstride=cs_t, algo_flops=up_flops * ntap, **kw)
        else:
            for py in (0, 1):
                for px in (0, 1):
                    taps = _up_phase_taps(py, px)
                    ops.conv_igemm(x, L.wq, taps, t, None, Cout=L.cout, out_y0=py, out_x0=px,
                                   algo_flops=up_flops * len(taps), **kw)

    def _run_flat(self, latent, start, *noise):
        lib = _lib.lib()
        B = self.B
        dev = self.device
        latent = latent.contiguous()
        st = torch.cuda.current_stream().cuda_stream
        with torch.cuda.device(dev):
            _lib.check(lib.fm_style_affine(self.style_desc.data_ptr(), self.n_style, self.max_cin, latent.data_ptr(), B,
                                           latent.shape[1], self.style_dim, st), "fm_style_affine")
            _lib.check(lib.fm_build_tables(self.table_desc.data_ptr(), len(self.convs), self.max_cout, self.max_cin, B, st),
                       "fm_build_tables")
            first = self.convs[0]
            _lib.check(lib.fm_nchw_to_nhwc_bf16(self.x0.data_ptr(), start.contiguous().float().data_ptr(),
                                                first.s.data_ptr(), B, first.cin, 4, 4, self.x0.shape[-1], st),
                       "fm_nchw_to_nhwc_bf16")
        x = self.x0
        outs = []
        skip = None
        rgb_i = 0
        for i, L in enumerate(self.convs):
            nz = noise[i]
            if nz is None:
                nz = torch.empty(B, 1, L.res_out, L.res_out, device=dev, dtype=torch.float32).normal_()
            nz = nz.contiguous().float()
            per_sample = nz.shape[0] != 1
            if nz.shape[0] not in (1, B):
                raise RuntimeError(f"noise batch {nz.shape[0]} does not match batch {B}")
            y = self.acts[i]
            h = L.res_in
            if L.up:
                t = self.tbuf[i]
                # (Tried: producing and blurring the intermediate a few samples at a time so it stays in L2 -- the extra
                # launches and tile tails cost more than the DRAM round trip saves: 3.8k -> 3.1-3.6k img/s.)
                self._up_conv(L, x, t, B, h)
                ops.blur_act_nhwc(t, L.kernel, L.tab, nz, per_sample, L.noise_w, L.cout, out=y, padded=True)
            else:
                rgb = self.rgb_acc[rgb_i] if L.rgb_mod is not None else None
                ops.conv_igemm(x, L.wq, ops.conv_taps(3, 3, 1), y, L.tab, B=B, H=h, W=h, Cin=L.cin, Cout=L.cout,
                               OH=h, OW=h, tab_per_sample=True, noise=nz, noise_per_sample=per_sample,
                               noise_w=L.noise_w, rgb=rgb, out_pitch_h=h + 1 if y is not None else 0,
                               out_pitch_w=h + 1 if y is not None else 0)
                if L.rgb_mod is not None:
                    to_rgb = L.rgb_mod[0]
                    kern = to_rgb.upsample.kernel if skip is not None else None
                    skip = ops.rgb_finalize(rgb, to_rgb.bias.detach().reshape(3), skip, kern)     # live view of the parameter
                    outs.append(skip)
                    rgb_i += 1
            x = y
        return outs


# plans live OUTSIDE the module, keyed weakly on it: ``nn.DataParallel.replicate`` shallow-copies ``__dict__`` (replicas
# would share -- and keep alive -- the first replica's plans), and a plan holds ctypes arrays and CUDA graphs that
# ``copy.deepcopy`` / ``torch.save`` of the generator must not meet
_PLANS = weakref.WeakKeyDictionary()


_EPOCHS = weakref.WeakKeyDictionary()


def plans_of(module):
    d = _PLANS.get(module)
    if d is None:
        d = _PLANS[module] = {}
    return d


def epoch_of(module):
    e = getattr(module, "_engine_epoch", None)
    return _EPOCHS.get(module, 0) if e is None else e


def invalidate(module):
    """Mark the engine's derived tensors of ``module`` (a Generator or an encoder) stale: parameter or buffer VALUES
    were changed in a way version counters do not show (``.data`` writes).  ``Generator.invalidate_engine()`` is
    the same thing as a method."""
    if hasattr(module, "_engine_epoch"):
        module._engine_epoch += 1
    else:
        _EPOCHS[module] = _EPOCHS.get(module, 0) + 1


def check_inputs(gen, latent, start, noise):
    """Shape contract of the engine (the reference raises a shape error inside the first mismatching op; the engine
    passes raw pointers to kernels, so it has to check up front).  Returns False for valid inputs the engine does not
    cover (the caller then takes the differentiable composition), raises on invalid ones."""
    B = latent.shape[0]
    if latent.ndim != 3 or latent.shape[1] < gen.n_latent or latent.shape[2] != gen.style_dim:
        raise RuntimeError(f"Generator: latent must be [B, >={gen.n_latent}, {gen.style_dim}], got {tuple(latent.shape)}")
    c0 = gen.conv1.conv.in_channel
    if start.ndim != 4 or start.shape[0] != B or start.shape[1] != c0:
        raise RuntimeError(f"Generator: start tensor must be [{B}, {c0}, 4, 4], got {tuple(start.shape)}")
    if tuple(start.shape[2:]) != (4, 4):
        return False                     # other start resolutions: valid for the module composition, not planned here
    if len(noise) != gen.num_layers:
        raise RuntimeError(f"Generator: {gen.num_layers} noise maps expected, got {len(noise)}")
    for i, n in enumerate(noise):
        if n is None:
            continue
        res = 2 ** ((i + 5) // 2)
        if n.ndim != 4 or n.shape[0] not in (1, B) or tuple(n.shape[1:]) != (1, res, res):
            raise RuntimeError(f"Generator: noise[{i}] must be [1|{B}, 1, {res}, {res}], got {tuple(n.shape)}")
    return True


def run_synthesis(gen, latent, start, noise):
    """Entry used by ``stylegan2.Generator.forward``: cached plan per (batch, device, engine slot)."""
    plans = plans_of(gen)
    key = (latent.shape[0], latent.device.index, ops.current_slot(), ops.current_cap())
    plan = plans.get(key)
    if plan is None:
        plan = plans[key] = SynthesisPlan(gen, latent.shape[0], latent.device)
    return plan.run(latent, start, noise)
