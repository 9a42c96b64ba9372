"""Tensor-level wrappers over the C ABI: unpack torch tensors (device pointer, sizes,
current stream), allocate outputs, call libfm3d.  Mirrors what the reference's pybind
shims do (op/fused_bias_act.cpp:11-17, op/upfirdn2d.cpp:12-19): CUDA checks, contiguity,
``torch.empty`` outputs, current-stream semantics.  No CPU path.
"""
import contextlib
import ctypes as C
import threading
import weakref

import torch

from . import _lib
from ._lib import ConvDesc, FM_BF16, FM_F16, FM_F32

# ---- engine slots: plans own their activation buffers and CUDA graphs, so two forwards of the same model that
# are in flight at once (different streams) must use different plans.  ``with engine_slot(k):`` selects plan set k
# for the calling thread (default 0).
class _Slot(threading.local):
    value = 0


_SLOT = _Slot()


@contextlib.contextmanager
def engine_slot(k):
    prev = _SLOT.value
    _SLOT.value = int(k)
    try:
        yield
    finally:
        _SLOT.value = prev


def current_slot():
    return _SLOT.value


# ---- SM partitions: ``with cta_cap(n):`` confines the persistent grids of the implicit-GEMM launches issued inside to
# n CTAs (one CTA owns an SM).  The two ResNet-18 encoders are ~40 launches of 10-40 us that are latency-bound on 148
# SMs; on a few SMs they take about as long and the rest of the chip keeps running the large layers of the other
# networks (fm_conv_desc.max_ctas).
class _Cap(threading.local):
    value = 0


_CAP = _Cap()


def current_cap():
    return _CAP.value


def sm_partition(device):
    """(CTAs per ResNet-18, CTAs of the W+ encoder's launches, CTAs of the generator's launches; 0 = every SM) for the
    3-encoder forward when its three encoders run concurrently.  Measured on B200 at batch 32 (tools/exp_partition.sh,
    profiles/r02_sm_partition.txt): 16 / 116 / all is +5.5 % images/s over every launch taking all 148 SMs.
    FM3D_PARTITION=0 turns it off, FM3D_PARTITION=r,p,g sets the three numbers."""
    env = __import__("os").environ.get("FM3D_PARTITION", "")
    if env == "0":
        return 0, 0, 0
    if env:
        r, p, g = (int(v) for v in env.split(","))
        return r, p, g
    sms = torch.cuda.get_device_properties(device).multi_processor_count
    r = max(2, (sms * 16 // 148) // 2 * 2)
    return r, sms - 2 * r, 0


@contextlib.contextmanager
def cta_cap(n):
    prev = _CAP.value
    _CAP.value = int(n or 0)
    try:
        yield
    finally:
        _CAP.value = prev


# bench.py sets this to a list to collect (start_event, end_event, algorithmic_flops, tag, max_ctas) per conv launch;
# PROFILE_TAG names the network the launches belong to (set by the engines: "generator", "resnet", "psp")
PROFILE = None
PROFILE_TAG = ""

_DTYPES = {torch.float32: FM_F32, torch.float16: FM_F16, torch.bfloat16: FM_BF16}


def _dtype_code(t):
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise RuntimeError(f"libfm3d: unsupported dtype {t.dtype} (float32, float16, bfloat16 only)")


def _check_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (libfm3d has no CPU path)")


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


# ------------------------------------------------------------------ bias + activation
def bias_act(x, bias=None, ref=None, act=3, grad=0, alpha=0.2, scale=2 ** 0.5):
    """fused_bias_act(input, bias, refer, act, grad, alpha, scale) of the reference
    (op/fused_bias_act.cpp:11-17); ``None`` / empty tensors mean "absent" (.cu:62-63)."""
    _check_cuda(x, "input")
    x = x.contiguous()
    if bias is not None and bias.numel() == 0:
        bias = None
    if ref is not None and ref.numel() == 0:
        ref = None
    if bias is not None:
        _check_cuda(bias, "bias")
        bias = bias.contiguous().to(x.dtype)
    if ref is not None:
        ref = ref.contiguous().to(x.dtype)
        if ref.numel() != x.numel():
            raise RuntimeError("refer must have as many elements as input")
    out = torch.empty_like(x)
    if x.numel() == 0:
        return out
    if x.ndim >= 2:
        n_outer, channels = x.shape[0], x.shape[1]
        inner = x.numel() // (n_outer * channels)
    else:                       # 1-D: the reference's step_b = 1, bias indexed by xi % size_b
        n_outer, channels, inner = 1, x.shape[0], 1
    if bias is not None and bias.numel() != channels:
        raise RuntimeError(f"bias has {bias.numel()} elements, expected {channels}")
    with torch.cuda.device(x.device):
        st = _lib.lib().fm_bias_act(_ptr(out), _ptr(x), _ptr(bias), _ptr(ref), n_outer, channels, inner,
                                    int(act), int(grad), float(alpha), float(scale), _dtype_code(x), _stream())
    _lib.check(st, "fm_bias_act")
    return out


def bias_act_grad_bias(grad_out, ref, act=3, alpha=0.2, scale=2 ** 0.5):
    """Gradient w.r.t. the input plus the per-channel bias gradient in one pass."""
    _check_cuda(grad_out, "grad_output")
    g = grad_out.contiguous()
    ref = ref.contiguous().to(g.dtype)
    gin = torch.empty_like(g)
    if g.ndim >= 2:
        n_outer, channels = g.shape[0], g.shape[1]
        inner = g.numel() // max(n_outer * channels, 1)
    else:
        n_outer, channels, inner = 1, g.shape[0], 1
    gb = torch.zeros(channels, device=g.device, dtype=torch.float32)
    if g.numel():
        with torch.cuda.device(g.device):
            st = _lib.lib().fm_bias_act_grad_bias(_ptr(gin), _ptr(gb), _ptr(g), _ptr(ref), n_outer, channels, inner,
                                                  int(act), float(alpha), float(scale), _dtype_code(g), _stream())
        _lib.check(st, "fm_bias_act_grad_bias")
    return gin, gb.to(g.dtype)


def channel_scale(x, s):
    """x [B, C, ...] * s [B, C] broadcast over the trailing dims (fm_channel_scale)."""
    _check_cuda(x, "input")
    _check_cuda(s, "scale")
    x = x.contiguous()
    rows = x.shape[0] * x.shape[1]
    s = s.contiguous().to(x.dtype)
    if s.numel() != rows:
        raise RuntimeError(f"channel_scale: scale has {s.numel()} elements for {rows} (sample, channel) rows")
    out = torch.empty_like(x)
    if x.numel():
        with torch.cuda.device(x.device):
            st = _lib.lib().fm_channel_scale(_ptr(out), _ptr(x), _ptr(s), rows, x.numel() // rows, _dtype_code(x), _stream())
        _lib.check(st, "fm_channel_scale")
    return out


def plane_add(x, n):
    """x [B, C, H, W] + n [B or 1, 1, H, W] broadcast over the channels (fm_plane_add)."""
    _check_cuda(x, "input")
    _check_cuda(n, "plane")
    x = x.contiguous()
    n = n.contiguous().to(x.dtype)
    B, Cc = x.shape[0], x.shape[1]
    inner = x.numel() // max(B * Cc, 1)
    if n.numel() not in (inner, B * inner):
        raise RuntimeError(f"plane_add: plane of {tuple(n.shape)} does not broadcast over the channels of {tuple(x.shape)}")
    out = torch.empty_like(x)
    if x.numel():
        with torch.cuda.device(x.device):
            st = _lib.lib().fm_plane_add(_ptr(out), _ptr(x), _ptr(n), B, Cc, inner, inner if n.numel() == B * inner and B > 1 else 0,
                                         _dtype_code(x), _stream())
        _lib.check(st, "fm_plane_add")
    return out


def channel_dot(a, b):
    """sum over the trailing dims of a * b, [B, C, ...] x [B, C, ...] -> [B, C] (fm_channel_dot)."""
    _check_cuda(a, "a")
    _check_cuda(b, "b")
    if a.shape != b.shape:
        raise RuntimeError(f"channel_dot: shapes differ: {tuple(a.shape)} vs {tuple(b.shape)}")
    a = a.contiguous()
    b = b.contiguous().to(a.dtype)
    rows = a.shape[0] * a.shape[1]
    out = torch.zeros(a.shape[0], a.shape[1], device=a.device, dtype=torch.float32)
    if a.numel():
        with torch.cuda.device(a.device):
            st = _lib.lib().fm_channel_dot(_ptr(out), _ptr(a), _ptr(b), rows, a.numel() // rows, _dtype_code(a), _stream())
        _lib.check(st, "fm_channel_dot")
    return out.to(a.dtype)


# ------------------------------------------------------------------ upfirdn2d
def upfirdn2d_planes(x, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1):
    """x [..., H, W] -> [..., out_h, out_w] (the reference's [major,H,W,1] layout,
    op/upfirdn2d.py:108)."""
    _check_cuda(x, "input")
    _check_cuda(kernel, "kernel")
    x = x.contiguous()
    k = kernel.contiguous().to(torch.float32)
    in_h, in_w = x.shape[-2], x.shape[-1]
    kh, kw = k.shape
    out_h = (in_h * up_y + pad_y0 + pad_y1 - kh + down_y) // down_y     # op/upfirdn2d_kernel.cu:237
    out_w = (in_w * up_x + pad_x0 + pad_x1 - kw + down_x) // down_x
    out_h, out_w = max(out_h, 0), max(out_w, 0)
    planes = x.numel() // max(in_h * in_w, 1) if in_h * in_w else 0
    out = torch.empty(*x.shape[:-2], out_h, out_w, device=x.device, dtype=x.dtype)
    if out.numel() == 0:
        return out
    with torch.cuda.device(x.device):
        st = _lib.lib().fm_upfirdn2d(_ptr(out), _ptr(x), _ptr(k), planes, in_h, in_w, kh, kw,
                                     up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1,
                                     _dtype_code(x), _stream())
    _lib.check(st, "fm_upfirdn2d")
    return out


# ------------------------------------------------------------------ implicit-GEMM conv
_SPLITK_WS = {}
SPLITK_WS_BYTES = 64 << 20


class _Scope(threading.local):
    owner = None


_SCOPE = _Scope()


@contextlib.contextmanager
def splitk_scope(owner):
    """Convs issued inside the scope reduce their split-K partial sums in a workspace owned by ``owner`` (an engine
    plan's GraphRunner).  The stream alone is not a safe key: every plan is captured on torch's one shared
    graph-capture stream and the captured graphs are then replayed concurrently on different streams."""
    prev = _SCOPE.owner
    _SCOPE.owner = owner
    try:
        yield
    finally:
        _SCOPE.owner = prev


def _splitk_workspace(device):
    """fp32 zero workspace: convs that share one run in stream order and each user re-zeroes what it touched.
    One per engine plan (see splitk_scope); outside a plan, one per (device, stream)."""
    owner = _SCOPE.owner
    if owner is not None:
        store = owner.__dict__.setdefault("_splitk_ws", {})
        key = device
    else:
        store = _SPLITK_WS
        key = (device, torch.cuda.current_stream(device).cuda_stream)
    ws = store.get(key)
    if ws is None:
        ws = store[key] = torch.zeros(SPLITK_WS_BYTES // 4, device=device, dtype=torch.float32)
    return ws


_DYN_TILES = __import__("os").environ.get("FM3D_DYNTILES", "0") != "0"
_DYN_TILES_ALL = __import__("os").environ.get("FM3D_DYNTILES", "0") == "2"     # tools: also eager launches, one pair per stream
_EAGER_CTR = {}


def _tile_counter(device):
    """Counter pair of the dynamic tile schedule (fm_conv_desc.tile_counter), one per engine plan: a plan's launches
    are stream-ordered, and plans that run concurrently (other streams, other batches in flight) are other owners.
    Outside a plan (eager calls from autograd, tools) launches keep the static schedule."""
    owner = _SCOPE.owner
    if not _DYN_TILES:
        return None
    if owner is None:
        if not _DYN_TILES_ALL:
            return None
        store, device = _EAGER_CTR, (device, torch.cuda.current_stream(device).cuda_stream)
    else:
        store = owner.__dict__.setdefault("_tile_ctr", {})
    ctr = store.get(device)
    if ctr is None:
        ctr = store[device] = torch.zeros(2, device=device[0] if isinstance(device, tuple) else device, dtype=torch.int32)
    return ctr


def conv_taps(kh, kw, pad):
    """Tap list (dy, dx, weight slab) of a plain correlation (F.conv2d semantics)."""
    return [(ky - pad, kx - pad, ky * kw + kx) for ky in range(kh) for kx in range(kw)]


def conv_igemm(x, w, taps, out, tab, *, B, H, W, Cin, Cout, OH, OW, stride=1, w_rows=None,
               out_H=None, out_W=None, out_y0=0, out_x0=0, out_ys=1, out_xs=1, out_nchw_f32=False,
               tab_per_sample=False, noise=None, noise_per_sample=True, noise_w=None, residual=None,
               rgb=None, block_n=0, tile_w=0, tile_h=0, stride_x=0, stride_y=0, x_pixstride=0, x_rowstride=0,
               x_imgstride=0, groups=1, border_tab=None, out_cgroup=0, out_gstride=0, out_cstride=None, ksplit=0, upmode=False,
               out_cgroup_ow_shrink=0, algo_flops=None, residual_up=None, out_pitch_h=0, out_pitch_w=0, phases=None,
               max_ctas=None, tile_counter=None, colsum=None):
    """Launch fm_conv_igemm.  x: bf16 NHWC [B,H,W,cs]; w: bf16 [slabs, w_rows, cin_stride];
    out: bf16 NHWC [B,out_H,out_W,cs_out] or fp32 NCHW; tab: fp32 [B|1, Cout, 8]."""
    d = ConvDesc()
    d.x = x.data_ptr(); d.B, d.H, d.W, d.Cin = B, H, W, Cin
    d.x_cstride = x.shape[-1]
    d.w = w.data_ptr(); d.ntaps = len(taps); d.Cout = Cout
    d.w_rows = w.shape[1] if w_rows is None else w_rows
    d.w_cstride = w.shape[2]
    for i, (dy, dx, wi) in enumerate(taps):
        d.tap_dy[i] = dy; d.tap_dx[i] = dx; d.tap_widx[i] = wi
    d.stride = stride
    d.stride_x, d.stride_y = stride_x, stride_y
    d.x_pixstride, d.x_rowstride, d.x_imgstride = x_pixstride, x_rowstride, x_imgstride
    d.groups = groups
    d.border_tab = _ptr(border_tab)
    d.out_cgroup, d.out_gstride = out_cgroup, out_gstride
    d.out_cgroup_ow_shrink = out_cgroup_ow_shrink
    d.OH, d.OW = OH, OW
    d.out = out.data_ptr() if out is not None else None
    d.out_H = OH if out_H is None else out_H
    d.out_W = OW if out_W is None else out_W
    d.out_cstride = out_cstride if out_cstride is not None else (Cout if out_nchw_f32 else (out.shape[-1] if out is not None else (Cout + 7) // 8 * 8))
    d.out_y0, d.out_x0, d.out_ys, d.out_xs = out_y0, out_x0, out_ys, out_xs
    d.out_nchw_f32 = 1 if out_nchw_f32 else 0
    d.out_pitch_h, d.out_pitch_w = out_pitch_h, out_pitch_w
    d.tab = tab.data_ptr() if tab is not None else None; d.tab_bstride = 1 if tab_per_sample else 0
    d.noise = _ptr(noise); d.noise_bstride = 1 if noise_per_sample else 0
    d.noise_w = _ptr(noise_w)
    d.residual = _ptr(residual)
    if residual_up is not None:          # low-resolution [B,h,w,cs] tensor, bilinearly upsampled in the epilogue
        d.residual = _ptr(residual_up)
        d.residual_up_h, d.residual_up_w = residual_up.shape[1], residual_up.shape[2]
    d.rgb = _ptr(rgb)
    d.block_n, d.tile_w, d.tile_h = block_n, tile_w, tile_h
    ws = _splitk_workspace(x.device)
    d.splitk_ws, d.splitk_ws_bytes, d.ksplit = ws.data_ptr(), ws.numel() * 4, ksplit
    d.upmode = 1 if upmode else 0
    d.max_ctas = _CAP.value if max_ctas is None else max_ctas
    d.tile_counter = _ptr(tile_counter if tile_counter is not None else _tile_counter(x.device))
    d.colsum = _ptr(colsum)
    if phases is not None:               # [(ntaps, out_y0, out_x0)]: several output phases in one launch (taps concatenated)
        d.nphases = len(phases)
        for i, (nt, py0, px0) in enumerate(phases):
            d.phase_ntaps[i], d.phase_out_y0[i], d.phase_out_x0[i] = nt, py0, px0
    prof = PROFILE
    with torch.cuda.device(x.device):
        if prof is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        st = _lib.lib().fm_conv_igemm(C.byref(d), _stream())
        if prof is not None:
            e1.record()
            # up mode: algorithmic FLOPs of the stride-2 transposed conv = 9 taps at the INPUT resolution
            # launches that pad their weights with zero blocks pass their algorithmic FLOPs explicitly
            prof.append((e0, e1, algo_flops if algo_flops is not None else
                         2.0 * B * (H * W if upmode else OH * OW) * Cin * Cout * len(taps), PROFILE_TAG, int(d.max_ctas)))
    _lib.check(st, "fm_conv_igemm")
    return out


def prep_weight(w_oikk, scale, cout_rows=None, cin_stride=None, want_wsq=True, out=None):
    """fp32 [O,I,kh,kw] -> (bf16 [kh*kw, rows, cin_stride], fp32 wsq [O,I] or None).
    ``out=(wq, wsq)``: refresh existing buffers in place (their addresses are baked into captured CUDA graphs)."""
    _check_cuda(w_oikk, "weight")
    w = w_oikk.contiguous().to(torch.float32)
    O, I, kh, kw = w.shape
    rows = O if cout_rows is None else cout_rows
    cs = (I + 7) // 8 * 8 if cin_stride is None else cin_stride
    if out is not None:
        wq, wsq = out
        if tuple(wq.shape) != (kh * kw, rows, cs) or (wsq is not None and tuple(wsq.shape) != (O, I)):
            raise RuntimeError("prep_weight: out buffers do not match the weight's shape")
    else:
        wq = torch.empty(kh * kw, rows, cs, device=w.device, dtype=torch.bfloat16)
        wsq = torch.empty(O, I, device=w.device, dtype=torch.float32) if want_wsq else None
    with torch.cuda.device(w.device):
        st = _lib.lib().fm_prep_weight(_ptr(wq), _ptr(wsq), _ptr(w), O, I, kh, kw, float(scale), rows, cs, _stream())
    _lib.check(st, "fm_prep_weight")
    return wq, wsq


def nchw_to_nhwc_bf16(x, scale_bc=None, cstride=None):
    _check_cuda(x, "input")
    x = x.contiguous().to(torch.float32)
    B, Cc, H, W = x.shape
    cs = (Cc + 7) // 8 * 8 if cstride is None else cstride
    out = torch.empty(B, H, W, cs, device=x.device, dtype=torch.bfloat16)
    if scale_bc is not None:
        scale_bc = scale_bc.contiguous().to(torch.float32)
    with torch.cuda.device(x.device):
        st = _lib.lib().fm_nchw_to_nhwc_bf16(_ptr(out), _ptr(x), _ptr(scale_bc), B, Cc, H, W, cs, _stream())
    _lib.check(st, "fm_nchw_to_nhwc_bf16")
    return out


def nhwc_bf16_to_nchw(x, channels, inv_scale_bc=None):
    _check_cuda(x, "input")
    B, H, W, cs = x.shape
    out = torch.empty(B, channels, H, W, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        st = _lib.lib().fm_nhwc_bf16_to_nchw(_ptr(out), _ptr(x), _ptr(inv_scale_bc), B, channels, H, W, cs, _stream())
    _lib.check(st, "fm_nhwc_bf16_to_nchw")
    return out


def _is_rank1(k):
    """True when the (host-inspected) 4x4 kernel is an outer product: enables the separable path."""
    kk = k.detach().double().cpu()
    tot = kk.sum()
    if float(tot.abs()) < 1e-12:
        return False
    return bool(torch.allclose(torch.outer(kk.sum(1), kk.sum(0)) / tot, kk, rtol=1e-6, atol=1e-9))


# keyed on the tensor OBJECT (id + weak reference): a data_ptr key alone would go stale when the allocator
# re-uses the address
_RANK1_CACHE = {}


def blur_act_nhwc(t, kernel4x4, tab, noise, noise_per_sample, noise_w, C_, out=None, padded=False):
    """t bf16 [B,OH+1,OW+1,cs] -> bf16 [B,OH,OW,cs] (see fm_blur_act_nhwc).  padded: t is [B,OH+2,OW+2,cs] with one
    spare row / column per image (pitch only; the spare entries are never read)."""
    B, IH, IW, cs = t.shape
    ph, pw = (IH, IW) if padded else (0, 0)
    if padded:
        IH, IW = IH - 1, IW - 1
    OH, OW = IH - 1, IW - 1
    if out is None:
        out = torch.empty(B, OH, OW, cs, device=t.device, dtype=torch.bfloat16)
    kid = id(kernel4x4)
    ent = _RANK1_CACHE.get(kid)
    if ent is None or ent[0]() is not kernel4x4 or ent[1] != (kernel4x4.data_ptr(), kernel4x4._version):
        if len(_RANK1_CACHE) > 256:
            _RANK1_CACHE.clear()
        # one host sync per kernel buffer (cached)
        ent = _RANK1_CACHE[kid] = (weakref.ref(kernel4x4), (kernel4x4.data_ptr(), kernel4x4._version), _is_rank1(kernel4x4))
    sep = ent[2]
    with torch.cuda.device(t.device):
        st = _lib.lib().fm_blur_act_nhwc(_ptr(out), _ptr(t), _ptr(kernel4x4), _ptr(tab), _ptr(noise),
                                         1 if noise_per_sample else 0, _ptr(noise_w), B, OH, OW, C_, cs,
                                         1 if sep else 0, ph, pw, _stream())
    _lib.check(st, "fm_blur_act_nhwc")
    return out


def rgb_finalize(acc, bias3, skip, kernel4x4, out=None):
    """acc fp32 [B,H,W,4] (+bias +Upsample(skip)) -> fp32 NCHW [B,3,H,W]; acc is zeroed."""
    B, H, W, _ = acc.shape
    if out is None:
        out = torch.empty(B, 3, H, W, device=acc.device, dtype=torch.float32)
    with torch.cuda.device(acc.device):
        st = _lib.lib().fm_rgb_finalize(_ptr(out), _ptr(acc), _ptr(bias3), _ptr(skip), _ptr(kernel4x4), B, H, W, _stream())
    _lib.check(st, "fm_rgb_finalize")
    return out


# ------------------------------------------------------------------ encoder helpers
def _call(name, *args):
    _lib.check(getattr(_lib.lib(), name)(*args), name)


def image_to_nhwc8_padded(x, pad_t, pad_l, Hp, Wp, out=None):
    _check_cuda(x, "input")
    x = x.contiguous().float()
    B, Cc, H, W = x.shape
    if out is None:
        out = torch.empty(B, Hp, Wp, 8, device=x.device, dtype=torch.bfloat16)
    with torch.cuda.device(x.device):
        _call("fm_image_to_nhwc8_padded", _ptr(out), _ptr(x), B, Cc, H, W, pad_t, pad_l, Hp, Wp, _stream())
    return out


def maxpool3x3s2_nhwc(x, out):
    B, H, W, cs = x.shape
    with torch.cuda.device(x.device):
        _call("fm_maxpool3x3s2_nhwc", _ptr(out), _ptr(x), B, H, W, cs, _stream())
    return out


def avgpool_nhwc_to_nchw(x, channels, ph, pw):
    B, H, W, cs = x.shape
    out = torch.empty(B, channels, H // ph, W // pw, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _call("fm_avgpool_nhwc_to_nchw", _ptr(out), _ptr(x), B, H, W, channels, cs, ph, pw, _stream())
    return out


def se_block_nhwc(r, channels, w1, w2, shortcut, sc_stride, sum_buf, gate_buf, out, summed=False):
    """out = r * sigmoid(w2 relu(w1 mean_hw(r))) + shortcut[:, ::s, ::s]  (3 launches; 2 when the conv that produced r
    already accumulated the channel sums into sum_buf: ``conv_igemm(..., colsum=sum_buf)``)."""
    B, H, W, cs = r.shape
    st = _stream()
    with torch.cuda.device(r.device):
        if not summed:
            _call("fm_channel_sum_nhwc", _ptr(sum_buf), _ptr(r), B, H * W, channels, cs, st)
        _call("fm_se_gate", _ptr(gate_buf), _ptr(sum_buf), 1.0 / (H * W), _ptr(w1), _ptr(w2), B, channels, w1.shape[0], st)
        _call("fm_se_combine_nhwc", _ptr(out), _ptr(r), _ptr(gate_buf), _ptr(shortcut), B, H, W, channels, cs,
              shortcut.shape[1], shortcut.shape[2], shortcut.shape[3], sc_stride, st)
    return out


def bilinear_up_nhwc(x, OH, OW, out):
    B, IH, IW, cs = x.shape
    with torch.cuda.device(x.device):
        _call("fm_bilinear_up_nhwc", _ptr(out), _ptr(x), B, IH, IW, OH, OW, cs, _stream())
    return out


def tensor2im_batch(img, cent=1.0, factor=255.0 / 2.0, out=None):
    """fp32 NCHW [B,3,H,W] in [-1,1] -> uint8 NHWC [B,H,W,3] on the device (the reference's ``tensor2im``,
    Evaluation/visual_eval.py:24-38, for the whole batch: clip, shift, scale, truncate, transpose)."""
    _check_cuda(img, "image")
    img = img.contiguous().float()
    B, Cc, H, W = img.shape
    if Cc != 3:
        raise RuntimeError(f"tensor2im_batch expects 3 channels, got {Cc}")
    if out is None:
        out = torch.empty(B, H, W, 3, device=img.device, dtype=torch.uint8)
    with torch.cuda.device(img.device):
        _call("fm_tensor2im_u8", _ptr(out), _ptr(img), B, H, W, float(cent), float(factor), _stream())
    return out


def im2tensor_batch(img_u8, mean=0.5, std=0.5, out=None):
    """uint8 NHWC [B,H,W,3] -> fp32 NCHW [B,3,H,W] = (x / 255 - mean) / std on the device: torchvision's ``ToTensor()`` +
    ``Normalize((mean,)*3, (std,)*3)`` of the reference's transform (train_3_encoder.py:231-237), bit-exact."""
    _check_cuda(img_u8, "image")
    if img_u8.dtype != torch.uint8 or img_u8.ndim != 4 or img_u8.shape[3] != 3:
        raise RuntimeError(f"im2tensor_batch expects uint8 [B,H,W,3], got {img_u8.dtype} {tuple(img_u8.shape)}")
    img_u8 = img_u8.contiguous()
    B, H, W, _ = img_u8.shape
    if out is None:
        out = torch.empty(B, 3, H, W, device=img_u8.device, dtype=torch.float32)
    with torch.cuda.device(img_u8.device):
        _call("fm_im2tensor_f32", _ptr(out), _ptr(img_u8), B, H, W, float(mean), float(std), _stream())
    return out
