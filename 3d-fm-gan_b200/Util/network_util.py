"""3-encoder forward funnel (mirror of the reference's Util/network_util.py:293-338).

``Forward_Inference_3_Encoder`` is the single entry every training step and evaluation loop
of the reference goes through; same name, arguments and return structure here.  The W / W+
product that the reference builds with a 14-iteration Python loop + stack + transpose is one
broadcast multiply.

Every other name of the reference module (``Forward_Inference``, ``Build_Generator_From_Dict``,
``Get_Network_Shape`` ... -- imported by ``train_3_encoder.py:32`` and ``Evaluation/visual_eval.py:16``) is
provided by executing the shadowed reference file in this namespace first (``fm3d/_overlay.py``); those
functions then run on the mirrored ``stylegan2.Generator``.
"""
from fm3d._overlay import load_shadowed

load_shadowed(globals())          # reference names first; the definitions below replace what this path owns

import os  # noqa: E402

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402,F401

_SIDE_STREAMS = {}
_RESNET_PAIR = os.environ.get("FM3D_RESNET_PAIR", "0") != "0"


def _side_streams(device):
    """Two side streams per (device, calling stream): forwards in flight on different streams do not share them."""
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    st = _SIDE_STREAMS.get(key)
    if st is None:
        st = _SIDE_STREAMS[key] = (torch.cuda.Stream(device), torch.cuda.Stream(device))
    return st


_GATES = {}


def _gate(mask, device, dtype):
    """[1,n,1] 0/1 mask of the sliced layers, cached on the device: building it from a Python list every call is a
    pageable host->device copy, i.e. a host-side wait for everything queued on the stream (the encoders)."""
    key = (mask, device, dtype)
    g = _GATES.get(key)
    if g is None:
        if len(_GATES) > 64:
            _GATES.clear()
        g = _GATES[key] = torch.tensor(mask, device=device, dtype=dtype).view(1, len(mask), 1)
    return g


def _n_latent(g_ema):
    return (g_ema.module if hasattr(g_ema, "module") else g_ema).n_latent


def Forward_Inference_3_Encoder(p_input, r_input, E_Tsr, E_W, E_W_Plus, g_ema, tsr_encode='Photo Image',
                                sliced_layer=None, use_tanh=False, PPL_regularize=False):
    """p_input / r_input: normalised photo and rendered-face batches [B,3,H,W].
    E_Tsr -> 4x4 start tensor (from the photo or the render, per ``tsr_encode``), E_W(render) -> W,
    E_W_Plus(photo) -> W+;  latent[b,i] = W[b] * W+[b,i] for i in ``sliced_layer`` (default: all),
    else W[b].  Returns the generator output (``(image, path_lengths)`` under PPL_regularize)."""
    if tsr_encode == 'Photo Image':
        tsr_in = p_input
    elif tsr_encode == 'Render Image':
        tsr_in = r_input
    else:
        raise ValueError(f"unknown tsr_encode {tsr_encode!r}")
    concurrent = (p_input.is_cuda and not torch.is_grad_enabled() and os.environ.get("FM3D_STREAMS", "1") != "0"
                  and not torch.cuda.is_current_stream_capturing())
    if concurrent:
        # the three encoders are independent: the two small ResNets run on side streams and fill the
        # tails / launch gaps of the large W+ encoder
        main = torch.cuda.current_stream(p_input.device)
        s1, s2 = _side_streams(p_input.device)
        s1.wait_stream(main)
        s2.wait_stream(main)
        # SM partition: the ResNets' ~40 launches of 10-40 us each are latency-bound on the whole chip; confined to a
        # few SMs they take about as long, and the W+ encoder's large layers keep the rest instead of time-slicing
        from fm3d import ops as _ops
        cap_r, cap_p, cap_g = _ops.sm_partition(p_input.device)
        pair = None
        if _RESNET_PAIR and tsr_in is r_input:
            # both ResNet-18s encode the render: one grouped launch sequence (fm3d/encoder_engine.py:ResNetPlan)
            from fm3d.encoder_engine import run_resnet_pair
            ma, mb = getattr(E_Tsr, "module", E_Tsr), getattr(E_W, "module", E_W)
            if all(hasattr(m, "_engine_ok") and hasattr(m, "layer1") and m._engine_ok(r_input) for m in (ma, mb)):
                with torch.cuda.stream(s1), _ops.cta_cap(2 * cap_r):
                    pair = run_resnet_pair(ma, mb, r_input)
        if pair is not None:
            encoded_tensor, encoded_W = pair
        else:
            with torch.cuda.stream(s1), _ops.cta_cap(cap_r):
                encoded_tensor = E_Tsr(tsr_in)
            with torch.cuda.stream(s2), _ops.cta_cap(cap_r):
                encoded_W = E_W(r_input)
        with _ops.cta_cap(cap_p):
            encoded_W_plus = E_W_Plus(p_input)
        main.wait_stream(s1)
        main.wait_stream(s2)
        for t in (encoded_tensor, encoded_W, p_input, r_input):
            t.record_stream(main)
    else:
        cap_r = cap_p = cap_g = 0
        if p_input.is_cuda and os.environ.get("FM3D_PARTITION_SERIAL", "0") != "0":
            # measurement aid (bench.py's per-launch roofline pass): the launches of the concurrent path, one stream
            from fm3d import ops as _ops
            cap_r, cap_p, cap_g = _ops.sm_partition(p_input.device)
            with _ops.cta_cap(cap_r):
                encoded_tensor = E_Tsr(tsr_in)
                encoded_W = E_W(r_input)
            with _ops.cta_cap(cap_p):
                encoded_W_plus = E_W_Plus(p_input)
        else:
            encoded_tensor = E_Tsr(tsr_in)
            encoded_W = E_W(r_input)
            encoded_W_plus = E_W_Plus(p_input)

    n = encoded_W_plus.shape[1]
    if sliced_layer is None:
        sliced_layer = range(_n_latent(g_ema))
    gate = _gate(tuple(1.0 if i in sliced_layer else 0.0 for i in range(n)), encoded_W.device, encoded_W.dtype)
    w = encoded_W.unsqueeze(1)
    encoded_latent = w * encoded_W_plus * gate + w * (1.0 - gate)

    if cap_g:
        with _ops.cta_cap(cap_g):
            g_output = g_ema(noise_z=None, latent_styles=[encoded_latent], input_is_latent=True,
                             use_external_input_tensor=True, external_input_tensor=encoded_tensor,
                             PPL_regularize=PPL_regularize)
    else:
        g_output = g_ema(noise_z=None, latent_styles=[encoded_latent], input_is_latent=True,
                         use_external_input_tensor=True, external_input_tensor=encoded_tensor,
                         PPL_regularize=PPL_regularize)
    if use_tanh:
        if PPL_regularize:
            g_output = (torch.tanh(g_output[0]), g_output[1])
        else:
            g_output = torch.tanh(g_output)
    return g_output
