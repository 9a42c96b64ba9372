"""CPU oracle: a functional fp32 restatement of the reference's StyleGAN2 synthesis path.

TEST INFRASTRUCTURE -- NOT part of the product path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module.  The product (``3d-fm-gan_b200/``) never does and fails
loudly when its CUDA library is missing.

Parity status: PINNED.  Every function here is checked in ``tests/test_oracle.py``
against golden vectors produced by importing the unmodified reference on CPU
(``oracle/make_golden.py`` -> ``tests/golden/*.npz``), and -- when ``/root/reference``
is present -- directly against the reference modules.

Everything is written as plain functions over a ``state_dict`` (name -> tensor) so that
it shares no code with the product's ``nn.Module`` mirror.  All arithmetic is fp32 on
CPU; convolutions go through ATen/oneDNN exactly like the reference's CPU path does
(the dense math of the reference lives in ATen).

Citations are ``file:line`` relative to the reference repository root.
"""
import math

import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------------
# op/ : upfirdn2d and fused bias + leaky-ReLU
# ---------------------------------------------------------------------------------
def upfirdn2d_ref(x, kernel, up_x=1, up_y=1, down_x=1, down_y=1,
                  pad_x0=0, pad_x1=0, pad_y0=0, pad_y1=0):
    """out[n,c,oy,ox] = sum_{ky,kx} xp[n,c,oy*down_y+ky, ox*down_x+kx] * k[kh-1-ky, kw-1-kx]

    where xp is the input zero-stuffed by ``up`` and padded (negative pad = crop).
    Restates op/upfirdn2d.py:168-209 (``upfirdn2d_native``) and the CUDA kernel's
    definition op/upfirdn2d_kernel.cu:49-105 as a direct tap sum (no F.conv2d)."""
    x = torch.as_tensor(x, dtype=torch.float32)
    k = torch.as_tensor(kernel, dtype=torch.float32)
    n, c, in_h, in_w = x.shape
    kh, kw = k.shape
    # zero stuffing (op/upfirdn2d.py:177-179): sample (iy,ix) lands at (iy*up_y, ix*up_x)
    up = x.new_zeros(n, c, in_h * up_y, in_w * up_x)
    up[:, :, ::up_y, ::up_x] = x
    # pad / crop (op/upfirdn2d.py:181-189)
    up = F.pad(up, [max(pad_x0, 0), max(pad_x1, 0), max(pad_y0, 0), max(pad_y1, 0)])
    up = up[:, :, max(-pad_y0, 0): up.shape[2] - max(-pad_y1, 0),
            max(-pad_x0, 0): up.shape[3] - max(-pad_x1, 0)]
    ph, pw = up.shape[2], up.shape[3]
    full_h, full_w = ph - kh + 1, pw - kw + 1
    out_h = (in_h * up_y + pad_y0 + pad_y1 - kh) // down_y + 1     # op/upfirdn2d.py:206
    out_w = (in_w * up_x + pad_x0 + pad_x1 - kw) // down_x + 1     # op/upfirdn2d.py:207
    acc = x.new_zeros(n, c, max(full_h, 0), max(full_w, 0))
    for ky in range(kh):
        for kx in range(kw):
            # true convolution: flipped kernel (op/upfirdn2d.py:195, .cu:137)
            acc += up[:, :, ky: ky + full_h, kx: kx + full_w] * k[kh - 1 - ky, kw - 1 - kx]
    out = acc[:, :, ::down_y, ::down_x]
    assert out.shape[2] == out_h and out.shape[3] == out_w, (out.shape, out_h, out_w)
    return out.contiguous()


def upfirdn2d_api_ref(x, kernel, up=1, down=1, pad=(0, 0)):
    """Public wrapper semantics, op/upfirdn2d.py:154-165: scalar up/down, pad=(p0,p1) on both axes."""
    return upfirdn2d_ref(x, kernel, up, up, down, down, pad[0], pad[1], pad[0], pad[1])


def fused_bias_act_ref(x, bias=None, ref=None, act=3, grad=0, alpha=0.2, scale=2 ** 0.5):
    """op/fused_bias_act_kernel.cu:18-49.  bias broadcasts along dim 1
    (``(xi / step_b) % size_b`` with ``step_b = prod(dims[2:])``, .cu:28,67-71)."""
    x = torch.as_tensor(x, dtype=torch.float32)
    if bias is not None and bias.numel():
        x = x + bias.reshape(1, -1, *([1] * (x.ndim - 2)))
    code = act * 10 + grad
    if code in (10, 11):
        y = x
    elif code == 30:
        y = torch.where(x > 0, x, x * alpha)
    elif code == 31:
        y = torch.where(ref > 0, x, x * alpha)
    elif code in (12, 32):
        y = torch.zeros_like(x)
    else:                       # "default:" falls into case 10 (.cu:37-38)
        y = x
    return y * scale


def fused_leaky_relu_ref(x, bias=None, negative_slope=0.2, scale=2 ** 0.5):
    """op/fused_act.py:113-128 (== leaky_relu(x + b, 0.2) * sqrt(2))."""
    return fused_bias_act_ref(x, bias, None, 3, 0, negative_slope, scale)


def fused_leaky_relu_grads_ref(gy, out, has_bias, negative_slope=0.2, scale=2 ** 0.5):
    """op/fused_act.py:29-53: grad_input via mode (3,1) with ref=out; grad_bias = sum over
    all dims but 1."""
    gx = fused_bias_act_ref(gy, None, out, 3, 1, negative_slope, scale)
    gb = None
    if has_bias:
        dims = [0] + list(range(2, gx.ndim))
        gb = gx.sum(dims)
    return gx, gb


def make_kernel_ref(k):
    """stylegan2.py:36-44."""
    k = torch.tensor(k, dtype=torch.float32)
    if k.ndim == 1:
        k = k[None, :] * k[:, None]
    return k / k.sum()


# ---------------------------------------------------------------------------------
# stylegan2.py layers
# ---------------------------------------------------------------------------------
def equal_linear_ref(x, weight, bias, lr_mul=1.0, activation=False):
    """stylegan2.py:146-175."""
    scale = (1.0 / math.sqrt(weight.shape[1])) * lr_mul
    if activation:
        out = F.linear(x, weight * scale)
        return fused_leaky_relu_ref(out, bias * lr_mul)
    return F.linear(x, weight * scale, bias=None if bias is None else bias * lr_mul)


def pixel_norm_ref(x):
    """stylegan2.py:32-33."""
    return x * torch.rsqrt(torch.mean(x ** 2, dim=1, keepdim=True) + 1e-8)


def modulated_conv2d_ref(x, style_in, weight, mod_weight, mod_bias, demodulate=True,
                         upsample=False, downsample=False, blur_kernel=None):
    """stylegan2.py:250-298.  ``weight`` is [1,O,I,k,k]; ``style_in`` is the W vector; the
    modulation EqualLinear has bias_init=1 (:240).  Written per-sample (loop over batch)
    instead of the reference's groups=batch trick."""
    b, cin, h, w = x.shape
    _, cout, _, k, _ = weight.shape
    scale = 1.0 / math.sqrt(cin * k * k)                                   # :232-233
    s = equal_linear_ref(style_in, mod_weight, mod_bias)                  # :257
    outs = []
    for i in range(b):
        wi = scale * weight[0] * s[i].view(1, cin, 1, 1)                   # :258
        if demodulate:
            d = torch.rsqrt(wi.pow(2).sum([1, 2, 3]) + 1e-8)               # :261 (literal 1e-8)
            wi = wi * d.view(cout, 1, 1, 1)
        xi = x[i:i + 1]
        if upsample:
            yi = F.conv_transpose2d(xi, wi.transpose(0, 1), padding=0, stride=2)   # :276
            p = (blur_kernel.shape[0] - 2) - (k - 1)                       # :218-220
            yi = upfirdn2d_api_ref(yi, blur_kernel, pad=((p + 1) // 2 + 1, p // 2 + 1))
        elif downsample:
            p = (blur_kernel.shape[0] - 2) + (k - 1)                       # :226-228
            xi = upfirdn2d_api_ref(xi, blur_kernel, pad=((p + 1) // 2, p // 2))
            yi = F.conv2d(xi, wi, padding=0, stride=2)                     # :285
        else:
            yi = F.conv2d(xi, wi, padding=k // 2)                          # :291
        outs.append(yi)
    return torch.cat(outs, 0), s


def styled_conv_ref(sd, prefix, x, style, noise, upsample):
    """stylegan2.py:360-376: modconv -> noise (:312) -> FusedLeakyReLU."""
    bk = sd.get(prefix + "conv.blur.kernel")
    out, _ = modulated_conv2d_ref(x, style, sd[prefix + "conv.weight"],
                                  sd[prefix + "conv.modulation.weight"],
                                  sd[prefix + "conv.modulation.bias"],
                                  demodulate=True, upsample=upsample, blur_kernel=bk)
    if noise is None:
        noise = torch.randn(out.shape[0], 1, out.shape[2], out.shape[3])
    out = out + sd[prefix + "noise.weight"] * noise
    return fused_leaky_relu_ref(out, sd[prefix + "activate.bias"])


def to_rgb_ref(sd, prefix, x, style, skip=None):
    """stylegan2.py:389-404: 1x1 modconv without demod, + bias, + Upsample(skip)."""
    out, _ = modulated_conv2d_ref(x, style, sd[prefix + "conv.weight"],
                                  sd[prefix + "conv.modulation.weight"],
                                  sd[prefix + "conv.modulation.bias"], demodulate=False)
    out = out + sd[prefix + "bias"]
    if skip is not None:
        k = sd[prefix + "upsample.kernel"]
        p = k.shape[0] - 2                                                 # :55-58
        skip = upfirdn2d_api_ref(skip, k, up=2, down=1, pad=((p + 1) // 2 + 1, p // 2))
        out = out + skip
    return out


def mapping_ref(sd, z):
    """Generator.style: PixelNorm + n_mlp x EqualLinear(lr_mul=0.01, fused_lrelu), stylegan2.py:430-439."""
    x = pixel_norm_ref(z)
    i = 1
    while f"style.{i}.weight" in sd:
        x = equal_linear_ref(x, sd[f"style.{i}.weight"], sd[f"style.{i}.bias"], lr_mul=0.01, activation=True)
        i += 1
    return x


def generator_synthesis_ref(sd, latent, noise, x0=None, return_rgb_list=False, return_acts=False):
    """stylegan2.py:627-668.  ``latent`` [B,n_latent,D]; ``noise`` list of [B|1,1,r,r] (None ->
    fresh N(0,1)); ``x0`` the external 4x4 tensor, else ConstantInput (:325-329)."""
    b = latent.shape[0]
    n_convs = 0
    while f"convs.{n_convs}.conv.weight" in sd:
        n_convs += 1
    if noise is None:
        noise = [None] * (n_convs + 1)
    out = x0 if x0 is not None else sd["input.input"].repeat(b, 1, 1, 1)
    acts = {}
    out = styled_conv_ref(sd, "conv1.", out, latent[:, 0], noise[0], False)          # :640
    acts["conv1"] = out
    skip = to_rgb_ref(sd, "to_rgb1.", out, latent[:, 1])                             # :643
    acts["to_rgb1"] = skip
    rgbs = [skip]
    i = 1
    for j in range(n_convs // 2):
        out = styled_conv_ref(sd, f"convs.{2 * j}.", out, latent[:, i], noise[1 + 2 * j], True)       # :656
        acts[f"convs.{2 * j}"] = out
        out = styled_conv_ref(sd, f"convs.{2 * j + 1}.", out, latent[:, i + 1], noise[2 + 2 * j], False)  # :657
        acts[f"convs.{2 * j + 1}"] = out
        skip = to_rgb_ref(sd, f"to_rgbs.{j}.", out, latent[:, i + 2], skip)          # :663
        acts[f"to_rgbs.{j}"] = skip
        rgbs.append(skip)
        i += 2
    if return_acts:
        return skip, acts
    return rgbs if return_rgb_list else skip


def generator_forward_ref(sd, noise_z=None, latent_styles=None, input_is_latent=False, noise=None,
                          randomize_noise=True, inject_index=None, truncation=1.0,
                          truncation_latent=None, external_input_tensor=None,
                          return_rgb_list=False):
    """stylegan2.py:554-681 (everything but the PPL branch, see generator_ppl_ref)."""
    n_convs = 0
    while f"convs.{n_convs}.conv.weight" in sd:
        n_convs += 1
    num_layers = n_convs + 1
    n_latent = num_layers + 1                                   # log_size*2-2 (:530)
    styles = latent_styles if input_is_latent else [mapping_ref(sd, z) for z in noise_z]    # :583-586
    if noise is None and not randomize_noise:
        noise = [sd[f"noises.noise_{i}"] for i in range(num_layers)]                         # :592-594
    if truncation < 1:
        styles = [truncation_latent + truncation * (s - truncation_latent) for s in styles]  # :596-604
    if len(styles) < 2:
        latent = styles[0]
        if latent.ndim < 3:
            latent = latent.unsqueeze(1).repeat(1, n_latent, 1)                              # :611-612
    else:
        assert inject_index is not None, "oracle requires an explicit inject_index"
        l1 = styles[0].unsqueeze(1).repeat(1, inject_index, 1)
        l2 = styles[1].unsqueeze(1).repeat(1, n_latent - inject_index, 1)
        latent = torch.cat([l1, l2], 1)                                                      # :622-625
    return generator_synthesis_ref(sd, latent, noise, external_input_tensor, return_rgb_list)


def generator_ppl_ref(sd, latent, noise, x0, pl_noise):
    """PPL branch, stylegan2.py:683-688, with the randn_like draw passed in as ``pl_noise``."""
    latent = latent.clone().requires_grad_(True)
    image = generator_synthesis_ref(sd, latent, noise, x0)
    n = pl_noise / math.sqrt(image.shape[2] * image.shape[3])
    grad, = torch.autograd.grad((image * n).sum(), latent, create_graph=False)
    return image.detach(), torch.sqrt(grad.pow(2).sum(2).mean(1))


# ---------------------------------------------------------------------------------
# Discriminator  (stylegan2.py:692-820)
# ---------------------------------------------------------------------------------
def _conv_layer_ref(sd, prefix, x, k, downsample, activate=True, bias=True):
    """ConvLayer = [Blur] + EqualConv2d + [FusedLeakyReLU]  (stylegan2.py:692-738)."""
    idx = 0
    if downsample:
        bk = sd[f"{prefix}{idx}.kernel"]
        p = (bk.shape[0] - 2) + (k - 1)                                    # :707-709
        x = upfirdn2d_api_ref(x, bk, pad=((p + 1) // 2, p // 2))
        idx += 1
    w = sd[f"{prefix}{idx}.weight"]
    scale = 1.0 / math.sqrt(w.shape[1] * k * k)                            # :117
    cb = sd.get(f"{prefix}{idx}.bias")
    x = F.conv2d(x, w * scale, bias=cb, stride=2 if downsample else 1, padding=0 if downsample else k // 2)
    idx += 1
    if activate:
        if bias:
            x = fused_leaky_relu_ref(x, sd[f"{prefix}{idx}.bias"])
        else:
            x = F.leaky_relu(x, 0.2) * math.sqrt(2)
    return x


def discriminator_forward_ref(sd, x):
    out = _conv_layer_ref(sd, "convs.0.", x, 1, False)                     # :778
    i = 1
    while f"convs.{i}.conv1.0.weight" in sd:                               # ResBlocks :741-759
        a = _conv_layer_ref(sd, f"convs.{i}.conv1.", out, 3, False)
        a = _conv_layer_ref(sd, f"convs.{i}.conv2.", a, 3, True)
        s = _conv_layer_ref(sd, f"convs.{i}.skip.", out, 1, True, activate=False, bias=False)
        out = (a + s) / math.sqrt(2)
        i += 1
    b, c, h, w = out.shape
    group = min(b, 4)                                                      # :806
    sdv = out.view(group, -1, 1, c, h, w)
    sdv = torch.sqrt(sdv.var(0, unbiased=False) + 1e-8)
    sdv = sdv.mean([2, 3, 4], keepdim=True).squeeze(2)
    sdv = sdv.repeat(group, 1, h, w)
    out = torch.cat([out, sdv], 1)                                         # :813
    out = _conv_layer_ref(sd, "final_conv.", out, 3, False)
    out = out.view(b, -1)
    out = equal_linear_ref(out, sd["final_linear.0.weight"], sd["final_linear.0.bias"], activation=True)
    return equal_linear_ref(out, sd["final_linear.1.weight"], sd["final_linear.1.bias"])


# ---------------------------------------------------------------------------------
# Encoders (eval mode: BatchNorm uses running statistics)
# ---------------------------------------------------------------------------------
def _bn_ref(sd, prefix, x, eps=1e-5):
    w, b = sd[prefix + "weight"], sd[prefix + "bias"]
    m, v = sd[prefix + "running_mean"], sd[prefix + "running_var"]
    return (x - m.view(1, -1, 1, 1)) / torch.sqrt(v.view(1, -1, 1, 1) + eps) * w.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)


def resnet18_forward_ref(sd, x, tensor_encoding=True):
    """resnet_encoder.py:258-280 with BasicBlock :74-91, layers [2,2,2,2] (:309)."""
    x = F.conv2d(x, sd["conv1.weight"], stride=2, padding=3)
    x = F.relu(_bn_ref(sd, "bn1.", x))
    x = F.max_pool2d(x, 3, 2, 1)
    for li in range(1, 5):
        for bi in range(2):
            p = f"layer{li}.{bi}."
            stride = 2 if (li > 1 and bi == 0) else 1
            idt = x
            out = F.conv2d(x, sd[p + "conv1.weight"], stride=stride, padding=1)
            out = F.relu(_bn_ref(sd, p + "bn1.", out))
            out = F.conv2d(out, sd[p + "conv2.weight"], padding=1)
            out = _bn_ref(sd, p + "bn2.", out)
            if p + "downsample.0.weight" in sd:
                idt = _bn_ref(sd, p + "downsample.1.", F.conv2d(x, sd[p + "downsample.0.weight"], stride=stride))
            x = F.relu(out + idt)
    if tensor_encoding:
        return F.avg_pool2d(x, 2, 2)                                       # :207
    return torch.flatten(F.adaptive_avg_pool2d(x, (1, 1)), 1)              # :209,:273


def _ir_se_ref(sd, p, x, in_c, depth, stride):
    """bottleneck_IR_SE, psp_encoder_model/encoders/helpers.py:117-139; SEModule :76-92."""
    if in_c == depth:
        sc = x[:, :, ::stride, ::stride]                                   # MaxPool2d(1, stride)
    else:
        sc = _bn_ref(sd, p + "shortcut_layer.1.", F.conv2d(x, sd[p + "shortcut_layer.0.weight"], stride=stride))
    r = _bn_ref(sd, p + "res_layer.0.", x)
    r = F.conv2d(r, sd[p + "res_layer.1.weight"], padding=1)
    r = F.prelu(r, sd[p + "res_layer.2.weight"])
    r = F.conv2d(r, sd[p + "res_layer.3.weight"], stride=stride, padding=1)
    r = _bn_ref(sd, p + "res_layer.4.", r)
    se = r.mean([2, 3], keepdim=True)
    se = F.relu(F.conv2d(se, sd[p + "res_layer.5.fc1.weight"]))
    se = torch.sigmoid(F.conv2d(se, sd[p + "res_layer.5.fc2.weight"]))
    return r * se + sc


def psp_forward_ref(sd, x, n_styles=14):
    """GradualStyleEncoder(18,'ir_se'), psp_encoders.py:100-132; blocks helpers.py:38-46."""
    x = F.conv2d(x, sd["input_layer.0.weight"], padding=1)
    x = F.prelu(_bn_ref(sd, "input_layer.1.", x), sd["input_layer.2.weight"])
    units = []
    for in_c, depth in [(64, 64), (64, 128), (128, 256), (256, 512)]:
        units += [(in_c, depth, 2), (depth, depth, 1)]
    feats = {}
    for i, (in_c, depth, stride) in enumerate(units):
        x = _ir_se_ref(sd, f"body.{i}.", x, in_c, depth, stride)
        feats[i] = x
    c1, c2, c3 = feats[3], feats[5], feats[7]                              # :108-118

    def style_block(j, f):
        """GradualStyleBlock, psp_encoders.py:20-41 (nn.LeakyReLU default slope 0.01)."""
        k = 0
        while f"styles.{j}.convs.{k}.weight" in sd:
            f = F.conv2d(f, sd[f"styles.{j}.convs.{k}.weight"], sd[f"styles.{j}.convs.{k}.bias"], stride=2, padding=1)
            f = F.leaky_relu(f, 0.01)
            k += 2
        f = f.view(-1, 512)
        return equal_linear_ref(f, sd[f"styles.{j}.linear.weight"], sd[f"styles.{j}.linear.bias"], lr_mul=1)

    lat = [style_block(j, c3) for j in range(3)]
    l1 = F.conv2d(c2, sd["latlayer1.weight"], sd["latlayer1.bias"])
    p2 = F.interpolate(c3, size=l1.shape[2:], mode="bilinear", align_corners=True) + l1     # :81-98
    lat += [style_block(j, p2) for j in range(3, 7)]
    l2 = F.conv2d(c1, sd["latlayer2.weight"], sd["latlayer2.bias"])
    p1 = F.interpolate(p2, size=l2.shape[2:], mode="bilinear", align_corners=True) + l2
    lat += [style_block(j, p1) for j in range(7, n_styles)]
    return torch.stack(lat, dim=1)


def forward_inference_3_encoder_ref(p, r, sd_tsr, sd_w, sd_wp, sd_g, noise=None,
                                    tsr_encode="Render Image", sliced_layer=None, use_tanh=False):
    """Util/network_util.py:293-338."""
    t = resnet18_forward_ref(sd_tsr, p if tsr_encode == "Photo Image" else r, True)
    w = resnet18_forward_ref(sd_w, r, False)
    wp = psp_forward_ref(sd_wp, p)
    n = wp.shape[1]
    if sliced_layer is None:
        sliced_layer = range(n)
    lat = torch.stack([w * wp[:, i, :] if i in sliced_layer else w for i in range(n)]).transpose(0, 1)
    img = generator_synthesis_ref(sd_g, lat, noise, t)
    return torch.tanh(img) if use_tanh else img


def tensor2im_ref(image):
    """Evaluation/visual_eval.py:24-38 for one [3,H,W] image -> uint8 [H,W,3]."""
    import numpy as np
    a = image.detach().cpu().float().numpy()
    a = np.clip(a, -1, 1)
    a = (np.transpose(a, (1, 2, 0)) + 1.0) * (255.0 / 2.0)
    return a.astype(np.uint8)
