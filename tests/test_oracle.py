"""Pin the CPU oracle (oracle/fm_oracle.py) against golden vectors produced by the unmodified
reference (oracle/make_golden.py) and against analytic known answers.  CPU only."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import fm_oracle as orc


def _sd(g, prefix):
    return {k[len(prefix):]: torch.from_numpy(g[k]) for k in g.files if k.startswith(prefix)}


def test_upfirdn2d_golden():
    g = load_golden("upfirdn2d.npz")
    assert len(g["names"]) >= 12
    for name in g["names"]:
        y = orc.upfirdn2d_ref(torch.from_numpy(g[f"{name}.x"]), torch.from_numpy(g[f"{name}.k"]), *[int(c) for c in g[f"{name}.cfg"]])
        np.testing.assert_allclose(y.numpy(), g[f"{name}.y"], rtol=1e-5, atol=1e-5, err_msg=str(name))
    y = orc.upfirdn2d_api_ref(torch.from_numpy(g["api.x"]), torch.from_numpy(g["api.k"]), up=2, down=1, pad=(2, 1))
    np.testing.assert_allclose(y.numpy(), g["api.y"], rtol=1e-5, atol=1e-5)


def test_analytic_known_answers():
    k = orc.make_kernel_ref([1, 3, 3, 1])
    assert abs(float(k[0, 0]) - 1 / 64) < 1e-9 and abs(float(k[1, 1]) - 9 / 64) < 1e-9
    up = orc.upfirdn2d_api_ref(torch.ones(1, 1, 8, 8), k * 4, up=2, pad=(2, 1))
    assert up.shape == (1, 1, 16, 16) and abs(float(up[0, 0, 0, 0]) - 0.5625) < 1e-6
    assert torch.allclose(up[0, 0, 2:-2, 2:-2], torch.ones(12, 12), atol=1e-6)
    # output-size algebra (SURVEY 8c vii)
    assert orc.upfirdn2d_api_ref(torch.ones(1, 1, 17, 17), k * 4, pad=(1, 1)).shape[-1] == 16
    assert orc.upfirdn2d_api_ref(torch.ones(1, 1, 16, 16), k, pad=(2, 2)).shape[-1] == 17
    assert orc.upfirdn2d_api_ref(torch.ones(1, 1, 16, 16), k, pad=(1, 1)).shape[-1] == 15
    assert orc.upfirdn2d_api_ref(torch.ones(1, 1, 16, 16), k, down=2, pad=(1, 1)).shape[-1] == 8
    x = torch.randn(2, 3, 4, 4); b = torch.randn(3)
    ref = torch.nn.functional.leaky_relu(x + b.view(1, 3, 1, 1), 0.2) * 2 ** 0.5
    assert torch.allclose(orc.fused_leaky_relu_ref(x, b), ref, atol=1e-7)


def test_bias_act_golden():
    g = load_golden("bias_act.npz")
    for name in g["names"]:
        b = torch.from_numpy(g[f"{name}.b"]) if f"{name}.b" in g.files else None
        y = orc.fused_leaky_relu_ref(torch.from_numpy(g[f"{name}.x"]), b)
        np.testing.assert_allclose(y.numpy(), g[f"{name}.y"], rtol=1e-6, atol=1e-6)
        gx, gb = orc.fused_leaky_relu_grads_ref(torch.from_numpy(g[f"{name}.gy"]), y, b is not None)
        np.testing.assert_allclose(gx.numpy(), g[f"{name}.gx"], rtol=1e-6, atol=1e-6)
        if b is not None:
            np.testing.assert_allclose(gb.numpy(), g[f"{name}.gb"], rtol=1e-5, atol=1e-5)


def test_modconv_golden():
    g = load_golden("modconv.npz")
    for name in g["names"]:
        cin, cout, k, sdim, demod, up, down = [int(v) for v in g[f"{name}.cfg"]]
        sd = _sd(g, f"{name}.sd.")
        y, _ = orc.modulated_conv2d_ref(torch.from_numpy(g[f"{name}.x"]), torch.from_numpy(g[f"{name}.style"]),
                                        sd["weight"], sd["modulation.weight"], sd["modulation.bias"],
                                        demodulate=bool(demod), upsample=bool(up), downsample=bool(down),
                                        blur_kernel=sd.get("blur.kernel"))
        np.testing.assert_allclose(y.numpy(), g[f"{name}.y"], rtol=1e-4, atol=2e-5, err_msg=str(name))


def test_generator_small_golden():
    g = load_golden("generator_small.npz")
    sd = _sd(g, "sd.")
    noise = [torch.from_numpy(g[f"noise.{i}"]) for i in range(7)]
    lat, ext = torch.from_numpy(g["latent"]), torch.from_numpy(g["ext"])
    rgbs = orc.generator_forward_ref(sd, latent_styles=[lat], input_is_latent=True, noise=noise,
                                     external_input_tensor=ext, return_rgb_list=True)
    for i, r in enumerate(rgbs):
        np.testing.assert_allclose(r.numpy(), g[f"rgb.{i}"], rtol=1e-4, atol=1e-4)
    _, acts = orc.generator_synthesis_ref(sd, lat, noise, ext, return_acts=True)
    for name, a in acts.items():
        np.testing.assert_allclose(a.numpy(), g[f"act.{name}"], rtol=1e-4, atol=1e-4, err_msg=name)
    y_z = orc.generator_forward_ref(sd, noise_z=[torch.from_numpy(g["z"])], randomize_noise=False)
    np.testing.assert_allclose(y_z.numpy(), g["y_z"], rtol=1e-4, atol=1e-4)
    y_mix = orc.generator_forward_ref(sd, noise_z=[torch.from_numpy(g["z"]), torch.from_numpy(g["z2"])], inject_index=3,
                                      truncation=0.7, truncation_latent=torch.from_numpy(g["mean_latent"]),
                                      randomize_noise=False)
    np.testing.assert_allclose(y_mix.numpy(), g["y_mix"], rtol=1e-4, atol=1e-4)
    # PPL branch: the golden drew randn_like under manual_seed(403) on CPU
    torch.manual_seed(403)
    pl_noise = torch.randn(g["ppl.image"].shape)
    img, pl = orc.generator_ppl_ref(sd, lat, noise, ext, pl_noise)
    np.testing.assert_allclose(img.numpy(), g["ppl.image"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(pl.numpy(), g["ppl.path_lengths"], rtol=1e-3, atol=1e-5)


def _checksum(sd):
    s = a = 0.0
    for k in sorted(sd.keys()):
        v = sd[k].detach().double()
        s += float(v.sum()); a += float(v.abs().sum())
    return np.array([s, a])


def test_discriminator_golden():
    """Weights come from the product mirror's constructor under the golden's seed: the checksum
    proves the mirror consumes the RNG exactly like the reference's Discriminator."""
    import stylegan2
    g = load_golden("discriminator32.npz")
    torch.manual_seed(500)
    d = stylegan2.Discriminator(32)
    with torch.no_grad():
        for n, p in d.named_parameters():
            if n.endswith("bias"):
                p.add_(torch.randn_like(p) * 0.1)
    np.testing.assert_allclose(_checksum(d.state_dict()), g["sd_checksum"], rtol=1e-12)
    y = orc.discriminator_forward_ref({k: v.detach() for k, v in d.state_dict().items()}, torch.from_numpy(g["x"]))
    np.testing.assert_allclose(y.numpy(), g["y"], rtol=1e-4, atol=1e-5)


def _randomize_fused_terms(g, seed):
    gen = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in g.named_parameters():
            if name.endswith("noise.weight") or name.endswith("activate.bias") or \
               (name.endswith(".bias") and "to_rgb" in name and p.ndim == 4):
                p.copy_(torch.randn(p.shape, generator=gen) * 0.1)


def test_generator_cfg1_golden():
    """BASELINE config 1: Generator(256,512,8), batch 4 (first two samples here to bound CPU time)."""
    import stylegan2
    g = load_golden("generator_cfg1.npz")
    torch.manual_seed(0)
    gen = stylegan2.Generator(256, 512, 8, channel_multiplier=2)
    _randomize_fused_terms(gen, 1)
    sd = {k: v.detach() for k, v in gen.state_dict().items()}
    np.testing.assert_allclose(_checksum(sd), g["sd_checksum"], rtol=1e-12)
    rg = torch.Generator().manual_seed(2)
    z = torch.randn(4, 512, generator=rg); lat = torch.randn(4, 14, 512, generator=rg); ext = torch.randn(4, 512, 4, 4, generator=rg)
    rg3 = torch.Generator().manual_seed(3)
    noise = [torch.randn(4, 1, 2 ** ((i + 5) // 2), 2 ** ((i + 5) // 2), generator=rg3) for i in range(13)]
    with torch.no_grad():
        y = orc.generator_forward_ref(sd, latent_styles=[lat[:1]], input_is_latent=True, noise=[n[:1] for n in noise],
                                      external_input_tensor=ext[:1])
        yz = orc.generator_forward_ref(sd, noise_z=[z[:1]], randomize_noise=False)
    np.testing.assert_allclose(y[0].numpy(), g["y_latent.img0"].astype(np.float32), rtol=2e-3, atol=4e-3)
    np.testing.assert_allclose(y[:, :, ::8, ::8].numpy(), g["y_latent.ds8"][:1], rtol=1e-3, atol=2e-4)
    np.testing.assert_allclose(yz[:, :, ::8, ::8].numpy(), g["y_z.ds8"][:1], rtol=1e-3, atol=2e-4)


def test_three_encoder_golden():
    """Encoders + funnel (Util/network_util.py:293-338) vs the reference run, B=2 at 256x256.
    Also proves the encoder mirrors build identical parameters under the same seed."""
    from conftest import build_three_encoder_models, state_checksum
    g = load_golden("three_encoder.npz")
    (e_tsr, e_w, e_wp, gen), p, r, noise = build_three_encoder_models()
    for m, key in ((e_tsr, "cs.e_tsr"), (e_w, "cs.e_w"), (e_wp, "cs.e_wp"), (gen, "cs.g")):
        np.testing.assert_allclose(state_checksum(m.state_dict()), g[key], rtol=1e-12, err_msg=key)
    np.testing.assert_array_equal(p[:, :, ::16, ::16].numpy(), g["p.ds"])
    sds = [{k: v.detach() for k, v in m.state_dict().items()} for m in (e_tsr, e_w, e_wp, gen)]
    with torch.no_grad():
        t = orc.resnet18_forward_ref(sds[0], r[:1], True)
        w = orc.resnet18_forward_ref(sds[1], r[:1], False)
        wp = orc.psp_forward_ref(sds[2], p[:1])
        img = orc.forward_inference_3_encoder_ref(p[:1], r[:1], *sds, noise=[n[:1] for n in noise])
    np.testing.assert_allclose(t.numpy(), g["e_tsr"][:1], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(w.numpy(), g["e_w"][:1], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(wp.numpy(), g["e_wp"][:1], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(img[:, :, ::8, ::8].numpy(), g["img.ds8"][:1], rtol=2e-3, atol=2e-3)
