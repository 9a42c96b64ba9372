import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "3d-fm-gan_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def build_three_encoder_models(device="cpu", B=2):
    """Re-create the models of tests/golden/three_encoder*.npz (same seeds and RNG order as
    oracle/make_golden.py:gen_encoders / gen_encoders_big) from the product mirror classes.  B=2 reproduces the inputs
    of three_encoder.npz, any other B those of three_encoder_b<B>.npz."""
    import types
    import torch
    import resnet_encoder as rn
    import stylegan2
    from psp_encoder_model.encoders import psp_encoders as psp
    torch.manual_seed(600)
    e_tsr = rn.resnet18(tensor_encoding=True).eval()
    e_w = rn.resnet18(tensor_encoding=False).eval()
    e_wp = psp.GradualStyleEncoder(18, 'ir_se', types.SimpleNamespace(input_nc=3, n_styles=14)).eval()
    g = stylegan2.Generator(256, 512, 8, channel_multiplier=2).eval()
    gen = torch.Generator().manual_seed(601)
    with torch.no_grad():
        for name, p in g.named_parameters():
            if name.endswith("noise.weight") or name.endswith("activate.bias") or \
               (name.endswith(".bias") and "to_rgb" in name and p.ndim == 4):
                p.copy_(torch.randn(p.shape, generator=gen) * 0.1)
    gen = torch.Generator().manual_seed(602)
    with torch.no_grad():
        for m in list(e_tsr.modules()) + list(e_w.modules()) + list(e_wp.modules()):
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=gen) * 0.1)
                m.running_var.copy_(1.0 + 0.2 * torch.rand(m.running_var.shape, generator=gen))
                m.weight.copy_(1.0 + 0.1 * torch.randn(m.weight.shape, generator=gen))
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=gen))
    if B != 2:
        gen = torch.Generator().manual_seed(700 + B)
    p = torch.rand(B, 3, 256, 256, generator=gen) * 2 - 1
    r = torch.rand(B, 3, 256, 256, generator=gen) * 2 - 1
    gen3 = torch.Generator().manual_seed(603 if B == 2 else 800 + B)
    noise = [torch.randn(B, 1, 2 ** ((i + 5) // 2), 2 ** ((i + 5) // 2), generator=gen3) for i in range(13)]
    models = [m.to(device) for m in (e_tsr, e_w, e_wp, g)]
    return models, p, r, noise


def state_checksum(sd):
    s = a = 0.0
    for k in sorted(sd.keys()):
        v = sd[k].detach().double().cpu()
        s += float(v.sum())
        a += float(v.abs().sum())
    return np.array([s, a])
