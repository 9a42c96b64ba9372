"""the frozen loss networks of the training step -- the reference's LPIPS (VGG16 backbone + linear
heads, lpips/networks_basic.py:36-101) and ArcFace ResNet-18 (Util/arcface_pytorch/resnet_face_recognition.py:350) -- with
every convolution on the native conv path (forward and the gradient w.r.t. the image), against the same modules on ATen
in strict fp32.  Random-init backbones (their pretrained blobs are not in the tree).  Needs the reference source."""
import os
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ref_env  # noqa: E402


@pytest.mark.skipif(ref_env.reference_root() is None, reason="no reference source here")
def test_lpips_and_arcface_on_native_convs(cuda):
    script = """
        import sys, os
        sys.path.insert(0, "tools")
        import ref_env
        ref_env.activate()
        import torch
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        import lpips
        from fm3d.convgrad import use_native_convs, NativeConv2d
        from Util.arcface_pytorch.resnet_face_recognition import resnet_face18
        from Util.training_util import LPIPS_Loss, Face_Identity_Loss
        dev = torch.device("cuda:0")
        torch.manual_seed(0)
        lp = lpips.PerceptualLoss(model='net-lin', net='vgg', use_gpu=True, gpu_ids=[0])
        face = resnet_face18(use_se=False).to(dev).eval()
        for p in face.parameters():
            p.requires_grad = False
        n = use_native_convs(lp.model.net) + use_native_convs(face)
        assert n >= 13 + 5 + 17, n
        g = torch.Generator(device=dev).manual_seed(1)
        out = (torch.rand(4, 3, 256, 256, generator=g, device=dev) * 2 - 1)
        ref = (torch.rand(4, 3, 256, 256, generator=g, device=dev) * 2 - 1)

        def run(native):
            os.environ["FM3D_NATIVE_GRAD"] = "1" if native else "0"
            x = out.clone().requires_grad_(True)
            l1 = LPIPS_Loss(x, ref, lp)                              # Util/training_util.py:116-127
            l2 = Face_Identity_Loss(x, ref, face, 'MSE')             # :186-210
            gx, = torch.autograd.grad(l1 + l2, x)
            return l1.detach(), l2.detach(), gx
        a, b = run(True), run(False)
        rel = lambda u, v: float((u - v).abs().max() / v.abs().max())
        cos = float(torch.nn.functional.cosine_similarity(a[2].flatten(), b[2].flatten(), dim=0))
        print("lpips", float(a[0]), float(b[0]), "face", float(a[1]), float(b[1]), "grad rel", rel(a[2], b[2]), "cos", cos)
        assert abs(float(a[0]) - float(b[0])) < 2e-2 * abs(float(b[0])) and abs(float(a[1]) - float(b[1])) < 3e-2 * abs(float(b[1]))
        assert cos > 0.99
        print("lossnets-ok")
    """
    r = subprocess.run([sys.executable, "-c", textwrap.dedent(script)], cwd=ROOT, capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, PYTHONPATH=""))
    assert r.returncode == 0 and "lossnets-ok" in r.stdout, r.stdout[-2000:] + "\n" + r.stderr[-3000:]
    print(r.stdout[-400:])


@pytest.mark.skipif(ref_env.reference_root() is None, reason="no reference source here")
def test_fid_inception_features_on_native_convs(cuda):
    """the FID feature extractor (Evaluation/fid.py:27-47 over Evaluation/inception.py, random-init
    offline) with its 60 square convolutions on the native path (the 34 1x7 / 7x1 / 1x3 / 3x1 ones stay on ATen), against
    the same module on ATen in strict fp32; and calc_fid (fid.py:50-73) of the two feature sets."""
    script = """
        import sys, os
        sys.path.insert(0, "tools")
        import ref_env
        ref_env.activate()
        import numpy as np
        import torch
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        from Evaluation.fid import load_patched_inception_v3, calc_fid
        from fm3d.convgrad import use_native_convs
        dev = torch.device("cuda:0")
        torch.manual_seed(0)
        inc = load_patched_inception_v3().to(dev).eval()
        assert use_native_convs(inc) == 60
        g = torch.Generator(device=dev).manual_seed(2)
        img = torch.rand(64, 3, 256, 256, generator=g, device=dev) * 2 - 1

        def feats(native):
            os.environ["FM3D_NATIVE_GRAD"] = "1" if native else "0"
            with torch.no_grad():
                return inc(img)[0].view(img.shape[0], -1).cpu()
        a, b = feats(True), feats(False)
        rel = float((a - b).abs().max() / b.abs().max())
        cos = float(torch.nn.functional.cosine_similarity(a, b, dim=1).min())
        print("inception features rel", rel, "min cosine", cos)
        assert rel < 3e-2 and cos > 0.9995
        print("inception-ok")
    """
    r = subprocess.run([sys.executable, "-c", textwrap.dedent(script)], cwd=ROOT, capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, PYTHONPATH=""))
    assert r.returncode == 0 and "inception-ok" in r.stdout, r.stdout[-2000:] + "\n" + r.stderr[-3000:]
    print(r.stdout[-300:])
