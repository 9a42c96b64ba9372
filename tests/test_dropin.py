"""The drop-in boundary of SURVEY 8(b), exercised the way INTEGRATION.md prescribes it.

CPU part (needs the reference checkout, so it is skipped on a box without one): the reference's *callers*
-- ``train_3_encoder.py`` (valid prefix, lines 1-879) and ``Evaluation/visual_eval.py`` -- import unchanged with
the mirror in front of the reference on ``sys.path``, every name of ``train_3_encoder.py:26-37`` resolves, and the
training-step functions are bound to the mirrored classes.  Each scenario runs in a fresh interpreter because it
rearranges ``sys.path`` / ``sys.modules``.

GPU part: the reference's own ``op/*.py`` on the ctypes stub of INTEGRATION.md section 2 against the mirror ops.
"""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ref_env  # noqa: E402

needs_reference = pytest.mark.skipif(ref_env.reference_root() is None, reason="no reference checkout here")


def _run(script):
    env = dict(os.environ, PYTHONPATH="")
    r = subprocess.run([sys.executable, "-c", textwrap.dedent(script)], cwd=ROOT, env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout + "\n" + r.stderr
    return r.stdout


@needs_reference
def test_train_script_and_visual_eval_import_unchanged():
    """PYTHONPATH = mirror : reference.  The five imports the round-1 verdict found broken
    (train_3_encoder.py:32-37, visual_eval.py:16) plus the rest of lines 26-37."""
    out = _run("""
        import sys
        sys.path.insert(0, "tools")
        import ref_env
        ref = ref_env.activate()
        mirror = ref_env.PKG

        # -- the imports of train_3_encoder.py:26-37, verbatim
        from stylegan2 import Generator, Discriminator
        from resnet_encoder import resnet18
        from psp_encoder_model.encoders import psp_encoders
        from dataset import Synthetic_Dataset, FFHQ_Dataset_Reconstruction, FFHQ_Dataset_Editing, FFHQ_Dataset, DualSupervisionSampler, ExtremePoseDualSupervisionSampler, Data_Loading
        from Miscellaneous.distributed import reduce_loss_dict
        from Util.network_util import Build_Generator_From_Dict, Forward_Inference_3_Encoder
        from Evaluation.quant_eval import Get_Recon_Score, Get_Edit_Score
        from Evaluation.visual_eval import Get_Real_Img_Val_Sample, Get_Syn_Img_Val_Sample, Get_Batch_Eval_Result
        from Evaluation.fid import load_patched_inception_v3
        import lpips
        from Util.training_util import d_logistic_loss, d_r1_loss, g_nonsaturating_loss, L1_Loss, LPIPS_Loss, Face_Identity_Loss, Load_Face_Recognition_Network, Heat_Map_Loss, Face_Regional_Loss
        # -- Evaluation/visual_eval.py:16
        from Util.network_util import Forward_Inference, Forward_Inference_3_Encoder
        import Util.training_util, Evaluation.quant_eval

        import stylegan2, resnet_encoder, op, Util.network_util as nu, Evaluation.visual_eval as ve, Miscellaneous.distributed as dist
        import dataset
        inside = lambda m, root: m.__file__.startswith(root + "/")
        # what this path owns comes from the mirror ...
        for m in (stylegan2, resnet_encoder, op, psp_encoders, nu, ve, dist, dataset):
            assert inside(m, mirror), m.__file__
        # ... and everything else from the reference checkout
        for m in (lpips, Util.training_util, Evaluation.quant_eval, sys.modules["Evaluation.fid"]):
            assert inside(m, ref), m.__file__
        assert nu.__shadowed_file__ == ref + "/Util/network_util.py"
        assert ve.__shadowed_file__ == ref + "/Evaluation/visual_eval.py"
        assert dataset.__shadowed_file__ == ref + "/dataset.py"
        # dataset classes are the reference's, Data_Loading is the device-side one
        assert Synthetic_Dataset.__module__ == "dataset" and DualSupervisionSampler.__init__.__code__.co_filename.startswith(ref)
        assert Data_Loading.__code__.co_filename.startswith(mirror)
        # the funnel and tensor2im are the mirror's; the reference's loops call them
        assert Forward_Inference_3_Encoder.__code__.co_filename.startswith(mirror)
        assert ve.tensor2im.__code__.co_filename.startswith(mirror)
        assert ve.Get_Single_Eval_Result.__code__.co_filename.startswith(ref)
        assert ve.Get_Single_Eval_Result.__globals__["tensor2im"] is ve.tensor2im
        assert ve.Get_Single_Eval_Result.__globals__["Forward_Inference_3_Encoder"] is nu.Forward_Inference_3_Encoder
        assert Build_Generator_From_Dict.__code__.co_filename.startswith(ref)
        assert Build_Generator_From_Dict.__globals__["Generator"] is stylegan2.Generator

        # -- the training script itself (valid prefix): step functions bound to the mirrored classes
        ts = ref_env.load_train_script()
        for fn in ("D_Loss_BackProp", "D_Reg_BackProp", "G_Loss_BackProp", "G_Reg_BackProp", "accumulate", "train",
                   "Module_To_Train_Setup", "Optimizer_Initilization", "Sample_Eval_Save_Ckpt"):
            assert callable(getattr(ts, fn)), fn
        g = ts.D_Loss_BackProp.__globals__
        assert g["Forward_Inference_3_Encoder"] is nu.Forward_Inference_3_Encoder
        assert g["Generator"] is stylegan2.Generator and g["Discriminator"] is stylegan2.Discriminator
        assert g["resnet18"] is resnet_encoder.resnet18
        assert ts.G_Reg_BackProp.__globals__["psp_encoders"] is psp_encoders

        # a pruned generator rebuilt from a state dict by the reference's helper is a mirror Generator (CPU: no kernels run)
        small = stylegan2.Generator(32, 64, 2, generator_net_shape=[48, 40, 32, 24, 24, 16, 16, 8])
        g2 = Build_Generator_From_Dict(small.state_dict(), size=32, latent=64, n_mlp=2)
        assert type(g2) is stylegan2.Generator and nu.Get_Network_Shape(g2.state_dict()) == [48, 40, 32, 24, 24, 16, 16, 8]
        print("dropin-ok")
    """)
    assert "dropin-ok" in out


@needs_reference
def test_reference_op_package_binds_ctypes_stub():
    """the reference's own op/*.py with cpp_extension.load routed to libfm3d (no JIT)."""
    out = _run("""
        import sys, torch
        sys.path.insert(0, "tools"); sys.path.insert(0, "3d-fm-gan_b200")
        import ref_env
        ref = ref_env.reference_root()
        import fm3d.pybind_compat as pc
        pc.install()
        sys.path.insert(0, ref)                  # the reference's packages now shadow the mirror's
        for k in [k for k in sys.modules if k == "op" or k.startswith("op.")]:
            del sys.modules[k]
        import op
        assert op.__file__.startswith(ref + "/"), op.__file__
        # (op/__init__.py re-exports the *function* upfirdn2d over the submodule name)
        fused_act, up_mod = sys.modules["op.fused_act"], sys.modules["op.upfirdn2d"]
        assert fused_act.fused is pc.fused and up_mod.upfirdn2d_op is pc.upfirdn2d_op
        # the reference's CPU branch still works and never reaches the stub
        x = torch.randn(2, 4, 5, 5)
        y = op.fused_leaky_relu(x, torch.zeros(4))
        assert torch.allclose(y, torch.nn.functional.leaky_relu(x, 0.2) * 2 ** 0.5)
        pc.uninstall()
        print("stub-ok")
    """)
    assert "stub-ok" in out


@pytest.mark.gpu
@needs_reference
def test_reference_op_on_stub_matches_mirror_gpu(cuda):
    """Reference autograd classes over the ctypes stub vs the mirror ops: forward, backward, double backward."""
    out = _run("""
        import sys, torch
        sys.path.insert(0, "tools"); sys.path.insert(0, "3d-fm-gan_b200")
        import ref_env
        ref = ref_env.reference_root()
        import op as mirror_op
        import fm3d.pybind_compat as pc
        pc.install()
        sys.path.insert(0, ref)
        saved = {k: sys.modules.pop(k) for k in [k for k in sys.modules if k == "op" or k.startswith("op.")]}
        import op as ref_op
        assert ref_op.__file__.startswith(ref + "/")
        dev = torch.device("cuda:0")
        torch.manual_seed(0)
        x = torch.randn(3, 8, 17, 17, device=dev, requires_grad=True)
        b = torch.randn(8, device=dev, requires_grad=True)
        k = torch.tensor([1., 3., 3., 1.], device=dev); k = torch.outer(k, k); k = k / k.sum() * 4
        def run(o):
            y = o.fused_leaky_relu(o.upfirdn2d(x, k, up=2, pad=(2, 1)), b)
            y = o.upfirdn2d(y, k / 4, down=2, pad=(1, 1))
            g, = torch.autograd.grad(y.sum(), x, create_graph=True)
            gg = torch.autograd.grad(g.pow(2).sum(), [x, b])
            return [y.detach(), g.detach()] + [t.detach() for t in gg]
        for a, c in zip(run(ref_op), run(mirror_op)):
            assert torch.allclose(a, c, rtol=1e-5, atol=1e-5), float((a - c).abs().max())
        print("gpu-stub-ok")
    """)
    assert "gpu-stub-ok" in out
