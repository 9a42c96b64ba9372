"""BASELINE config 4 on the GPU: the reference's OWN four update functions (D, D-R1, G, G-path-length;
train_3_encoder.py:448-596, loaded from its source by tools/train_step.py) run end to end on the mirrored modules with
every convolution of G and D on the tcgen05 kernels, a tiny batch so the test stays in seconds.  Needs the reference
source (baseline/_ref on a GPU box, tools/stage_reference.sh)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ref_env  # noqa: E402

needs_reference = pytest.mark.skipif(ref_env.reference_root() is None, reason="no reference source here")


def _run(env_extra=None, iters=4, warmup=1):
    env = dict(os.environ, **(env_extra or {}))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "train_step.py"), "--size", "256", "--batch", "4",
                          "--iters", str(iters), "--warmup", str(warmup), "--d-reg-every", "2", "--g-reg-every", "2"],
                         capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stderr[-3000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


@needs_reference
def test_reference_train_functions_run_on_native_kernels(cuda):
    line = _run()
    assert line["native_grad"] and line["source"].endswith("train_3_encoder.py")
    assert line["losses_finite"], line
    assert set(line["losses"]) >= {"d", "r1", "g", "l1", "g_reg", "face_reg", "lpips", "face_id"}
    assert line["value"] > 0
    # the FIRST iteration from identical weights and data, native kernels vs the ATen cross-check path: the losses agree
    # to bf16 accuracy (later iterations of an adversarial game drift apart chaotically, so they are not compared)
    one = _run(iters=1, warmup=0)
    ref = _run({"FM3D_NATIVE_GRAD": "0"}, iters=1, warmup=0)
    assert one["native_grad"] and not ref["native_grad"]
    for k in ("d", "g", "l1", "lpips", "face_reg", "r1"):
        a, b = one["losses"][k], ref["losses"][k]
        assert abs(a - b) <= 0.05 * max(abs(b), 0.05), (k, a, b)


def test_two_devices_in_one_process_threads(cuda):
    """nn.DataParallel-style use (train_3_encoder.py:355-362): worker threads of ONE process drive different GPUs.
    Kernels that opt into > 48 KB of dynamic shared memory must opt in on every device."""
    import threading
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs in one process")
    import stylegan2
    torch.manual_seed(0)
    gen = stylegan2.Generator(64, 64, 2).eval()
    lat = torch.randn(4, gen.n_latent, 64)
    ext = torch.randn(4, 512, 4, 4)
    noise = [torch.randn(4, 1, 2 ** ((i + 5) // 2), 2 ** ((i + 5) // 2)) for i in range(gen.num_layers)]
    outs, errs = {}, []
    k = torch.tensor([1., 3., 3., 1.])
    k = torch.outer(k, k) / 64

    def work(d):
        try:
            dev = torch.device("cuda", d)
            with torch.cuda.device(dev), torch.no_grad():
                g = stylegan2.Generator(64, 64, 2).eval()
                g.load_state_dict(gen.state_dict())
                g = g.to(dev)
                import op
                big = torch.randn(2, 8, 257, 257, device=dev)             # upfirdn2d stream kernel (200 KB smem)
                outs[("blur", d)] = op.upfirdn2d(big, k.to(dev), pad=(2, 1)).cpu()
                outs[("blur_in", d)] = big.cpu()
                outs[d] = g(None, latent_styles=[lat.to(dev)], input_is_latent=True, noise=[n.to(dev) for n in noise],
                            use_external_input_tensor=True, external_input_tensor=ext.to(dev)).cpu()
        except Exception as e:      # noqa: BLE001
            errs.append((d, repr(e)))
    ths = [threading.Thread(target=work, args=(d,)) for d in (0, 1)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errs, errs
    assert torch.isfinite(outs[0]).all() and torch.allclose(outs[0], outs[1], rtol=1e-3, atol=1e-3)
    from oracle import fm_oracle as orc
    for d in (0, 1):
        ref = orc.upfirdn2d_api_ref(outs[("blur_in", d)], k, pad=(2, 1))
        assert torch.allclose(outs[("blur", d)], ref, atol=1e-5)
