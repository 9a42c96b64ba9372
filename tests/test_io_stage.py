"""the input stage (device-side ToTensor + Normalize, prefetcher, Data_Loading) and the
asynchronous checkpoint writer.  Host logic runs on CPU; the kernel and the stream-ordered paths under ``-m gpu``."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ref_env  # noqa: E402


def _loader(batches):
    for b in batches:
        yield b


def test_data_loading_matches_reference_semantics_cpu():
    """The three modes of dataset.py:361-413 against the reference function itself (CPU tensors)."""
    if ref_env.reference_root() is None:
        pytest.skip("no reference checkout here")
    import importlib.util
    ref_env.install_shims()
    spec = importlib.util.spec_from_file_location("ref_dataset", os.path.join(ref_env.reference_root(), "dataset.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    from fm3d import data
    g = torch.Generator().manual_seed(0)
    batches = [(torch.randn(8, 3, 4, 4, generator=g), torch.randn(8, 3, 4, 4, generator=g)) for _ in range(3)]
    for ds_flag, ex in ((False, False), (True, False), (True, True)):
        a = ref.Data_Loading(_loader(batches), _loader(batches), ds_flag, "cpu", extreme_loader=_loader(batches), extreme_ds_flag=ex)
        b = data.Data_Loading(_loader(batches), _loader(batches), ds_flag, "cpu", extreme_loader=_loader(batches), extreme_ds_flag=ex)
        assert len(a) == len(b) == 3
        for x, y in zip(a, b):
            assert torch.equal(x, y)


def test_async_checkpoint_cpu(tmp_path):
    from fm3d.checkpoint import AsyncCheckpointWriter
    sd = {"g": {"w": torch.randn(4, 5), "b": torch.zeros(3)}, "opt": {"state": {0: {"step": 3, "m": torch.ones(2)}}, "lr": 1e-3},
          "tsr_encode": "Render Image", "sliced_layer": None}
    w = AsyncCheckpointWriter()
    path = str(tmp_path / "000010.pt")
    w.save(sd, path)
    w.wait()
    back = torch.load(path)
    assert torch.equal(back["g"]["w"], sd["g"]["w"]) and back["opt"]["state"][0]["step"] == 3 and back["tsr_encode"] == "Render Image"
    w.save(sd, str(tmp_path / "missing_dir" / "x.pt"))
    with pytest.raises(Exception):
        w.wait()


@pytest.mark.gpu
def test_im2tensor_bit_exact(cuda):
    """uint8 NHWC -> normalised fp32 NCHW on the device == torchvision ToTensor() + Normalize((0.5,)*3, (0.5,)*3)."""
    from torchvision import transforms
    from fm3d import ops
    rng = np.random.default_rng(0)
    imgs = rng.integers(0, 256, size=(5, 64, 48, 3), dtype=np.uint8)
    imgs[0] = np.arange(64 * 48 * 3, dtype=np.uint32).reshape(64, 48, 3) % 256        # every byte value
    tf = transforms.Compose([transforms.ToTensor(), transforms.Normalize((0.5, 0.5, 0.5), (0.5, 0.5, 0.5), inplace=True)])
    ref = torch.stack([tf(img) for img in imgs])
    got = ops.im2tensor_batch(torch.from_numpy(imgs).to(cuda))
    assert got.shape == (5, 3, 64, 48) and torch.equal(got.cpu(), ref)
    # round trip through the output stage: the reference's tensor2im TRUNCATES (astype(uint8), visual_eval.py:38), so a byte
    # comes back as itself or one less, never more
    back = ops.tensor2im_batch(got).cpu().to(torch.int16) - torch.from_numpy(imgs).to(torch.int16)
    assert int(back.max()) == 0 and int(back.min()) >= -1


@pytest.mark.gpu
def test_prefetcher_and_device_data_loading(cuda):
    from fm3d import data
    rng = np.random.default_rng(1)
    batches = [(torch.from_numpy(rng.integers(0, 256, size=(6, 32, 32, 3), dtype=np.uint8)),
                torch.randn(6, 3, 32, 32)) for _ in range(4)]
    pf = data.DevicePrefetcher(_loader(batches), cuda)
    seen = 0
    for (a, b), (ua, fb) in zip(pf, batches):
        assert a.is_cuda and a.dtype == torch.float32 and a.shape == (6, 3, 32, 32)
        ref = (ua.permute(0, 3, 1, 2).float().div(255) - 0.5) / 0.5
        assert torch.equal(a.cpu(), ref) and torch.equal(b.cpu(), fb)
        seen += 1
    assert seen == 4
    pf = data.DevicePrefetcher(_loader(batches), cuda)
    g_in, r_in, g_ref = data.Data_Loading(pf, pf, True, cuda)
    swap = [1, 0, 3, 2, 5, 4]
    assert g_in.is_cuda and torch.equal(g_ref, g_in[swap]) and torch.equal(r_in.cpu(), batches[0][1][swap])


@pytest.mark.gpu
def test_async_checkpoint_gpu(cuda, tmp_path):
    """The snapshot is ordered after queued work and is not disturbed by later updates of the parameters."""
    from fm3d.checkpoint import AsyncCheckpointWriter
    p = torch.zeros(1 << 22, device=cuda)
    p.add_(1.0)                                         # queued before the snapshot: must be in it
    w = AsyncCheckpointWriter(cuda)
    path = str(tmp_path / "ckpt.pt")
    w.save({"p": p, "n": 7}, path)
    p.add_(1.0)                                         # after the snapshot: must not be in it
    w.wait()
    back = torch.load(path)
    assert back["n"] == 7 and float(back["p"].min()) == 1.0 and float(back["p"].max()) == 1.0
    w.save({"p": p}, path)                              # pinned buffers are reused
    w.wait()
    assert float(torch.load(path)["p"].max()) == 2.0
