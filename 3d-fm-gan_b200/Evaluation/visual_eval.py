"""Output stage of the reference's ``Evaluation/visual_eval.py`` on the device.

Only the conversion helpers live here -- the script part of the reference file (argument parsing,
checkpoint loading, figure layout) runs unchanged on top of the mirrored modules."""
import numpy as np
import torch

from fm3d import ops


def tensor2im_batch(image_tensor, cent=1., factor=255. / 2.):
    """[B,3,H,W] in [-1,1] -> uint8 device tensor [B,H,W,3]: one kernel for the batch, one D2H of 1/4 the bytes."""
    return ops.tensor2im_batch(image_tensor, cent, factor)


def tensor2im(image_tensor, imtype=np.uint8, cent=1., factor=255. / 2.):
    """Same signature and result as the reference (Evaluation/visual_eval.py:24-38): converts
    ``image_tensor[0]`` to a numpy HWC image in [0, 255]."""
    if not image_tensor.is_cuda:
        raise RuntimeError("tensor2im: the B200 path has no CPU fallback (got a CPU tensor)")
    out = ops.tensor2im_batch(image_tensor[:1], cent, factor)[0].cpu().numpy()
    return out if imtype == np.uint8 else out.astype(imtype)
