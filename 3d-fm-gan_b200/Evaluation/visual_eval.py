"""Output stage of the reference's ``Evaluation/visual_eval.py`` on the device.

Only the conversion helpers are re-implemented here.  Everything else the reference module exports
(``Get_Real_Img_Val_Sample``, ``Get_Syn_Img_Val_Sample``, ``Get_Single_Eval_Result``, ``Get_Batch_Eval_Result`` ...,
imported by ``train_3_encoder.py:34``) comes from executing the shadowed reference file in this namespace
(``fm3d/_overlay.py``), so its loops call the mirrored funnel and this file's ``tensor2im``."""
from fm3d._overlay import load_shadowed

load_shadowed(globals())          # Get_*_Val_Sample, Get_Batch_Eval_Result, ... of the reference file; tensor2im below wins

import numpy as np  # noqa: E402
import torch  # noqa: E402,F401

from fm3d import ops  # noqa: E402


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
