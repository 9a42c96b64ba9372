"""Native-level drop-in: the reference's two pybind modules re-created over ``libfm3d.so``.

The reference builds two extensions at import time, ``fused`` (``op/fused_act.py:20-26`` ->
``fused.fused_bias_act(input, bias, refer, act, grad, alpha, scale)``, ``op/fused_bias_act.cpp:11-17``) and
``upfirdn2d`` (``op/upfirdn2d.py:19-25`` -> ``upfirdn2d_op.upfirdn2d(input, kernel, up_x, up_y, down_x, down_y,
pad_x0, pad_x1, pad_y0, pad_y1)``, ``op/upfirdn2d.cpp:12-19``).  ``fused`` and ``upfirdn2d_op`` below have those
call signatures and semantics (empty tensor = absent argument, outputs allocated inside, current-stream launch,
``RuntimeError`` on failure) and call ``fm_bias_act`` / ``fm_upfirdn2d`` through the C ABI.

``install()`` makes ``torch.utils.cpp_extension.load(name, sources, ...)`` return them for exactly those two
module names, so the reference's *unmodified* ``op/*.py`` (and its autograd classes) run on the sm_100a
kernels:

    import fm3d.pybind_compat as pc; pc.install()
    import op                     # the reference's own package; no nvcc/ninja JIT happens

CUDA tensors only (the reference's Python wrappers keep their own CPU branch and never reach these objects
for CPU tensors, op/fused_act.py:114, op/upfirdn2d.py:155).
"""
import torch

from . import ops


class fused:                      # replaces the pybind module of op/fused_bias_act.cpp
    @staticmethod
    def fused_bias_act(input, bias, refer, act, grad, alpha, scale):
        return ops.bias_act(input, bias if bias.numel() else None, refer if refer.numel() else None,
                            act, grad, alpha, scale)


class upfirdn2d_op:               # replaces the pybind module of op/upfirdn2d.cpp
    @staticmethod
    def upfirdn2d(input, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1):
        if input.ndim != 4:
            raise RuntimeError("upfirdn2d: input must be [major, in_h, in_w, minor]")
        major, in_h, in_w, minor = input.shape
        if minor != 1:
            # the Python wrapper always passes minor == 1 (op/upfirdn2d.py:108); keep the general layout correct
            x = input.permute(0, 3, 1, 2).reshape(major * minor, in_h, in_w)
            out = ops.upfirdn2d_planes(x, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1)
            return out.view(major, minor, out.shape[-2], out.shape[-1]).permute(0, 2, 3, 1).contiguous()
        out = ops.upfirdn2d_planes(input.reshape(major, in_h, in_w), kernel, up_x, up_y, down_x, down_y,
                                   pad_x0, pad_x1, pad_y0, pad_y1)
        return out.unsqueeze(-1)


_MODULES = {"fused": fused, "upfirdn2d": upfirdn2d_op}
_ORIG_LOAD = None


def install():
    """Route ``cpp_extension.load("fused" | "upfirdn2d", ...)`` to the objects above; other names JIT as before."""
    global _ORIG_LOAD
    from torch.utils import cpp_extension
    if _ORIG_LOAD is not None:
        return
    _ORIG_LOAD = cpp_extension.load

    def load(name, *args, **kwargs):
        mod = _MODULES.get(name)
        if mod is not None:
            return mod
        return _ORIG_LOAD(name, *args, **kwargs)
    cpp_extension.load = load


def uninstall():
    global _ORIG_LOAD
    from torch.utils import cpp_extension
    if _ORIG_LOAD is not None:
        cpp_extension.load = _ORIG_LOAD
        _ORIG_LOAD = None
