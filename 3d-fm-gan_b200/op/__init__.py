"""Drop-in replacement for the reference's ``op`` package (op/__init__.py:10-11).

Same public names -- ``FusedLeakyReLU``, ``fused_leaky_relu``, ``upfirdn2d`` -- backed by the
hand-written sm_100a kernels in libfm3d.so instead of the JIT-compiled 2019 StyleGAN2
extensions.  CUDA tensors only: there is deliberately no CPU branch.
"""
from .fused_act import FusedLeakyReLU, fused_leaky_relu  # noqa: F401
from .upfirdn2d import upfirdn2d  # noqa: F401
