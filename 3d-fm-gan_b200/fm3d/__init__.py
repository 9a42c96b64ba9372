"""fm3d: Python runtime of the B200-native StyleGAN2 synthesis path (ctypes over libfm3d.so)."""
from . import _lib, ops  # noqa: F401
