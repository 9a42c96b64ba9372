"""Mirror of the reference's ``dataset.py``: every dataset / sampler class comes from the reference file this module
shadows (executed in this namespace, ``fm3d/_overlay.py``); ``Data_Loading`` (dataset.py:361-413) is replaced by the
device-side version of ``fm3d.data`` -- same arguments and return values, no ``.cpu().numpy()`` round trip, no blocking copies
when it is fed by ``fm3d.data.DevicePrefetcher``."""
from fm3d._overlay import load_shadowed

load_shadowed(globals())

from fm3d.data import Data_Loading, DevicePrefetcher, u8_transform  # noqa: E402,F401
