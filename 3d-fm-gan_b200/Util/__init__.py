"""Mirror package ``Util``: replaces ``network_util.Forward_Inference_3_Encoder``; every other module of the reference's ``Util``
package stays importable through the extended ``__path__`` (fm3d/_overlay.py)."""
from fm3d._overlay import extend

__path__ = extend(__path__, __name__)
