"""Mirror package ``psp_encoder_model.encoders``: replaces ``helpers`` and ``psp_encoders``; every other module of the reference's ``psp_encoder_model.encoders``
package stays importable through the extended ``__path__`` (fm3d/_overlay.py)."""
from fm3d._overlay import extend

__path__ = extend(__path__, __name__)
