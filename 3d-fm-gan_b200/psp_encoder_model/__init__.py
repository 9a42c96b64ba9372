"""Mirror package ``psp_encoder_model``: replaces ``encoders.helpers`` / ``encoders.psp_encoders``; every other module of the reference's ``psp_encoder_model``
package stays importable through the extended ``__path__`` (fm3d/_overlay.py)."""
from fm3d._overlay import extend

__path__ = extend(__path__, __name__)
