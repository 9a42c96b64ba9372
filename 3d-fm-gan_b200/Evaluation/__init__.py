"""Mirror package ``Evaluation``: replaces ``visual_eval.tensor2im``; every other module of the reference's ``Evaluation``
package stays importable through the extended ``__path__`` (fm3d/_overlay.py)."""
from fm3d._overlay import extend

__path__ = extend(__path__, __name__)
