"""Overlay loading for the mirror packages (the drop-in boundary of SURVEY 8b).

``3d-fm-gan_b200/`` sits in front of a checkout of the reference on ``sys.path``.  A mirror
package (``Util``, ``Evaluation``, ``psp_encoder_model`` ...) must not hide the modules of the
reference package of the same name that it does not replace (``Util.training_util``,
``Evaluation.quant_eval``, ``Evaluation.fid`` ...), and a mirror *module* that replaces one function
(``Util.network_util.Forward_Inference_3_Encoder``, ``Evaluation.visual_eval.tensor2im``) must still
export every other name the reference's callers import from it (``Forward_Inference``,
``Build_Generator_From_Dict``, ``Get_Real_Img_Val_Sample`` ... -- ``train_3_encoder.py:26-37``,
``Evaluation/visual_eval.py:16``).  Two helpers do that:

* ``extend(path, name)``      -- for a package ``__init__``: append the same-named directories found
                                 further down ``sys.path`` to the package's ``__path__``.
* ``load_shadowed(globals())`` -- for a module: find the module file this one shadows on the package's
                                 extended ``__path__`` and execute it in this module's namespace, so all of
                                 its names exist here; whatever the mirror module defines afterwards wins.

Nothing is copied: the reference's source is executed from wherever the user's checkout lives.  Without a
reference checkout on ``sys.path`` both helpers are no-ops and the mirror stands alone.
"""
import os
import pkgutil
import sys


def extend(path, name):
    """``__path__ = extend(__path__, __name__)`` in a mirror package's ``__init__``."""
    return pkgutil.extend_path(path, name)


def shadowed_file(module_name, own_file):
    """Path of the first ``<module>.py`` that ``module_name`` shadows, or None."""
    pkg_name, _, leaf = module_name.rpartition(".")
    if pkg_name:
        pkg = sys.modules.get(pkg_name)
        search = list(getattr(pkg, "__path__", []))
    else:
        search = list(sys.path)
    own = os.path.realpath(own_file)
    for d in search:
        cand = os.path.join(d or ".", leaf + ".py")
        if os.path.isfile(cand) and os.path.realpath(cand) != own:
            return cand
    return None


def load_shadowed(module_globals):
    """Execute the shadowed module's source in ``module_globals`` (call it first thing in the mirror module).
    Returns the path that was loaded, or None when this module shadows nothing."""
    name, own = module_globals["__name__"], module_globals["__file__"]
    path = shadowed_file(name, own)
    if path is None:
        return None
    with open(path, "rb") as f:
        source = f.read()
    code = compile(source, path, "exec")       # tracebacks point at the reference file
    exec(code, module_globals)
    module_globals["__shadowed_file__"] = path
    return path
