"""Harness for running the reference's *callers* unchanged on top of the mirror (SURVEY 8b "harness caveats").

Test / tool infrastructure only -- nothing under ``3d-fm-gan_b200/`` imports this.

* ``reference_root()``   the reference checkout: ``/root/reference`` in the build container, or the staged copy
                         ``baseline/_ref`` (git-ignored, made by ``tools/stage_reference.sh``) on a GPU box.
* ``install_shims()``    stub modules for the third-party packages the reference's scripts import at module level
                         and this image lacks (matplotlib, easydict, imageio, face_alignment, skimage.measure,
                         IPython): train_3_encoder.py:15-16,39-41, Util/landmark_util.py:15-17,
                         lpips/__init__.py:16, lpips/base_model.py:14, lpips/dist_model.py:25; plus the typing
                         names torch 1.9 re-exported from torch.utils.data.sampler (dataset.py:17).
* ``activate()``         sys.path order of INTEGRATION.md section 1: mirror first, reference second.
* ``load_train_script()`` the valid prefix of train_3_encoder.py (lines 1-879; a CRLF tail fragment follows,
                         SURVEY 0.6) executed as a module.
"""
import importlib.machinery
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "3d-fm-gan_b200")
TRAIN_SCRIPT_VALID_LINES = 879


def reference_root():
    for cand in (os.environ.get("FM3D_REFERENCE"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "stylegan2.py")):
            return cand
    return None


def _stub(name, **attrs):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, None)
        m.__fm3d_stub__ = True
        sys.modules[name] = m
        parent, _, leaf = name.rpartition(".")
        if parent:
            setattr(_stub(parent), leaf, m)
            if not hasattr(sys.modules[parent], "__path__"):
                sys.modules[parent].__path__ = []
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def _absent(name):
    try:
        __import__(name)
        return False
    except Exception:
        return True


def install_shims():
    """Idempotent.  Only packages that are really absent are stubbed."""
    def _unavailable(what):
        def f(*a, **k):
            raise RuntimeError(f"{what} is a stub in this environment (package not installed)")
        return f

    if _absent("matplotlib"):
        _stub("matplotlib")
        _stub("matplotlib.pyplot", figure=_unavailable("matplotlib.pyplot.figure"))
    if _absent("easydict"):
        class EasyDict(dict):                       # attribute access over a dict, all the reference uses
            def __init__(self, d=None, **kw):
                super().__init__()
                for k, v in dict(d or {}, **kw).items():
                    self[k] = v

            def __getattr__(self, k):
                try:
                    return self[k]
                except KeyError:
                    raise AttributeError(k)

            def __setattr__(self, k, v):
                self[k] = v
        _stub("easydict", EasyDict=EasyDict)
    if _absent("imageio"):
        _stub("imageio", mimsave=_unavailable("imageio.mimsave"), imread=_unavailable("imageio.imread"))
    if _absent("face_alignment"):
        class _LandmarksType:
            _2D, _2halfD, _3D = 1, 2, 3
        _stub("face_alignment", FaceAlignment=_unavailable("face_alignment.FaceAlignment"), LandmarksType=_LandmarksType)
        _stub("face_alignment.detection")
        _stub("face_alignment.detection.sfd")
        _stub("face_alignment.detection.sfd.detect", get_predictions=_unavailable("face_alignment get_predictions"))
        _stub("face_alignment.utils", transform=_unavailable("face_alignment.utils.transform"),
              get_preds_fromhm=_unavailable("face_alignment.utils.get_preds_fromhm"))
    if _absent("skimage"):
        _stub("skimage")
        _stub("skimage.measure", compare_ssim=_unavailable("skimage.measure.compare_ssim"))
        _stub("skimage.transform")                  # lpips/dist_model.py:25
        _stub("skimage.color")                      # lpips/networks_basic.py:20
    else:
        import skimage.measure as sm                # modern scikit-image dropped compare_ssim (lpips/__init__.py:16)
        if not hasattr(sm, "compare_ssim"):
            sm.compare_ssim = _unavailable("skimage.measure.compare_ssim")
    if _absent("IPython"):
        _stub("IPython", embed=_unavailable("IPython.embed"))
    # lpips/pretrained_networks.py builds its backbone with torchvision.models.vgg16(pretrained=True), a download: offline the
    # architecture is kept and the weights stay random-init (the LPIPS linear heads are in the tree: lpips/weights/v0.1)
    try:
        import torchvision
        if not getattr(torchvision.models.vgg16, "__fm3d_offline__", False):
            _orig_vgg16 = torchvision.models.vgg16

            def vgg16(pretrained=False, **kw):
                kw.pop("weights", None)
                return _orig_vgg16(weights=None, **kw)
            vgg16.__fm3d_offline__ = True
            torchvision.models.vgg16 = vgg16
    except Exception:
        pass
    # Evaluation/inception.py downloads the FID Inception weights (torch.utils.model_zoo.load_url) into a freshly built
    # torchvision inception_v3: offline the download is replaced by the state dict of that very instance (random-init)
    try:
        import torch.utils.model_zoo as _zoo
        import torchvision
        if not getattr(_zoo.load_url, "__fm3d_offline__", False):
            _orig_inc, _orig_load, _last = torchvision.models.inception_v3, _zoo.load_url, {}

            def inception_v3(*a, **k):
                k.pop("pretrained", None)
                k.setdefault("weights", None)
                k.setdefault("init_weights", False)
                _last["m"] = _orig_inc(*a, **k)
                return _last["m"]

            def load_url(url, *a, **k):
                if "pt_inception" in str(url) and "m" in _last:
                    return _last["m"].state_dict()
                return _orig_load(url, *a, **k)
            load_url.__fm3d_offline__ = True
            torchvision.models.inception_v3 = inception_v3
            _zoo.load_url = load_url
    except Exception:
        pass
    # torch 1.9 re-exported typing names from torch.utils.data.sampler; dataset.py:17 imports them from there
    import typing
    import torch.utils.data.sampler as _sampler
    for _n in ("Sized", "Optional", "Iterator"):
        if not hasattr(_sampler, _n):
            setattr(_sampler, _n, getattr(typing, _n))


def activate(ref=None):
    """Mirror first, reference second.  Returns the reference root."""
    ref = ref or reference_root()
    if ref is None:
        raise RuntimeError("no reference checkout: /root/reference or baseline/_ref (tools/stage_reference.sh)")
    for p in (ref, PKG):
        while p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, ref)
    sys.path.insert(0, PKG)
    install_shims()
    return ref


def load_train_script(ref=None, name="train_3_encoder"):
    """Execute lines 1-879 of the reference's train_3_encoder.py as module ``name`` (its ``__main__`` block does not run)."""
    ref = activate(ref)
    path = os.path.join(ref, "train_3_encoder.py")
    with open(path, "rb") as f:
        lines = f.read().splitlines(keepends=True)
    src = b"".join(lines[:TRAIN_SCRIPT_VALID_LINES])
    mod = types.ModuleType(name)
    mod.__file__ = path
    sys.modules[name] = mod
    exec(compile(src, path, "exec"), mod.__dict__)
    return mod
