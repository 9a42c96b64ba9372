"""The C-ABI library loads and exports every symbol include/fm3d.h declares (no GPU needed)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "fm3d.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fm_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    from fm3d import _lib
    syms = _declared_symbols()
    assert len(syms) >= 10
    handle = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in syms if not hasattr(handle, s)]
    assert not missing, f"libfm3d.so does not export {missing}"
    # the Python binding covers exactly the declared ABI
    assert sorted(_lib.EXPORTED_SYMBOLS) == syms


def test_version_and_error_text():
    from fm3d import _lib
    lib = _lib.lib()
    assert lib.fm_version() >= 100
    # argument validation happens before any CUDA call: usable without a GPU
    st = lib.fm_bias_act(None, None, None, None, 1, 1, 1, 3, 0, 0.2, 1.0, 0, None)
    assert st == 1
    assert b"null" in lib.fm_last_error()
    st = lib.fm_upfirdn2d(None, None, None, 1, 4, 4, 0, 0, 1, 1, 1, 1, 0, 0, 0, 0, 0, None)
    assert st == 1 and b"kernel" in lib.fm_last_error()


def test_conv_desc_layout_matches_header():
    """ctypes.Structure must mirror fm_conv_desc: compare against the C compiler's sizeof."""
    import subprocess
    import tempfile
    from fm3d import _lib
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "fm3d.h"\nint main(){printf("%zu %zu %zu %zu %zu",' \
          'sizeof(fm_conv_desc),offsetof(fm_conv_desc,stride),offsetof(fm_conv_desc,tab),offsetof(fm_conv_desc,block_n),' \
          'sizeof(fm_table_layer));return 0;}'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        out = subprocess.check_output([os.path.join(d, "t")]).decode().split()
    size, off_stride, off_tab, off_bn, size_tl = map(int, out)
    assert ctypes.sizeof(_lib.ConvDesc) == size
    assert _lib.ConvDesc.stride.offset == off_stride
    assert _lib.ConvDesc.tab.offset == off_tab
    assert _lib.ConvDesc.block_n.offset == off_bn
    assert ctypes.sizeof(_lib.TableLayer) == size_tl


def test_wgrad_desc_layout_matches_header():
    """fm_wgrad_desc vs its ctypes mirror, and argument validation of the training entry point without a GPU."""
    import subprocess
    import tempfile
    from fm3d import _lib
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "fm3d.h"\nint main(){printf("%zu %zu %zu %zu %zu",' \
          'sizeof(fm_wgrad_desc),offsetof(fm_wgrad_desc,GW),offsetof(fm_wgrad_desc,tap_dy_b),offsetof(fm_wgrad_desc,dw),' \
          'offsetof(fm_wgrad_desc,ksplit));return 0;}'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        out = subprocess.check_output([os.path.join(d, "t")]).decode().split()
    size, off_l, off_tb, off_dw, off_ks = map(int, out)
    assert ctypes.sizeof(_lib.WgradDesc) == size
    assert _lib.WgradDesc.GW.offset == off_l and _lib.WgradDesc.tap_dy_b.offset == off_tb
    assert _lib.WgradDesc.dw.offset == off_dw and _lib.WgradDesc.ksplit.offset == off_ks
    lib = _lib.lib()
    assert lib.fm_conv_wgrad(None, None) == 1 and b"null" in lib.fm_last_error()
