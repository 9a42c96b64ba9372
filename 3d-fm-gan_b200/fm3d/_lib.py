"""ctypes binding of libfm3d.so (the C ABI declared in include/fm3d.h).

The library is built in-tree by ``csrc/build.sh`` (``__graft_entry__.build()``).  There is
no fallback: if the shared object is missing or a call fails, a RuntimeError is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FM3D_LIB") or os.path.join(_HERE, "libfm3d.so")   # FM3D_LIB: A/B-test another build

FM_F32, FM_F16, FM_BF16 = 0, 1, 2
FM_MAX_TAPS = 49


class ConvDesc(C.Structure):
    """fm_conv_desc (include/fm3d.h)."""
    _fields_ = [
        ("x", C.c_void_p), ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("Cin", C.c_int32),
        ("x_cstride", C.c_int32),
        ("w", C.c_void_p), ("ntaps", C.c_int32), ("Cout", C.c_int32), ("w_rows", C.c_int32), ("w_cstride", C.c_int32),
        ("tap_dy", C.c_int8 * FM_MAX_TAPS), ("tap_dx", C.c_int8 * FM_MAX_TAPS), ("tap_widx", C.c_int8 * FM_MAX_TAPS),
        ("stride", C.c_int32), ("stride_x", C.c_int32), ("stride_y", C.c_int32),
        ("x_pixstride", C.c_int64), ("x_rowstride", C.c_int64), ("x_imgstride", C.c_int64),
        ("groups", C.c_int32),
        ("OH", C.c_int32), ("OW", C.c_int32),
        ("out", C.c_void_p),
        ("out_H", C.c_int32), ("out_W", C.c_int32), ("out_cstride", C.c_int32), ("out_y0", C.c_int32),
        ("out_x0", C.c_int32), ("out_ys", C.c_int32), ("out_xs", C.c_int32), ("out_nchw_f32", C.c_int32),
        ("tab", C.c_void_p), ("tab_bstride", C.c_int32),
        ("noise", C.c_void_p), ("noise_bstride", C.c_int32), ("noise_w", C.c_void_p),
        ("residual", C.c_void_p), ("rgb", C.c_void_p),
        ("border_tab", C.c_void_p), ("out_cgroup", C.c_int32), ("out_gstride", C.c_int64),
        ("splitk_ws", C.c_void_p), ("splitk_ws_bytes", C.c_int64), ("ksplit", C.c_int32), ("upmode", C.c_int32),
        ("block_n", C.c_int32), ("tile_w", C.c_int32), ("tile_h", C.c_int32), ("out_cgroup_ow_shrink", C.c_int32),
        ("residual_up_h", C.c_int32), ("residual_up_w", C.c_int32),
        ("out_pitch_h", C.c_int32), ("out_pitch_w", C.c_int32),
        ("nphases", C.c_int32), ("phase_ntaps", C.c_int32 * 4), ("phase_out_y0", C.c_int32 * 4), ("phase_out_x0", C.c_int32 * 4),
        ("max_ctas", C.c_int32), ("tile_counter", C.c_void_p), ("colsum", C.c_void_p),
    ]


class WgradOperand(C.Structure):
    """fm_wgrad_operand (include/fm3d.h)."""
    _fields_ = [("ptr", C.c_void_p), ("C", C.c_int32), ("cstride", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("stride", C.c_int32)]


class WgradDesc(C.Structure):
    """fm_wgrad_desc (include/fm3d.h)."""
    _fields_ = [
        ("a", WgradOperand), ("b", WgradOperand),
        ("B", C.c_int32), ("GH", C.c_int32), ("GW", C.c_int32), ("ntaps", C.c_int32),
        ("tap_dy_a", C.c_int8 * FM_MAX_TAPS), ("tap_dx_a", C.c_int8 * FM_MAX_TAPS),
        ("tap_dy_b", C.c_int8 * FM_MAX_TAPS), ("tap_dx_b", C.c_int8 * FM_MAX_TAPS),
        ("dw", C.c_void_p), ("dw_tap_stride", C.c_int64), ("dw_row_stride", C.c_int32), ("ksplit", C.c_int32),
    ]


class StyleLayer(C.Structure):
    """fm_style_layer."""
    _fields_ = [("wmod", C.c_void_p), ("bmod", C.c_void_p), ("s", C.c_void_p),
                ("cin", C.c_int32), ("latent_idx", C.c_int32)]


class TableLayer(C.Structure):
    """fm_table_layer."""
    _fields_ = [("s", C.c_void_p), ("wsq", C.c_void_p), ("act_bias", C.c_void_p), ("s_next", C.c_void_p),
                ("wrgb", C.c_void_p), ("s_rgb", C.c_void_p), ("tab", C.c_void_p),
                ("cin", C.c_int32), ("cout", C.c_int32), ("slope", C.c_float), ("gain", C.c_float)]


_SIGNATURES = {
    "fm_version": (C.c_int, []),
    "fm_last_error": (C.c_char_p, []),
    "fm_launch_count": (C.c_int64, []),
    "fm_add_launches": (None, [C.c_int64]),
    "fm_bias_act": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                              C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_void_p]),
    "fm_channel_scale": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p]),
    "fm_plane_add": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_void_p]),
    "fm_channel_dot": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p]),
    "fm_bias_act_grad_bias": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                        C.c_int64, C.c_int, C.c_float, C.c_float, C.c_int, C.c_void_p]),
    "fm_upfirdn2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64] + [C.c_int] * 12 + [C.c_int, C.c_void_p]),
    "fm_conv_igemm": (C.c_int, [C.POINTER(ConvDesc), C.c_void_p]),
    "fm_igemm_trace": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "fm_conv_wgrad": (C.c_int, [C.POINTER(WgradDesc), C.c_void_p]),
    "fm_style_affine": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "fm_build_tables": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "fm_tensor2im_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p]),
    "fm_im2tensor_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p]),
    "fm_nchw_to_nhwc_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 5 + [C.c_void_p]),
    "fm_nhwc_bf16_to_nchw": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 5 + [C.c_void_p]),
    "fm_blur_act_nhwc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
                         + [C.c_int] * 8 + [C.c_void_p]),
    "fm_rgb_finalize": (C.c_int, [C.c_void_p] * 5 + [C.c_int] * 3 + [C.c_void_p]),
    "fm_prep_weight": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                 C.c_int, C.c_int, C.c_void_p]),
    "fm_image_to_nhwc8_padded": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 8 + [C.c_void_p]),
    "fm_maxpool3x3s2_nhwc": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 4 + [C.c_void_p]),
    "fm_avgpool_nhwc_to_nchw": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 7 + [C.c_void_p]),
    "fm_channel_sum_nhwc": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 4 + [C.c_void_p]),
    "fm_se_gate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p] + [C.c_int] * 3 + [C.c_void_p]),
    "fm_se_combine_nhwc": (C.c_int, [C.c_void_p] * 4 + [C.c_int] * 9 + [C.c_void_p]),
    "fm_bilinear_up_nhwc": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 6 + [C.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES.keys())

_lib = None


def lib():
    """Load libfm3d.so (once).  Fails loudly: the product has no CPU / PyTorch fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"libfm3d.so not found at {LIB_PATH}: build it with 3d-fm-gan_b200/csrc/build.sh "
                "(or __graft_entry__.build()); there is no fallback path")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(status, what):
    if status != 0:
        msg = lib().fm_last_error()
        raise RuntimeError(f"{what} failed (fm_status {status}): {msg.decode() if msg else ''}")


def launch_count():
    return int(lib().fm_launch_count())
