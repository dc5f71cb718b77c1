"""ctypes binding of libsnb_b200.so (include/snb_b200.h).

This is the only place the package touches native code.  There is no CPU fallback: if the library is missing
the import of anything that computes fails loudly, and every entry point takes DEVICE pointers only.
Error codes are mapped onto the exception types the reference raises for the same mistakes
(`ValueError` lib/tiles.py:56-57,81-85,138-139; `AssertionError` lib/tiles.py:99-100; `RuntimeError` for CUDA
failures, mirroring `_check` in lib/modules/abn/functions.py:12-15).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsnb_b200.so")

SNB_OK = 0
SNB_E_INVALID = -1
SNB_E_SHAPE = -2
SNB_E_CUDA = -3
SNB_E_UNSUPPORTED = -4

DT_U8, DT_F32, DT_F64, DT_I64 = 0, 1, 2, 3
LAYOUT_NCHW_F32, LAYOUT_PATCH32, LAYOUT_PATCH32_F32, LAYOUT_NHWC3_BF16 = 0, 1, 2, 3
CONV_3X3, CONV_1X1, CONVT_4X4_S2, CONVT_3X3_S2, CONV_2X2, CONVT_3X3_S2_FULL, CONV_2X2_ADJ, CONV_FIRST_3X3 = 0, 1, 2, 3, 4, 5, 6, 7
CONV_BF16, CONV_TF32 = 0, 1

c_i64 = ctypes.c_int64
c_vp = ctypes.c_void_p
c_int = ctypes.c_int


class ConvDesc(ctypes.Structure):
    """struct snb_conv_desc (include/snb_b200.h)."""

    _fields_ = [
        ("kind", ctypes.c_int32),
        ("relu", ctypes.c_int32),
        ("n", c_i64),
        ("h", c_i64),
        ("w", c_i64),
        ("cin", c_i64),
        ("in_cstride", c_i64),
        ("cout", c_i64),
        ("out_cstride", c_i64),
        ("d_in", c_vp),
        ("d_out", c_vp),
        ("d_weight", c_vp),
        ("d_bias", c_vp),
        ("d_head_w", c_vp),
        ("head_b", ctypes.c_float),
        ("head_sigmoid", ctypes.c_int32),
        ("d_head_out", c_vp),
        ("d_pool_out", c_vp),
        ("pool_cstride", c_i64),
        ("d_pre_scale", c_vp),
        ("d_pre_shift", c_vp),
        ("act_slope", ctypes.c_float),
        ("res_after_act", ctypes.c_int32),
        ("d_residual", c_vp),
        ("res_cstride", c_i64),
        ("valid", ctypes.c_int32),
        ("out_upsample2x", ctypes.c_int32),
        ("dtype", ctypes.c_int32),
    ]


class WgradDesc(ctypes.Structure):
    """struct snb_wgrad_desc (include/snb_b200.h)."""

    _fields_ = [("kind", ctypes.c_int32), ("valid", ctypes.c_int32), ("n", c_i64), ("h", c_i64), ("w", c_i64),
                ("cin", c_i64), ("in_cstride", c_i64), ("cout", c_i64), ("dout_cstride", c_i64), ("d_in", c_vp),
                ("d_dout", c_vp), ("d_dweight", c_vp), ("dw_cout", c_i64), ("dw_cin", c_i64)]


# name -> (restype, argtypes); every symbol include/snb_b200.h declares
class ConvGeom(ctypes.Structure):
    """struct snb_conv_geom (include/snb_b200.h)."""
    _fields_ = [(k, c_i64) for k in ("n", "big_h", "big_w", "big_c", "big_cstride", "small_h", "small_w", "small_c",
                                     "small_cstride", "kh", "kw", "stride", "pad")]


SIGNATURES = {
    "snb_version": (c_int, []),
    "snb_last_error": (ctypes.c_char_p, []),
    "snb_device_sm_count": (c_int, []),
    "snb_slicer_create": (c_int, [c_i64, c_i64, c_i64, c_i64, c_i64, ctypes.POINTER(c_vp)]),
    "snb_slicer_destroy": (None, [c_vp]),
    "snb_slicer_info": (c_int, [c_vp, ctypes.POINTER(c_i64)]),
    "snb_slicer_crops": (c_int, [c_vp, ctypes.POINTER(c_i64)]),
    "snb_split_hwc": (c_int, [c_vp, c_vp, c_i64, c_i64, c_int, c_vp, c_vp, c_i64, c_i64, c_vp]),
    "snb_split_norm_u8": (c_int, [c_vp, c_vp, c_i64, c_vp, c_int, c_int, c_vp, c_i64, c_i64, c_vp]),
    "snb_nchw_f32_to_nhwc3": (c_int, [c_vp, c_i64, c_i64, c_i64, c_vp, c_vp]),
    "snb_nchw_f32_to_patch32": (c_int, [c_vp, c_i64, c_i64, c_i64, c_i64, c_vp, c_int, c_vp]),
    "snb_merge": (c_int, [c_vp, c_vp, c_int, c_i64, c_int, c_vp, c_vp, c_int, c_vp, ctypes.c_float, c_vp]),
    "snb_merge_rows": (c_int, [c_vp, c_vp, c_int, c_i64, c_int, c_vp, c_vp, c_int, c_vp, ctypes.c_float, c_i64, c_i64, c_vp]),
    "snb_conv_create": (c_int, [ctypes.POINTER(ConvDesc), ctypes.POINTER(c_vp)]),
    "snb_conv_launch": (c_int, [c_vp, c_vp]),
    "snb_conv_destroy": (None, [c_vp]),
    "snb_conv_flops": (ctypes.c_double, [c_vp]),
    "snb_conv_set_head_out": (c_int, [c_vp, c_vp]),
    "snb_wgrad_create": (c_int, [ctypes.POINTER(WgradDesc), ctypes.POINTER(c_vp)]),
    "snb_wgrad_launch": (c_int, [c_vp, c_vp]),
    "snb_wgrad_destroy": (None, [c_vp]),
    "snb_wgrad_flops": (ctypes.c_double, [c_vp]),
    "snb_conv_scatter_create": (c_int, [c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64,
                                        ctypes.POINTER(c_vp)]),
    "snb_conv_scatter_launch": (c_int, [c_vp, c_vp]),
    "snb_conv_scatter_destroy": (None, [c_vp]),
    "snb_conv_scatter_flops": (ctypes.c_double, [c_vp]),
    "snb_conv_generic_fwd": (c_int, [ctypes.POINTER(ConvGeom), c_vp, c_vp, c_vp, c_vp, c_vp]),
    "snb_conv_generic_dgrad": (c_int, [ctypes.POINTER(ConvGeom), c_vp, c_vp, c_vp, c_vp, c_vp]),
    "snb_conv_generic_wgrad": (c_int, [ctypes.POINTER(ConvGeom), c_vp, c_vp, c_vp, c_vp]),
    "snb_maxpool2x2": (c_int, [c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_vp, c_i64, c_int, c_vp]),
    "snb_space_to_depth2": (c_int, [c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_vp, c_i64, c_vp]),
    "snb_depth_to_space2": (c_int, [c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_vp, c_i64, c_int, c_vp]),
    "snb_scale_nc_nhwc": (c_int, [c_vp, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp, c_i64, c_vp]),
    "snb_gather_segments": (c_int, [c_vp, c_i64, c_i64, c_int, c_vp]),
    "snb_maxpool3x3s2": (c_int, [c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_vp, c_i64, c_vp]),
    "snb_stem7x7_rows": (c_int, [c_vp, c_i64, c_i64, c_i64, c_i64, c_vp, c_i64, c_vp]),
    "snb_abn_forward": (c_int, [c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_int, ctypes.c_float, ctypes.c_float,
                                c_int, ctypes.c_float, c_vp, c_vp, c_vp, c_vp]),
    "snb_abn_backward": (c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_int, ctypes.c_float, c_int,
                                 ctypes.c_float, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "snb_bn_train_nhwc": (c_int, [c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_int, ctypes.c_float, ctypes.c_float, c_vp, c_vp,
                                  ctypes.c_float, c_vp, c_i64, c_int, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "snb_bn_backward_nhwc": (c_int, [c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, ctypes.c_float,
                                     ctypes.c_float, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "snb_channel_sum_nhwc": (c_int, [c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp]),
    "snb_maxpool3x3s2_backward": (c_int, [c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp]),
    "snb_ew_nhwc": (c_int, [c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_int, ctypes.c_float, c_vp]),
    "snb_bn_relu_nhwc": (c_int, [c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp]),
    "snb_nhwc_bf16_to_nchw_f32": (c_int, [c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp]),
    "snb_reduce_workspace_bytes": (c_i64, []),
    "snb_loss_iou_reduce": (c_int, [c_vp, c_vp, c_int, c_i64, ctypes.c_float, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "snb_loss_grad": (c_int, [c_vp, c_vp, c_int, c_i64, c_vp, c_vp, c_int, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                              ctypes.c_float, ctypes.c_float, ctypes.c_float, c_vp, c_vp]),
    "snb_confusion_counts": (c_int, [c_vp, c_vp, c_int, c_i64, ctypes.c_float, c_vp, c_vp, c_vp]),
    "snb_pr_curve_update": (c_int, [c_vp, c_vp, c_int, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
}

_lib = None


def lib():
    """The loaded library; raises ImportError (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        path = os.environ.get("SNB_B200_LIB", LIB_PATH)   # another build of the SAME ABI, for A/B timing (tools/build_rev.py)
        if not os.path.exists(path):
            raise ImportError(
                "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback." % path
            )
        handle = ctypes.CDLL(path)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def last_error():
    msg = lib().snb_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc):
    """Raise the reference's exception type for a non-zero return code."""
    if rc == SNB_OK:
        return
    msg = last_error()
    if rc == SNB_E_INVALID:
        raise ValueError(msg)
    if rc == SNB_E_SHAPE:
        raise AssertionError(msg)
    if rc == SNB_E_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)


def require_cuda():
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("segmentation-networks-benchmark_b200 runs on CUDA devices only (no CPU fallback)")


def stream_ptr():
    import torch

    return c_vp(torch.cuda.current_stream().cuda_stream)


_reduce_ws = {}


def reduce_workspace():
    """The "zero at rest" workspace of the single-launch reductions (include/snb_b200.h) for the current device and
    stream: allocated and zeroed once, then only touched by the library."""
    import torch

    dev = torch.cuda.current_device()
    key = (dev, torch.cuda.current_stream().cuda_stream)
    ws = _reduce_ws.get(key)
    if ws is None:
        ws = torch.zeros(int(lib().snb_reduce_workspace_bytes()), dtype=torch.uint8, device=torch.device("cuda", dev))
        _reduce_ws[key] = ws
    return ws


def ptr(t):
    """Device pointer of a torch tensor (must be CUDA)."""
    if t is None:
        return c_vp(0)
    if not t.is_cuda:
        raise RuntimeError("expected a CUDA tensor (no CPU fallback)")
    return c_vp(t.data_ptr())
