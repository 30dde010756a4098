"""ctypes binding of libunetb200.so (C ABI declared in include/unet_b200.h).

There is no CPU fallback: importing this module never computes anything, but every compute entry
point raises ``RuntimeError`` when the shared library is missing or no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_PATH = os.path.join(HERE, "lib", "libunetb200.so")
HEADER_PATH = os.path.join(ROOT, "include", "unet_b200.h")

c_void_p, c_int, c_int64, c_float = C.c_void_p, C.c_int, C.c_int64, C.c_float


class UbView(C.Structure):
    """Mirror of ``ub_view``: NHWC bf16 view with element strides."""

    _fields_ = [("ptr", c_void_p), ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("C", C.c_int32), ("sN", c_int64), ("sH", c_int64), ("sW", c_int64)]


_VP = C.POINTER(UbView)
_P = c_void_p  # every device pointer travels as void*

# name -> (restype, argtypes); kept in the order of include/unet_b200.h
_SIGNATURES = {
    "ub_last_error": (C.c_char_p, []),
    "ub_version": (c_int, []),
    "ub_device_sm_count": (c_int, []),
    "ub_plan_create": (c_int, [C.POINTER(c_void_p)] + [c_int] * 8),
    "ub_plan_create_ex": (c_int, [C.POINTER(c_void_p)] + [c_int] * 9),
    "ub_plan_destroy": (c_int, [_P]),
    "ub_plan_out_hw": (c_int, [_P, C.POINTER(c_int), C.POINTER(c_int)]),
    "ub_plan_num_params": (c_int, [_P]),
    "ub_plan_num_bn": (c_int, [_P]),
    "ub_plan_param_numel": (c_int64, [_P, c_int]),
    "ub_plan_device_bytes": (c_int64, [_P]),
    "ub_plan_bind_params": (c_int, [_P, C.POINTER(c_void_p), c_int]),
    "ub_plan_bind_bn_buffers": (c_int, [_P, C.POINTER(c_void_p), C.POINTER(c_void_p),
                                        C.POINTER(c_void_p), c_int]),
    "ub_plan_set_bn_config": (c_int, [_P, C.POINTER(c_float), C.POINTER(c_float), c_int]),
    "ub_plan_pack_weights": (c_int, [_P, _P]),
    "ub_plan_forward": (c_int, [_P, _P, _P, _P, _P]),
    "ub_plan_graph_replays": (c_int64, [_P]),
    "ub_plan_num_stages": (c_int, [_P]),
    "ub_plan_stage_params": (c_int, [_P, c_int, C.POINTER(c_int), C.POINTER(c_int)]),
    "ub_plan_backward_stage": (c_int, [_P, c_int, _P, C.POINTER(c_void_p), _P]),
    "ub_plan_set_overlap": (c_int, [_P, c_int]),
    "ub_plan_join_side": (c_int, [_P, _P]),
    "ub_plan_sgd_step": (c_int, [_P, C.POINTER(c_void_p), C.POINTER(c_void_p), c_float, c_float,
                                 c_float, c_float, c_int, c_int, _P]),
    "ub_launch_count": (c_int64, []),
    "ub_plan_profile_enable": (c_int, [_P, c_int]),
    "ub_plan_profile_classes": (c_int, []),
    "ub_plan_profile_class_name": (C.c_char_p, [c_int]),
    "ub_plan_profile_collect": (c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                        C.POINTER(C.c_double), C.POINTER(c_int)]),
    "ub_wce_workspace_floats": (c_int64, []),
    "ub_wce_forward": (c_int, [_P, C.POINTER(c_int64), _P, C.POINTER(c_int64), _P,
                               C.POINTER(c_int64), c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P]),
    "ub_scale_by_device_scalar": (c_int, [_P, _P, _P, c_int64, _P]),
    "ub_prepare_batch": (c_int, [_P, _P, c_int, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P,
                                 _P, _P]),
    "ub_weight_map": (c_int, [_P, c_int, c_int, c_int, c_int, C.c_double, C.c_double, _P, c_int, _P,
                              _P]),
    "ub_elastic_workspace_bytes": (c_int64, [c_int, c_int, c_int]),
    "ub_elastic_deform": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P, _P, c_int, C.c_double, _P, _P,
                                  c_int, _P, _P]),
    "ub_ccl_workspace_bytes": (c_int64, [c_int, c_int]),
    "ub_ccl_label": (c_int, [_P, c_int, c_int, c_int, _P, _P, _P]),
    "ub_extract_tiles": (c_int, [_P, c_int, c_int, _P, c_int, c_int, c_int, _P, _P]),
    "ub_stitch_tiles": (c_int, [_P, _P, c_int, c_int, _P, c_int, c_int, _P]),
    "ub_op_pack_conv3x3": (c_int, [_P, c_int, c_int, _P, _P, _P]),
    "ub_op_pack_convT": (c_int, [_P, c_int, c_int, _P, _P, _P, _P, _P]),
    "ub_op_conv_stats_floats": (c_int64, [c_int]),
    "ub_op_conv3x3_forward": (c_int, [_VP, _VP, _P, _P, c_int, c_int, _P, _P, _P, _P,
                                      C.POINTER(c_int), _P]),
    "ub_op_conv3x3_affine_relu_head": (c_int, [_VP, _VP, _P, _P, _P, _P, _P, c_int, _P, _P, _P]),
    "ub_op_conv3x3_dgrad": (c_int, [_VP, _P, c_int, _P, _P]),
    "ub_op_upsample2x_forward": (c_int, [_VP, _P, _P]),
    "ub_op_upsample2x_backward": (c_int, [_VP, _P, _P]),
    "ub_op_conv3x3_dgrad_bnred": (c_int, [_VP, _P, c_int, _P, _P, _P, _P, _P, _P, C.POINTER(c_int), _P]),
    "ub_op_bn_relu_backward_fused": (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _VP, _P,
                                             C.POINTER(c_int), _P, _P, _P, _P]),
    "ub_op_wgrad_workspace_floats": (c_int64, [c_int, c_int, c_int64]),
    "ub_op_conv3x3_wgrad": (c_int, [_VP, _VP, _P, c_int, _P, c_int64, _P, _P]),
    "ub_op_convT_forward": (c_int, [_VP, _P, _P, c_int, _VP, _P]),
    "ub_op_convT_dgrad": (c_int, [_VP, _P, c_int, _P, _P]),
    "ub_op_convT_wgrad": (c_int, [_VP, _P, c_int, _P, c_int64, _P, _P]),
    "ub_op_bn_finalize": (c_int, [_P, C.POINTER(c_int), c_int, _P, _P, _P, _P, _P, c_float, c_float,
                                  _P, _P, _P, _P, _P]),
    "ub_op_bn_apply_relu": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P, _P]),
    "ub_op_bn_apply_relu_head": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P, _P, c_int, _P, _P,
                                         _P, _P]),
    "ub_op_bn_bwd_workspace_floats": (c_int64, [c_int]),
    "ub_op_bn_relu_backward": (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _VP, _VP,
                                       _VP, c_int, c_int, _P, _P, _P, _P, _P, _P]),
    "ub_op_first_conv_workspace_floats": (c_int64, [c_int]),
    "ub_op_first_conv_forward": (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P, c_int, _P, _P, _P,
                                         _P, _P, c_float, c_float, _P, _P, _P, _P, _P, _P, _P]),
    "ub_op_first_conv_affine_relu": (c_int, [_P, c_int, c_int, c_int, c_int, _P, c_int, _P, _P, _P, _P]),
    "ub_op_first_conv_backward": (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P, c_int, _P, _P, _P,
                                          _P, _VP, _P, _P, _P, _P, _P, _P]),
    "ub_op_head_forward": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P]),
    "ub_op_head_bwd_workspace_floats": (c_int64, [c_int, c_int]),
    "ub_op_head_backward": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P,
                                    _P]),
    "ub_op_maxpool2": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P]),
}


def declared_symbols() -> list[str]:
    """Every function name declared in include/unet_b200.h."""
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ub_[a-zA-Z0-9_]+)\s*\(", text)))


_lib = None


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load the shared library, building it with nvcc if it is absent. Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build

    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise RuntimeError(f"{LIB_PATH} is missing; run `python -m unet_segmentation_b200.build`")
        _build.build()
    elif not _build.is_current():
        # a library from older sources must never load silently; rebuild when nvcc is here (the
        # authoring container), otherwise refuse (a GPU box receives the in-tree .so with its stamp)
        if build_if_missing and _build.have_nvcc():
            _build.build()
        else:
            raise RuntimeError(f"{LIB_PATH} does not match the sources under csrc/ (hash stamp "
                               "differs); run `python -m unet_segmentation_b200.build`")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    msg = load().ub_last_error()
    return msg.decode() if msg else ""


def check(status: int, what: str = "libunetb200") -> None:
    if status != 0:
        raise RuntimeError(f"{what} failed with status {status}: {last_error()}")


def require_cuda() -> None:
    """Raise unless a CUDA device is usable; the product path never falls back to the CPU."""
    n = load().ub_device_sm_count()
    if n <= 0:
        raise RuntimeError(f"libunetb200 needs a CUDA device (sm_100a): {last_error()}")
