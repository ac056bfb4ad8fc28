"""ctypes binding of libyms_b200.so (the C ABI declared in include/yms_b200.h).

There is no CPU implementation: if the library is missing or a CUDA device is absent the
callers raise.  The library is built in-tree by ``python -m yolo_ms_b200.build``.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("YMS_LIB") or os.path.join(_HERE, "libyms_b200.so")   # YMS_LIB: profiling build (scripts/role_prof.py)

EXPORTS = [
    "yms_abi_version", "yms_last_error", "yms_launch_count",
    "yms_conv_plan_create", "yms_conv_plan_run", "yms_conv_plan_destroy", "yms_conv_plan_cost", "yms_conv_plan_fuse_decode", "yms_conv_plan_add_upsampled",
    "yms_ms_plan_create", "yms_ms_plan_run", "yms_ms_plan_destroy", "yms_ms_plan_cost",
    "yms_stem_conv", "yms_stem_conv_u8", "yms_resample_u8", "yms_dwconv", "yms_sppf_pool", "yms_upsample2x",
    "yms_head_decode", "yms_select_candidates",
    "yms_nms_workspace_bytes", "yms_nms_batched", "yms_gather_detections",
]

DTYPE_F32, DTYPE_BF16 = 0, 1


class YmsError(RuntimeError):
    pass


class ConvParams(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("in_h", C.c_int32), ("in_w", C.c_int32),
        ("c_in", C.c_int32), ("c_out", C.c_int32), ("ksize", C.c_int32), ("stride", C.c_int32),
        ("act", C.c_int32), ("out_dtype", C.c_int32), ("c_in2", C.c_int32), ("variant", C.c_int32),
        ("x", C.c_void_p), ("x_pixel_stride", C.c_int64),
        ("x2", C.c_void_p), ("x2_pixel_stride", C.c_int64),
        ("y", C.c_void_p), ("y_pixel_stride", C.c_int64),
        ("residual", C.c_void_p), ("res_pixel_stride", C.c_int64),
        ("weight", C.c_void_p), ("bias", C.c_void_p),
    ]


class DecodeFusion(C.Structure):
    """yms_decode_fusion (include/yms_b200.h)."""
    _fields_ = [
        ("branch", C.c_int32), ("map_h", C.c_int32), ("map_w", C.c_int32), ("anchor_base", C.c_int32),
        ("anchors", C.c_int32), ("num_classes", C.c_int32),
        ("stride", C.c_void_p), ("pred", C.c_void_p),
        ("cand_boxes", C.c_void_p), ("cand_scores", C.c_void_p), ("cand_labels", C.c_void_p),
    ]


class MsParams(C.Structure):
    """yms_ms_params (include/yms_b200.h)."""
    _fields_ = [
        ("mode", C.c_int32), ("batch", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("ksize", C.c_int32),
        ("e_ch", C.c_int32), ("c_out", C.c_int32), ("act2", C.c_int32), ("c_in", C.c_int32), ("c_in2", C.c_int32),
        ("e", C.c_void_p), ("e_pixel_stride", C.c_int64),
        ("x", C.c_void_p), ("x_pixel_stride", C.c_int64),
        ("x2", C.c_void_p), ("x2_pixel_stride", C.c_int64),
        ("y", C.c_void_p), ("y_pixel_stride", C.c_int64),
        ("w1", C.c_void_p), ("bias1", C.c_void_p), ("dw_weight", C.c_void_p), ("dw_bias", C.c_void_p),
        ("w2", C.c_void_p), ("bias2", C.c_void_p),
    ]


_lib = None


def load() -> C.CDLL:
    """Load (building first if the .so is absent and nvcc exists).  Raises YmsError otherwise."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
            from . import build as _build
            _build.build()
        else:
            raise YmsError(f"{LIB_PATH} is missing and nvcc is unavailable; build with `python -m yolo_ms_b200.build`")
    lib = C.CDLL(LIB_PATH)
    missing = [s for s in EXPORTS if not hasattr(lib, s)]
    if missing:
        raise YmsError(f"libyms_b200.so lacks symbols {missing}")
    vp, i32, i64, f32, f64, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double, C.c_size_t
    lib.yms_abi_version.restype = C.c_int
    lib.yms_last_error.restype = C.c_char_p
    lib.yms_launch_count.restype = C.c_longlong
    lib.yms_conv_plan_create.argtypes = [C.POINTER(ConvParams), C.POINTER(vp)]
    lib.yms_conv_plan_run.argtypes = [vp, vp]
    lib.yms_conv_plan_destroy.argtypes = [vp]
    lib.yms_conv_plan_cost.argtypes = [vp, C.POINTER(f64), C.POINTER(f64)]
    lib.yms_conv_plan_fuse_decode.argtypes = [vp, C.POINTER(DecodeFusion)]
    lib.yms_conv_plan_add_upsampled.argtypes = [vp, vp, i64, i32, i32]
    lib.yms_ms_plan_create.argtypes = [C.POINTER(MsParams), C.POINTER(vp)]
    lib.yms_ms_plan_run.argtypes = [vp, vp]
    lib.yms_ms_plan_destroy.argtypes = [vp]
    lib.yms_ms_plan_cost.argtypes = [vp, C.POINTER(f64), C.POINTER(f64)]
    lib.yms_stem_conv.argtypes = [vp, i32, i32, i32, i32, vp, vp, vp, i64, vp]
    lib.yms_stem_conv_u8.argtypes = [vp, i32, i32, i32, i32, vp, vp, C.POINTER(f32), C.POINTER(f32), vp, i64, vp]
    lib.yms_resample_u8.argtypes = [vp, i32, i32, i32, i64, vp, i32, i32, i64, vp, vp, i32, i32, vp]
    lib.yms_dwconv.argtypes = [vp, i64, i32, i32, i32, i32, i32, vp, vp, vp, i64, vp]
    lib.yms_sppf_pool.argtypes = [vp, i64, i32, i32, i32, i32, vp]
    lib.yms_upsample2x.argtypes = [vp, i64, i32, i32, i32, i32, vp, i64, vp]
    lib.yms_head_decode.argtypes = [vp, vp, vp, i32, i32, C.POINTER(i32), i32, C.POINTER(f32), vp, vp, vp, vp, vp]
    lib.yms_select_candidates.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp]
    lib.yms_nms_workspace_bytes.restype = sz
    lib.yms_nms_workspace_bytes.argtypes = [i32, i32]
    lib.yms_nms_batched.argtypes = [vp, vp, vp, vp, i32, i32, i32, f32, f64, vp, vp, vp, sz, vp]
    lib.yms_gather_detections.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, vp, vp]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int and name not in ("yms_abi_version",):
            fn.restype = C.c_int
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().yms_last_error().decode(errors="replace")
        raise YmsError(f"{what} failed (code {rc}): {msg}")


def set_debug_option(name: str, value: int) -> None:
    """Experiment switches of the library (csrc/common.cuh::DebugOptions): 'pdl_off', 'conv_half', 'stem_gather', 'nms_groups',
    'nms_mask_tiles', 'nms_poll_ns', 'nms_sort_bitonic'.  Not part of the reference-facing ABI; affects plans created and
    kernels launched after the call.  (The library never reads environment variables.)"""
    lib = load()
    lib.yms_debug_set_option.argtypes = [C.c_char_p, C.c_int]
    check(lib.yms_debug_set_option(name.encode(), int(value)), "yms_debug_set_option")


def launch_count() -> int:
    return int(load().yms_launch_count())
