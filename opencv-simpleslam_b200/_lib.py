"""ctypes binding of libb200slam.so (include/b200slam.h).  There is no CPU fallback: if the
shared library has not been built (`python -c "import __graft_entry__ as g; g.build()"` or
`make -C opencv-simpleslam_b200/csrc`) importing this module raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2S_LIB_PATH") or os.path.join(_HERE, "libb200slam.so")   # B2S_LIB_PATH: A/B builds of the library (development)

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing - build the CUDA library first (make -C {_HERE}/csrc). "
        "b200slam has no CPU fallback.")

lib = C.CDLL(LIB_PATH)


class AlikedCfg(C.Structure):
    _fields_ = [("model", C.c_int), ("max_kp", C.c_int), ("det_thresh", C.c_float),
                ("nms_radius", C.c_int), ("resize_long", C.c_int), ("precision", C.c_int)]


class LgCfg(C.Structure):
    _fields_ = [("n_layers", C.c_int), ("heads", C.c_int), ("dim", C.c_int), ("in_dim", C.c_int),
                ("depth_conf", C.c_float), ("width_conf", C.c_float), ("filter_thresh", C.c_float),
                ("pruning_min_kpts", C.c_int), ("precision", C.c_int), ("max_kp", C.c_int)]


FP32, BF16, FP32X3 = 0, 1, 2
PRECISIONS = {"fp32": FP32, "bf16": BF16, "fp32x3": FP32X3}
LG_RANGE = -2   # n_matches of a pair whose launch sequence left the fp16 operand range (B2S_FP32)
ALIKED_RANGE = -2   # keypoint count of a frame whose fp16x2 convolutions left the fp16 range (device-resident APIs)
ALIKED_RANGE_MSG = ("b200slam: an ALIKED activation left the fp16 range of the two-plane convolutions; "
                    "re-create the extractor with B2S_ALIKED_CONV_NP=3 (three bf16 planes)")
IMG_BGR_U8_HWC, IMG_RGB_F32_CHW = 0, 1
vp, i32p, f32p = C.c_void_p, C.c_void_p, C.c_void_p   # raw addresses (device or host)

SIGNATURES = {
    "b2s_version": (C.c_int, []),
    "b2s_last_error": (C.c_char_p, []),
    "b2s_device_count": (C.c_int, []),
    "b2s_aliked_default_cfg": (None, [C.POINTER(AlikedCfg)]),
    "b2s_lg_default_cfg": (None, [C.POINTER(LgCfg)]),
    "b2s_aliked_create": (C.c_int, [C.POINTER(AlikedCfg), vp, C.c_size_t, C.c_int, C.POINTER(vp)]),
    "b2s_aliked_destroy": (None, [vp]),
    "b2s_aliked_extract": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, f32p, f32p, f32p, i32p]),
    "b2s_aliked_extract_batch": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, f32p, f32p, f32p, i32p]),
    "b2s_aliked_set_batch_renorm": (C.c_int, [vp, C.c_float]),
    "b2s_aliked_extract_host": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, f32p, f32p, f32p, i32p]),
    "b2s_aliked_extract_host_ex": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, f32p, f32p, f32p, i32p, C.c_float]),
    "b2s_aliked_extract_host_begin": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float]),
    "b2s_aliked_extract_host_keypoints": (C.c_int, [vp, f32p, i32p]),
    "b2s_aliked_extract_host_finish": (C.c_int, [vp, f32p, f32p]),
    "b2s_aliked_copy_last_features": (C.c_int, [vp, f32p, f32p, C.c_int, vp]),
    "b2s_lightglue_create": (C.c_int, [C.POINTER(LgCfg), vp, C.c_size_t, C.c_int, C.POINTER(vp)]),
    "b2s_lg_destroy": (None, [vp]),
    "b2s_lightglue_match": (C.c_int, [vp, f32p, f32p, C.c_int, f32p, f32p, C.c_int, f32p, f32p, vp,
                                      i32p, f32p, i32p, i32p, i32p, i32p, f32p, f32p, i32p, i32p]),
    "b2s_lightglue_match_host": (C.c_int, [vp, f32p, f32p, C.c_int, f32p, f32p, C.c_int, f32p, f32p,
                                           i32p, f32p, i32p, i32p, i32p, i32p, f32p, f32p, i32p, i32p]),
    "b2s_lightglue_match_batch": (C.c_int, [vp, f32p, f32p, i32p, C.c_int, i32p, i32p, C.c_int, vp, C.c_int,
                                            i32p, f32p, i32p]),
    "b2s_lightglue_match_batch_ex": (C.c_int, [vp, f32p, f32p, i32p, i32p, i32p, C.c_int, i32p, i32p, C.c_int, vp, C.c_int, C.c_int,
                                               i32p, f32p, i32p, i32p]),
    "b2s_lg_max_batch": (C.c_int, []),
    "b2s_lg_workspace_bytes": (C.c_size_t, [C.POINTER(LgCfg), C.c_int, C.c_int]),
    "b2s_lg_reserve": (C.c_int, [vp, C.c_int, C.c_int]),
    "b2s_lg_fallback": (C.c_int, [vp, C.POINTER(vp)]),
    "b2s_lg_range_fallbacks": (C.c_longlong, [vp]),
    "b2s_lg_planes": (C.c_int, [vp]),
    "b2s_test_gemm_h2": (C.c_int, [f32p, f32p, f32p, C.c_int, C.c_int, C.c_int, f32p]),
    "b2s_test_attn_h2": (C.c_int, [f32p, f32p, f32p, C.c_int, C.c_int, f32p]),
    "b2s_bench_attn_h2": (C.c_int, [C.c_int, C.c_int, C.c_int, f32p]),
    "b2s_trace_attn_h2": (C.c_int, [C.c_int, C.c_int, C.c_int, f32p, vp]),
    "b2s_bench_gemm_h2": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, f32p, vp, i32p]),
    "b2s_aliked_debug_get": (C.c_int, [vp, C.c_char_p, f32p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "b2s_lg_set_debug": (C.c_int, [vp, C.c_int]),
    "b2s_lg_debug_get": (C.c_int, [vp, C.c_char_p, f32p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "b2s_lg_profile": (C.c_int, [vp, C.c_int]),
    "b2s_lg_profile_read": (C.c_int, [vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "b2s_lg_profile_work": (C.c_int, [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "b2s_test_gemm_tc": (C.c_int, [f32p, f32p, f32p, C.c_int, C.c_int, C.c_int, f32p]),
    "b2s_test_attn_tc": (C.c_int, [f32p, f32p, f32p, C.c_int, C.c_int, f32p]),
    "b2s_bench_attn_tc": (C.c_int, [C.c_int, C.c_int, C.c_int, f32p]),
    "b2s_test_gemm_tc3": (C.c_int, [f32p, f32p, f32p, C.c_int, C.c_int, C.c_int, f32p]),
    "b2s_test_attn_tc3": (C.c_int, [f32p, f32p, f32p, C.c_int, C.c_int, f32p]),
    "b2s_bench_attn_tc3": (C.c_int, [C.c_int, C.c_int, C.c_int, f32p]),
    "b2s_trace_attn_tc3": (C.c_int, [C.c_int, C.c_int, C.c_int, f32p, vp]),
    "b2s_bench_gemm_tc3": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, f32p, vp, i32p]),
    "b2s_aliked_launch_count": (C.c_longlong, [vp]),
    "b2s_lg_launch_count": (C.c_longlong, [vp]),
    "b2s_fm_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(vp)]),
    "b2s_fm_destroy": (None, [vp]),
    "b2s_fm_ransac": (C.c_int, [vp, f32p, f32p, i32p, C.c_int, C.c_float, C.c_int, C.c_uint64, vp, vp, vp, i32p]),
    "b2s_fm_ransac_host": (C.c_int, [vp, f32p, f32p, C.c_int, C.c_float, C.c_int, C.c_uint64, vp, vp, i32p, i32p]),
    "b2s_fm_cv_ransac_host": (C.c_int, [vp, f32p, f32p, C.c_int, C.c_double, C.c_double, C.c_int, vp, vp, i32p, i32p]),
    "b2s_fm_debug_models": (C.c_int, [vp, C.c_int, vp, i32p, i32p]),
    "b2s_fm_launch_count": (C.c_longlong, [vp]),
    "b2s_remap_create": (C.c_int, [C.c_int, f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]),
    "b2s_remap_destroy": (None, [vp]),
    "b2s_remap_bgr": (C.c_int, [vp, vp, C.c_int, vp, vp, C.c_int]),
    "b2s_remap_bgr_host": (C.c_int, [vp, vp, C.c_int, vp, C.c_int]),
    "b2s_remap_output_dev": (vp, [vp]),
    "b2s_remap_launch_count": (C.c_longlong, [vp]),
    "b2s_reproj_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]),
    "b2s_reproj_destroy": (None, [vp]),
    "b2s_reproj_match": (C.c_int, [vp, vp, f32p, i32p, i32p, C.c_int, C.c_int, vp, vp, f32p, f32p, C.c_int, C.c_int, C.c_int,
                                   C.c_double, C.c_double, vp, i32p, f32p, i32p]),
    "b2s_reproj_launch_count": (C.c_longlong, [vp]),
    "b2s_remap_dims": (None, [vp, i32p, i32p, i32p, i32p]),
    "b2s_aliked_set_undistort": (C.c_int, [vp, vp]),
}
for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)   # AttributeError here == header/library mismatch
    _fn.restype, _fn.argtypes = _res, _args


class B2SError(RuntimeError):
    pass


def check(rc: int, what: str):
    if rc != 0:
        raise B2SError(f"{what} failed ({rc}): {lib.b2s_last_error().decode(errors='replace')}")
