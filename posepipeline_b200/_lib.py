"""ctypes binding of libposeengine.so (include/poseengine.h).  No fallback: if the library is missing it
is built from csrc/ with nvcc; if that fails, importing raises."""
from __future__ import annotations

import atexit
import ctypes as C
import os
import re
import sys
import weakref

_PKG = os.path.dirname(os.path.abspath(__file__))
# PE_PRECISION=tf32 selects the wide-range TF32x3 build (libposeengine_tf32.so); default is the fp16x2 build
LIB_PATH = os.path.join(_PKG, "libposeengine_tf32.so" if os.environ.get("PE_PRECISION", "fp16") == "tf32" else "libposeengine.so")
HEADER_PATH = os.path.join(os.path.dirname(_PKG), "include", "poseengine.h")

PE_OK, PE_ERR_INVALID, PE_ERR_CUDA, PE_ERR_STATE, PE_ERR_NOGPU, PE_ERR_RANGE = 0, -1, -2, -3, -4, -5
PE_OP_STEM, PE_OP_CONV, PE_OP_FUSE, PE_OP_HEAD = 0, 1, 2, 3
PE_POST = {None: 0, "none": 0, "default": 1, "unbiased": 2, "udp": 3}


class PoseEngineError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libposeengine error {code}: {msg}")
        self.code = code


class OpDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("out", C.c_int32), ("inp", C.c_int32 * 4), ("up", C.c_int32 * 4),
                ("n_in", C.c_int32), ("ksize", C.c_int32), ("stride", C.c_int32), ("cin", C.c_int32),
                ("cout", C.c_int32), ("relu", C.c_int32), ("residual", C.c_int32), ("reserved", C.c_int32),
                ("w_off", C.c_int64), ("b_off", C.c_int64), ("wtc_off", C.c_int64)]


class TensorDesc(C.Structure):
    _fields_ = [("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("slot", C.c_int32)]


class ModelDesc(C.Structure):
    _fields_ = [("in_h", C.c_int32), ("in_w", C.c_int32), ("hm_h", C.c_int32), ("hm_w", C.c_int32),
                ("num_joints", C.c_int32), ("n_ops", C.c_int32), ("n_tensors", C.c_int32), ("n_slots", C.c_int32),
                ("max_crops", C.c_int32), ("flip_test", C.c_int32), ("shift_heatmap", C.c_int32),
                ("post_process", C.c_int32), ("blur_kernel", C.c_int32), ("swap_rb", C.c_int32),
                ("use_tensor_cores", C.c_int32), ("reserved", C.c_int32), ("padding", C.c_float),
                ("pixel_std", C.c_float)]


class GopDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("inp", C.c_int32), ("out", C.c_int32), ("res", C.c_int32), ("in_coff", C.c_int32),
                ("out_coff", C.c_int32), ("res_coff", C.c_int32), ("cin", C.c_int32), ("cout", C.c_int32), ("ksize", C.c_int32),
                ("stride", C.c_int32), ("act", C.c_int32), ("w_off", C.c_int64), ("b_off", C.c_int64), ("wtc_off", C.c_int64)]


class DetDesc(C.Structure):
    _fields_ = [("frame_h", C.c_int32), ("frame_w", C.c_int32), ("resized_h", C.c_int32), ("resized_w", C.c_int32),
                ("net_h", C.c_int32), ("net_w", C.c_int32), ("n_ops", C.c_int32), ("n_tensors", C.c_int32), ("n_slots", C.c_int32),
                ("max_frames", C.c_int32), ("max_candidates", C.c_int32), ("reserved", C.c_int32), ("score_thr", C.c_float),
                ("nms_iou", C.c_float), ("pad_val", C.c_float), ("reserved_f", C.c_float)]


_lib = None


def declared_symbols():
    """Function names declared in include/poseengine.h."""
    src = open(HEADER_PATH).read()
    return sorted(set(re.findall(r"^(?:int|const char\*)\s+(pe_[a-z0-9_]+)\s*\(", src, flags=re.M)))


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from .csrc.build import build
        build()
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    P = C.POINTER
    sig = {
        "pe_abi_version": (C.c_int, []),
        "pe_precision_mode": (C.c_int, []),
        "pe_last_error": (C.c_char_p, []),
        "pe_device_count": (C.c_int, [P(C.c_int)]),
        "pe_engine_create": (C.c_int, [C.c_int, vp, P(vp)]),
        "pe_engine_destroy": (C.c_int, [vp]),
        "pe_engine_sync": (C.c_int, [vp]),
        "pe_shutdown": (C.c_int, []),
        "pe_stage_frames": (C.c_int, [vp, vp, i32, i32, i32, i64]),
        "pe_stage_frames_device": (C.c_int, [vp, vp, i32, i32, i32]),
        "pe_frames_upload": (C.c_int, [vp, i32, vp, i32, i32, i32, i64]),
        "pe_frames_upload_to": (C.c_int, [vp, i32, vp, vp, i32, i32, i32, i64]),
        "pe_frames_select": (C.c_int, [vp, i32]),
        "pe_frames_slot_ptr": (C.c_int, [vp, i32, P(vp)]),
        "pe_warp_affine": (C.c_int, [vp, vp, vp, i32, i32, i32, i32, vp]),
        "pe_person_bbox": (C.c_int, [vp, i32, vp, vp, vp, i32, vp, vp]),
        "pe_model_create": (C.c_int, [vp, P(ModelDesc), P(OpDesc), P(TensorDesc), vp, vp, i64, vp, vp, P(vp)]),
        "pe_model_destroy": (C.c_int, [vp]),
        "pe_topdown": (C.c_int, [vp, vp, vp, i32, vp]),
        "pe_topdown_async": (C.c_int, [vp, vp, vp, i32, vp]),
        "pe_box_to_affine": (C.c_int, [P(ModelDesc), vp, vp, vp, vp]),
        "pe_warp_crops": (C.c_int, [vp, vp, vp, i32, vp, vp, vp]),
        "pe_forward_heatmaps": (C.c_int, [vp, vp, i32, vp, vp]),
        "pe_decode_heatmaps": (C.c_int, [vp, vp, vp, vp, vp, i32, vp]),
        "pe_debug_tensor": (C.c_int, [vp, i32, i32, vp]),
        "pe_model_launch_count": (C.c_int, [vp, P(i64)]),
        "pe_model_profile": (C.c_int, [vp, i32]),
        "pe_model_profile_read": (C.c_int, [vp, P(C.c_double), P(C.c_double), P(i64)]),
        "pe_model_profile_ops": (C.c_int, [vp, vp, i32]),
        "pe_conv_test": (C.c_int, [vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]),
        "pe_tc_plan_candidates": (C.c_int, [i32, i32, i32, i32, i32, i32, i32, i32, vp, i32]),
        "pe_tc_work_item": (C.c_int, [i32, i32, i32, i32, i32, i32, i32, i32, vp, vp]),
        "pe_lifter_create": (C.c_int, [vp, vp, i64, vp, i32, i32, P(vp)]),
        "pe_lifter_destroy": (C.c_int, [vp]),
        "pe_lift3d": (C.c_int, [vp, vp, i32, vp]),
        "pe_lifter_uses_tensor_cores": (C.c_int, [vp]),
        "pe_lifter_launch_count": (C.c_int, [vp, P(i64)]),
        "pe_detector_create": (C.c_int, [vp, P(DetDesc), P(GopDesc), P(TensorDesc), vp, vp, i64, P(vp)]),
        "pe_detector_destroy": (C.c_int, [vp]),
        "pe_detect": (C.c_int, [vp, vp, i32, vp, vp, i32]),
        "pe_detector_debug_tensor": (C.c_int, [vp, i32, i32, i32, i32, vp]),
        "pe_detector_launch_count": (C.c_int, [vp, P(i64)]),
        "pe_bytetrack_create": (C.c_int, [vp, i32, P(vp)]),
        "pe_bytetrack_destroy": (C.c_int, [vp]),
        "pe_bytetrack_reset": (C.c_int, [vp]),
        "pe_bytetrack_update": (C.c_int, [vp, i32, vp, i32, vp, i32, P(i32)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if lib.pe_abi_version() != 1:
        raise ImportError("libposeengine ABI mismatch; rebuild with python -m posepipeline_b200.csrc.build --force")
    _lib = lib
    atexit.register(shutdown)
    return lib


# Python objects holding a live C handle (PoseEngine / TopDownModel / Lifter ...).  The atexit hook destroys every engine
# (and with it every model / lifter) BEFORE interpreter finalisation and CUDA teardown, and clears the Python-side handles,
# so no __del__ touches CUDA afterwards: a worker process that used the wrappers exits 0.
_handles = weakref.WeakSet()


def track(obj):
    _handles.add(obj)
    return obj


def finalizing() -> bool:
    return sys.is_finalizing()


def shutdown():
    """Destroy all engines, models and lifters of this process (idempotent; also usable between tests)."""
    for o in list(_handles):
        try:
            o.h = None
        except Exception:
            pass
    if _lib is not None:
        _lib.pe_shutdown()


def check(rc):
    if rc != 0:
        raise PoseEngineError(rc, load().pe_last_error().decode())


def ptr(a):
    """numpy array -> void* (array must be C-contiguous and kept alive by the caller)."""
    return a.ctypes.data_as(C.c_void_p)
