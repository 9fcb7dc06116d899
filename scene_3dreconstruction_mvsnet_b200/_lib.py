"""ctypes binding of libmvsnet_b200.so (the C ABI in include/mvsnet_b200.h).

There is no CPU fallback and no alternative backend: if the shared library is missing or a call
fails, a RuntimeError is raised.  Build it with `python -m scene_3dreconstruction_mvsnet_b200.build`
(or `__graft_entry__.build()`).
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmvsnet_b200.so")

MVS_OK = 0
PRECISION_FP32 = 0
PRECISION_BF16 = 1
COSTREG_LAYERS = 11
FEATURENET_LAYERS = 8

_c_float_p = ctypes.c_void_p  # device or host pointers are passed as raw addresses
_i = ctypes.c_int


class CostRegParams(ctypes.Structure):
    _fields_ = [("w", ctypes.c_void_p * COSTREG_LAYERS), ("shift", ctypes.c_void_p * COSTREG_LAYERS)]


class FeatureNetParams(ctypes.Structure):
    _fields_ = [("w", ctypes.c_void_p * FEATURENET_LAYERS), ("shift", ctypes.c_void_p * FEATURENET_LAYERS)]


# name -> (restype, argtypes); every symbol declared in include/mvsnet_b200.h
SIGNATURES = {
    "mvs_abi_version": (_i, []),
    "mvs_last_error": (ctypes.c_char_p, []),
    "mvs_launch_count": (ctypes.c_uint64, []),
    "mvs_arch": (ctypes.c_char_p, []),
    "mvs_homo_warping": (_i, [_c_float_p] * 5 + [_i] * 5 + [ctypes.c_void_p]),
    "mvs_homo_warping_bwd": (_i, [_c_float_p] * 5 + [_i] * 5 + [ctypes.c_void_p]),
    "mvs_warp_variance_workspace_bytes": (ctypes.c_size_t, [_i] * 5),
    "mvs_warp_variance_fwd": (_i, [_c_float_p] * 5 + [_i] * 6 + [ctypes.c_void_p]),
    "mvs_warp_variance_bwd_workspace_bytes": (ctypes.c_size_t, [_i] * 5),
    "mvs_warp_variance_bwd": (_i, [_c_float_p] * 6 + [_i] * 6 + [ctypes.c_void_p]),
    "mvs_conv3d_bn_relu": (_i, [_c_float_p] * 3 + [_i, _c_float_p] + [_i] * 7 + [ctypes.c_void_p]),
    "mvs_conv_transpose3d_bn_relu": (_i, [_c_float_p] * 3 + [_i, _c_float_p, _c_float_p] + [_i] * 6 + [ctypes.c_void_p]),
    "mvs_conv3d_bn_relu_tc": (_i, [_c_float_p] * 3 + [_i, _c_float_p] + [_i] * 7 + [ctypes.c_void_p]),
    "mvs_conv_transpose3d_bn_relu_tc": (_i, [_c_float_p] * 3 + [_i, _c_float_p, _c_float_p] + [_i] * 6 +
                                        [ctypes.c_void_p]),
    "mvs_weight_cache_clear": (_i, []),
    "mvs_tc_set_debug_buffer": (_i, [ctypes.c_void_p]),
    "mvs_tc_plan_describe": (_i, [_i] * 8 + [ctypes.c_char_p, _i]),
    "mvs_costreg_workspace_bytes": (ctypes.c_size_t, [_i] * 5),
    "mvs_costreg_fwd": (_i, [_c_float_p, ctypes.POINTER(CostRegParams), _c_float_p, ctypes.c_void_p] + [_i] * 5 +
                        [ctypes.c_void_p]),
    "mvs_volume_cp8_bytes": (ctypes.c_size_t, [_i] * 4),
    "mvs_warp_variance_fwd_cp8": (_i, [_c_float_p] * 5 + [_i] * 6 + [ctypes.c_void_p]),
    "mvs_warp_variance_fwd_cp8_f16": (_i, [_c_float_p] * 5 + [_i] * 6 + [ctypes.c_void_p]),
    "mvs_warp_variance_fwd_cp8_feat": (_i, [_c_float_p] * 5 + [_i] * 6 + [ctypes.c_void_p]),
    "mvs_warp_variance_fwd_cp8_pool": (_i, [_c_float_p, _i, ctypes.POINTER(ctypes.c_int)] + [_c_float_p] * 4 + [_i] * 6 +
                                       [ctypes.c_void_p]),
    "mvs_featurenet_tc_workspace_bytes": (ctypes.c_size_t, [_i] * 3),
    "mvs_featurenet_tc_fwd": (_i, [_c_float_p, ctypes.POINTER(FeatureNetParams), _c_float_p, ctypes.c_void_p] + [_i] * 3 +
                              [ctypes.c_void_p]),
    "mvs_featurenet_tc_fwd_u8": (_i, [_c_float_p, ctypes.POINTER(FeatureNetParams), _c_float_p, ctypes.c_void_p] + [_i] * 3 +
                                 [ctypes.c_void_p]),
    "mvs_featurenet_workspace_bytes": (ctypes.c_size_t, [_i] * 3),
    "mvs_featurenet_fwd": (_i, [_c_float_p, ctypes.POINTER(FeatureNetParams), _c_float_p, ctypes.c_void_p] + [_i] * 3 +
                           [ctypes.c_void_p]),
    "mvs_conv2d_bn_relu": (_i, [_c_float_p] * 3 + [_i, _c_float_p] + [_i] * 7 + [ctypes.c_void_p]),
    "mvs_conv2d_bn_relu_tc": (_i, [_c_float_p] * 3 + [_i, _c_float_p] + [_i] * 8 + [ctypes.c_void_p]),
    "mvs_costreg_fwd_cp8": (_i, [_c_float_p, ctypes.POINTER(CostRegParams), _c_float_p, ctypes.c_void_p] + [_i] * 4 +
                            [ctypes.c_void_p]),
    "mvs_filter_depth": (_i, [_c_float_p] * 7 + [_i] * 3 + [ctypes.c_double, ctypes.c_double, _i, ctypes.c_double] +
                         [_c_float_p] * 8 + [ctypes.c_void_p]),
    "mvs_softmax_depth_conf": (_i, [_c_float_p] * 5 + [_i] * 4 + [ctypes.c_void_p]),
    "mvs_depth_regression": (_i, [_c_float_p, _c_float_p, _i, _c_float_p] + [_i] * 4 + [ctypes.c_void_p]),
    "mvs_depth_from_features_host": (_i, [_c_float_p] * 3 + [ctypes.POINTER(CostRegParams), _c_float_p, _c_float_p] +
                                     [_i] * 7),
}

_lib = None
_lock = threading.Lock()


def load():
    """Load the native library (once).  Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    "libmvsnet_b200.so is missing (%s). This package has no CPU or PyTorch fallback: build the "
                    "CUDA library with `python -m scene_3dreconstruction_mvsnet_b200.build`." % LIB_PATH)
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
                fn.restype = res
                fn.argtypes = args
            if lib.mvs_abi_version() != 2:
                raise RuntimeError("libmvsnet_b200.so ABI version mismatch")
            _lib = lib
    return _lib


def check(rc, what):
    if rc != MVS_OK:
        msg = load().mvs_last_error().decode("utf-8", "replace")
        raise RuntimeError("%s failed (status %d): %s" % (what, rc, msg))


def launch_count():
    return int(load().mvs_launch_count())
