"""ctypes loader for libvarpro_b200.so (the C ABI in include/varpro_b200.h).

The product path has no CPU fallback: if the CUDA extension is missing this
module raises at import time of the bindings, and every compute entry point
fails with VP_ERR_CUDA when no device is present.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libvarpro_b200.so")

VP_MAX_BASIS_PARAMS = 4
VP_MAX_N = 8
VP_MAX_Q = 8
VP_MAX_P = 12

VP_F64, VP_F32 = 0, 1


class BasisDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_params", C.c_int32),
                ("param_idx", C.c_int32 * VP_MAX_BASIS_PARAMS), ("scale", C.c_double)]


class LmOptions(C.Structure):
    _fields_ = [("ftol", C.c_double), ("xtol", C.c_double), ("gtol", C.c_double),
                ("stepbound", C.c_double), ("patience", C.c_int32), ("scale_diag", C.c_int32)]


class FitReport(C.Structure):
    _fields_ = [("termination", C.c_int32), ("number_of_evaluations", C.c_int32),
                ("objective_function", C.c_double), ("successful", C.c_int32),
                ("reserved", C.c_int32)]


class Reduced(C.Structure):
    _fields_ = [("rnorm2", C.c_double), ("g", C.c_double * VP_MAX_Q),
                ("H", C.c_double * (VP_MAX_Q * VP_MAX_Q)), ("finite", C.c_int32), ("q", C.c_int32)]


HOST_EVAL_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double))

# every symbol include/varpro_b200.h declares: name -> (restype, argtypes)
_vp = C.c_void_p
_pp = C.POINTER(C.c_void_p)
_dp = C.POINTER(C.c_double)
SYMBOLS = {
    "vp_abi_version": (C.c_int, []),
    "vp_status_string": (C.c_char_p, [C.c_int]),
    "vp_ctx_create": (C.c_int, [C.c_int, _pp]),
    "vp_ctx_destroy": (C.c_int, [_vp]),
    "vp_last_error": (C.c_char_p, [_vp]),
    "vp_ctx_kernel_launches": (C.c_int64, [_vp]),
    "vp_ctx_stream": (C.c_void_p, [_vp]),
    "vp_ctx_set_option": (C.c_int, [_vp, C.c_char_p, C.c_char_p]),
    "vp_ctx_trim": (C.c_int, [_vp]),
    "vp_model_create": (C.c_int, [_vp, C.c_int, C.c_int64, _vp, C.c_int32, C.c_int32,
                                  C.POINTER(BasisDesc), _pp]),
    "vp_model_create_hosteval": (C.c_int, [_vp, C.c_int, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                           C.POINTER(C.c_int32), C.c_void_p, _vp, _pp]),
    "vp_model_destroy": (C.c_int, [_vp]),
    "vp_problem_create": (C.c_int, [_vp, _vp, C.c_int64, _vp, C.c_int64, _vp, C.c_double, _dp, _pp]),
    "vp_problem_create_device": (C.c_int, [_vp, _vp, C.c_int64, _vp, C.c_int64, _vp, C.c_double, _dp, _pp]),
    "vp_problem_destroy": (C.c_int, [_vp]),
    "vp_set_params": (C.c_int, [_vp, _dp]),
    "vp_params": (C.c_int, [_vp, _dp]),
    "vp_residuals": (C.c_int, [_vp, _vp]),
    "vp_jacobian": (C.c_int, [_vp, _vp]),
    "vp_linear_coefficients": (C.c_int, [_vp, _vp]),
    "vp_best_fit": (C.c_int, [_vp, _vp]),
    "vp_residuals_device": (C.c_int, [_vp, _vp]),
    "vp_jacobian_device": (C.c_int, [_vp, _vp]),
    "vp_best_fit_device": (C.c_int, [_vp, _vp]),
    "vp_problem_set_jacobian": (C.c_int, [_vp, C.c_int]),
    "vp_problem_set_rank_policy": (C.c_int, [_vp, C.c_int]),
    "vp_reduce": (C.c_int, [_vp, C.POINTER(Reduced)]),
    "vp_comm_create": (C.c_int, [_vp, C.c_int, C.c_int, _pp, _vp]),
    "vp_comm_connect": (C.c_int, [_vp, _vp]),
    "vp_comm_connect_local": (C.c_int, [_pp, C.c_int]),
    "vp_comm_destroy": (C.c_int, [_vp]),
    "vp_problem_set_comm": (C.c_int, [_vp, _vp]),
    "vp_fit": (C.c_int, [_vp, C.POINTER(LmOptions), C.POINTER(FitReport)]),
    "vp_fit_many": (C.c_int, [_pp, C.c_int64, C.POINTER(LmOptions), C.POINTER(FitReport), C.c_int32]),
    "vp_fit_host_batch": (C.c_int, [_vp, C.c_int, C.c_int64, _vp, C.c_int32, C.c_int32, C.POINTER(BasisDesc), C.c_int64, C.c_int64,
                                    _pp, C.c_int64, _vp, C.c_double, _dp, C.POINTER(LmOptions), C.c_int32, C.POINTER(FitReport),
                                    _dp, _pp]),
    "vp_statistics": (C.c_int, [_vp, _dp, _dp, _dp]),
    "vp_batch_create": (C.c_int, [_vp, _vp, C.c_int64, _vp, C.c_int64, _vp, C.c_double, _dp, _pp]),
    "vp_batch_create_device": (C.c_int, [_vp, _vp, C.c_int64, _vp, C.c_int64, _vp, C.c_double, _dp, _pp]),
    "vp_batch_destroy": (C.c_int, [_vp]),
    "vp_batch_fit": (C.c_int, [_vp, C.POINTER(LmOptions), C.POINTER(FitReport)]),
    "vp_batch_params": (C.c_int, [_vp, _dp]),
    "vp_batch_set_params": (C.c_int, [_vp, _dp]),
    "vp_batch_set_rank_policy": (C.c_int, [_vp, C.c_int]),
    "vp_batch_linear_coefficients": (C.c_int, [_vp, _dp]),
    "vp_measure_fp64_peaks": (C.c_int, [_vp, _dp, _dp]),
    "vp_debug_timeline": (C.c_int, [_vp, C.POINTER(C.c_longlong), C.c_int64, C.POINTER(C.c_int64)]),
    "vp_profile_evaluation": (C.c_int, [_vp, C.c_int, C.c_int64, _dp, _dp, C.POINTER(C.c_int64),
                                        C.POINTER(C.c_int64)]),
}

_lib = None


class ExtensionMissingError(ImportError):
    pass


def load():
    """Load the CUDA extension. Raises loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ExtensionMissingError(
            f"{LIB_PATH} not found: build it with `python -m varpro_b200.build` "
            "(there is no CPU fallback for the VarPro hot path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
