"""ctypes binding of libhvs_b200.so (the C ABI in include/hvs_b200.h).

There is no CPU or PyTorch fallback: if the library is missing or a call
fails, the caller gets an exception.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint32, c_uint64, c_void_p

from . import build as _build

_LIB = None

HVS_MHC_SPLIT_PHI = 1
HVS_MHC_SAVED_STRIDE = 28
HVS_DTYPE_F32, HVS_DTYPE_F16, HVS_DTYPE_BF16 = 0, 1, 2
HVS_NMS_AGNOSTIC, HVS_NMS_CLASS_AWARE, HVS_NMS_BOXES_XYXY = 0, 1, 16

# name -> (restype, argtypes); mirrors include/hvs_b200.h one to one
_SIGNATURES = {
    "hvs_abi_version": (c_int, []),
    "hvs_error_string": (c_char_p, [c_int]),
    "hvs_launch_count": (c_uint64, []),
    "hvs_mhc_stream_profile": (c_int, [c_int]),
    "hvs_mhc_stream_kernel_ms": (c_int, [POINTER(c_float)]),
    "hvs_mhc_stream_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_int64, c_int, c_int, c_int, c_float, c_float, c_uint32, c_void_p]),
    "hvs_mhc_stream_fwd_save": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_int64, c_int, c_int, c_int, c_float, c_float, c_uint32, c_void_p]),
    "hvs_mhc_stream_post": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "hvs_mhc_stream_bwd_workspace": (c_size_t, [c_int64, c_int, c_int]),
    "hvs_mhc_stream_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_float, c_float,
                                   c_uint32, c_void_p, c_size_t, c_void_p]),
    "hvs_mhc_stream_bwd_saved_workspace": (c_size_t, [c_int64, c_int, c_int]),
    "hvs_mhc_stream_bwd_saved": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_float,
                                         c_float, c_uint32, c_void_p, c_size_t, c_void_p]),
    "hvs_sinkhorn": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_float, c_float, c_void_p, c_void_p]),
    "hvs_mhc_constrained_matrices": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                             c_int, c_float, c_void_p, c_void_p]),
    "hvs_yolo_decode": (c_int, [c_void_p, c_int, POINTER(c_int64), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "hvs_nms": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_float, c_float, c_int, c_int,
                        c_void_p, c_void_p, c_void_p, c_void_p]),
    "hvs_post_process_workspace": (c_size_t, [c_int, c_int, c_int]),
    "hvs_post_process": (c_int, [POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), POINTER(c_int), c_int, c_int,
                                 c_float, c_float, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                 c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


class HvsError(RuntimeError):
    pass


def lib_path() -> str:
    return _build.LIB_PATH


def load(build_if_missing: bool = True):
    """Load (building first if the in-tree .so is absent or stale and nvcc is here)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB_PATH
    if build_if_missing and not _build.is_fresh():
        try:
            _build.build()
        except Exception as exc:  # no nvcc on this box: use the prebuilt library if there is one
            if not os.path.exists(path):
                raise HvsError(f"libhvs_b200.so is not built and cannot be built here: {exc}") from exc
    if not os.path.exists(path):
        raise HvsError(f"{path} not found; run `python -m hvs_b200.build`")
    lib = ctypes.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the ABI and the header drifted apart
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(code: int, what: str):
    if code != 0:
        msg = load().hvs_error_string(code).decode()
        raise HvsError(f"{what} failed: {msg} (code {code})")


def launch_count() -> int:
    return int(load().hvs_launch_count())
