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
HVS_MHC_ADAPTIVE_ITERS = 2
HVS_MHC_SAVED_STRIDE = 28
HVS_DTYPE_F32, HVS_DTYPE_F16, HVS_DTYPE_BF16 = 0, 1, 2
HVS_NMS_AGNOSTIC, HVS_NMS_CLASS_AWARE, HVS_NMS_BOXES_XYXY = 0, 1, 16
HVS_GEMM_EPI_NONE, HVS_GEMM_EPI_BIAS_GELU, HVS_GEMM_EPI_LAYERNORM = 0, 1, 2
HVS_GEMM_EPI_BIAS_GELU_SAVE, HVS_GEMM_EPI_DGELU = 3, 4


class CoeffJob(ctypes.Structure):
    """struct hvs_coeff_job (include/hvs_b200.h)."""
    _fields_ = [("h_pre_raw", c_void_p), ("h_post_raw", c_void_p), ("h_res_raw", c_void_p),
                ("h_pre", c_void_p), ("h_post", c_void_p), ("h_res", c_void_p),
                ("h_pre_t", c_void_p), ("h_post_t", c_void_p), ("h_res_t", c_void_p),
                ("uv_history", c_void_p), ("convergence", c_void_p),
                ("D", c_int32), ("H", c_int32), ("Dp", c_int32), ("reserved", c_int32)]


class GemmArgs(ctypes.Structure):
    """struct hvs_gemm_args."""
    _fields_ = [("a0", c_void_p), ("lda0", c_int64), ("b0", c_void_p), ("ldb0", c_int64), ("K0", c_int),
                ("a1", c_void_p), ("lda1", c_int64), ("b1", c_void_p), ("ldb1", c_int64), ("K1", c_int),
                ("a_mn_major", c_int), ("b_mn_major", c_int), ("bias", c_void_p),
                ("ln_w", c_void_p), ("ln_b", c_void_p), ("ln_eps", c_float),
                ("aux", c_void_p), ("ld_aux", c_int64),
                ("out", c_void_p), ("out_dtype", c_int), ("ldo", c_int64),
                ("out2", c_void_p), ("ldo2", c_int64),
                ("M", c_int64), ("N", c_int), ("epilogue", c_int),
                ("dropout_p", c_float), ("dropout_seed", c_uint32),
                ("split_k", c_int), ("split_stride", c_int64), ("dropout_seed_dev", c_void_p),
                ("colsum_partials", c_void_p)]


class GradTensor(ctypes.Structure):
    """struct hvs_grad_tensor."""
    _fields_ = [("grad", c_void_p), ("numel", c_int64), ("group", c_int32), ("reserved", c_int32)]


class DecodeScale(ctypes.Structure):
    """struct hvs_decode_scale."""
    _fields_ = [("pred", c_void_p), ("pred_stride", c_int64 * 5), ("anchor_wh", c_void_p), ("boxes", c_void_p),
                ("class_scores", c_void_p), ("class_idx", c_void_p), ("objectness", c_void_p),
                ("A", c_int), ("H", c_int), ("W", c_int)]


class CoeffGrad(ctypes.Structure):
    """struct hvs_coeff_grad."""
    _fields_ = [("d_h_pre", c_void_p), ("d_h_post", c_void_p), ("d_h_res", c_void_p),
                ("d_h_pre_raw", c_void_p), ("d_h_post_raw", c_void_p), ("d_h_res_raw", c_void_p)]


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
    "hvs_rmsnorm_fwd": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int64, c_int, c_float, c_void_p]),
    "hvs_rmsnorm_bwd_workspace": (c_size_t, [c_int64, c_int]),
    "hvs_rmsnorm_bwd": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_float, c_void_p,
                                c_size_t, c_void_p]),
    "hvs_layernorm_fwd": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int64, c_int, c_int,
                                  c_int, c_float, c_void_p]),
    "hvs_layernorm_bwd_workspace": (c_size_t, [c_int64, c_int]),
    "hvs_layernorm_bwd": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_float,
                                  c_void_p, c_size_t, c_void_p]),
    "hvs_mhc_static_coeffs_workspace": (c_size_t, [POINTER(CoeffJob), c_int, c_int, c_int]),
    "hvs_mhc_static_coeffs": (c_int, [POINTER(CoeffJob), c_int, c_int, c_float, c_void_p, c_size_t, c_void_p]),
    "hvs_mhc_static_coeffs_bwd": (c_int, [POINTER(CoeffJob), POINTER(CoeffGrad), c_int, c_int, c_float, c_void_p, c_size_t,
                                          c_void_p]),
    "hvs_gemm_bf16": (c_int, [c_void_p, c_int64, c_void_p, c_int, c_void_p, c_int64, c_void_p, c_int, c_void_p, c_void_p,
                              c_void_p, c_float, c_void_p, c_int, c_int64, c_int64, c_int, c_int, c_void_p]),
    "hvs_gemm_bf16_ex": (c_int, [POINTER(GemmArgs), c_void_p]),
    "hvs_gemm_choose_split": (c_int, [c_int64, c_int, c_int64]),
    "hvs_reduce_partials": (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p]),
    "hvs_colsum_f32_workspace": (c_size_t, [c_int64, c_int]),
    "hvs_colsum_f32": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "hvs_colsum_bf16_workspace": (c_size_t, [c_int64, c_int]),
    "hvs_colsum_bf16": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "hvs_signal_ratio_workspace": (c_size_t, [c_int64]),
    "hvs_signal_ratio": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int64, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "hvs_profile_kernel_ms": (c_int, [POINTER(c_float)]),
    "hvs_head_decode_fused": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_int, c_int, c_int, c_int, c_void_p]),
    "hvs_mhc_module_fwd_supported": (c_int, [c_int, c_int]),
    "hvs_mhc_module_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_float, c_void_p, c_void_p, c_float, c_void_p, c_int, c_int64, c_int, c_int, c_void_p]),
    "hvs_grad_clip_dual_workspace": (c_size_t, [POINTER(GradTensor), c_int]),
    "hvs_grad_clip_dual": (c_int, [POINTER(GradTensor), c_int, c_float, c_float, c_void_p, c_void_p, c_size_t, c_void_p]),
    "hvs_gate_residual_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p]),
    "hvs_bias_act_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "hvs_se_gate_workspace": (c_size_t, [c_int64, c_int64, c_int]),
    "hvs_se_gate_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_int,
                                 c_void_p, c_size_t, c_void_p]),
    "hvs_preprocess_u8": (c_int, [c_void_p, c_int, c_int, c_int, c_int64, c_void_p, c_int, c_int, c_int, c_int,
                                  POINTER(c_float), POINTER(c_float), c_void_p]),
    "hvs_yolo_decode": (c_int, [c_void_p, c_int, POINTER(c_int64), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "hvs_yolo_decode_scales": (c_int, [POINTER(DecodeScale), c_int, c_int, c_int, c_int, c_void_p]),
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
            _build.have_nvcc()
        except RuntimeError as exc:
            # no compiler on this box: a prebuilt library that travelled with the tree is all there is
            if not os.path.exists(path):
                raise HvsError(f"libhvs_b200.so is not built and cannot be built here: {exc}") from exc
            import warnings
            warnings.warn("hvs_b200: nvcc not found, loading the prebuilt libhvs_b200.so without a freshness check")
        else:
            try:
                _build.build()                      # a compile error must surface, never fall back to a stale .so
            except Exception as exc:
                raise HvsError(f"building libhvs_b200.so failed: {exc}") from exc
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
