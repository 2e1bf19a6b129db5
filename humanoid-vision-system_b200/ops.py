"""Functional wrappers: torch CUDA tensors in, C-ABI calls on the current stream, tensors out.

PyTorch is used for device memory and streams only; every computation below runs in
libhvs_b200.so.  CPU tensors are rejected (no fallback).
"""
from __future__ import annotations

import ctypes
import functools
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import check


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _on_device(fn):
    """Run the wrapped op with the first tensor argument's device current, so the stream, the SM count and the
    per-device kernel attributes the library looks up are those of the device that owns the data."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        dev = None
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, torch.Tensor):
                dev = a.device
                break
            if isinstance(a, (list, tuple)) and a and isinstance(a[0], dict):
                for v in a[0].values():
                    if isinstance(v, torch.Tensor):
                        dev = v.device
                        break
                if dev is not None:
                    break
        if dev is None or dev.type != "cuda":
            return fn(*args, **kwargs)            # the op itself rejects CPU tensors
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapper


def _need_cuda(*tensors: Optional[torch.Tensor]):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.HvsError("hvs_b200 ops need CUDA tensors (there is no CPU path)")


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


# ----------------------------------------------------------------------------- K1 stream mHC
@_on_device
def mhc_stream_fwd(x: torch.Tensor, phi: torch.Tensor, bias: torch.Tensor, alpha: torch.Tensor,
                   scale: torch.Tensor, sk_iters: int = 20, eps_rms: float = 1e-8, eps_sk: float = 1e-8,
                   want_y: bool = True, want_u: bool = False, want_coeffs: bool = False,
                   split_phi: bool = False, out: Optional[torch.Tensor] = None,
                   saved: Optional[torch.Tensor] = None, adaptive: bool = False
                   ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor], Optional[torch.Tensor]]:
    """x [T,n,C] bf16 -> (y [T,n,C] bf16, u [T,C] bf16, coeffs [T,n*n+2n] fp32).
    `saved` ([T, SAVED_STRIDE] fp32, from `new_saved`) receives the statistics `mhc_stream_bwd_saved` consumes.
    `adaptive` (HVS_MHC_ADAPTIVE_ITERS, n = 4, C = 512): stop the Sinkhorn loop at a bitwise fixed point (same result)."""
    _need_cuda(x, phi, bias, alpha, scale)
    if x.dtype != torch.bfloat16 or x.dim() != 3 or not x.is_contiguous():
        raise _lib.HvsError("x must be a contiguous [T,n,C] bf16 tensor")
    t, n, c = x.shape
    k = n * n + 2 * n
    for name, p, shape in (("phi", phi, (n * c, k)), ("bias", bias, (k,)), ("alpha", alpha, (3,)), ("scale", scale, (n * c,))):
        if p.dtype != torch.float32 or tuple(p.shape) != shape or not p.is_contiguous():
            raise _lib.HvsError(f"{name} must be contiguous fp32 of shape {shape}")
    y = (out if out is not None else torch.empty_like(x)) if want_y else None
    u = torch.empty((t, c), dtype=torch.bfloat16, device=x.device) if want_u else None
    co = torch.empty((t, k), dtype=torch.float32, device=x.device) if want_coeffs else None
    flags = (_lib.HVS_MHC_SPLIT_PHI if split_phi else 0) | (_lib.HVS_MHC_ADAPTIVE_ITERS if adaptive else 0)
    if saved is not None:
        _need_cuda(saved)
        if saved.dtype != torch.float32 or tuple(saved.shape) != (t, _lib.HVS_MHC_SAVED_STRIDE) or not saved.is_contiguous():
            raise _lib.HvsError("saved must be contiguous [T, HVS_MHC_SAVED_STRIDE] fp32")
    check(_lib.load().hvs_mhc_stream_fwd_save(_ptr(x), _ptr(phi), _ptr(bias), _ptr(alpha), _ptr(scale), _ptr(y), _ptr(u),
                                              _ptr(co), _ptr(saved), t, n, c, sk_iters, eps_rms, eps_sk, flags, _stream()),
          "hvs_mhc_stream_fwd_save")
    return y, u, co


def new_saved(x: torch.Tensor) -> torch.Tensor:
    """Buffer for the forward's per-token statistics (112 B/token)."""
    return torch.empty((x.shape[0], _lib.HVS_MHC_SAVED_STRIDE), dtype=torch.float32, device=x.device)


@_on_device
def mhc_stream_bwd_saved(x: torch.Tensor, dy: torch.Tensor, saved: torch.Tensor, phi: torch.Tensor, bias: torch.Tensor,
                         alpha: torch.Tensor, scale: torch.Tensor, sk_iters: int = 20, eps_rms: float = 1e-8,
                         eps_sk: float = 1e-8, out: Optional[torch.Tensor] = None,
                         workspace: Optional[torch.Tensor] = None, adaptive: bool = False) -> Dict[str, torch.Tensor]:
    """Fused single-pass backward (dx and every parameter gradient) from the statistics saved by the forward.
    `adaptive` (HVS_MHC_ADAPTIVE_ITERS): replay / differentiate the Sinkhorn iterations only up to their bitwise fixed point."""
    _need_cuda(x, dy, saved, phi, bias, alpha, scale)
    if dy.dtype != torch.bfloat16 or dy.shape != x.shape or not dy.is_contiguous() or not x.is_contiguous():
        raise _lib.HvsError("dy must be contiguous bf16 with x's shape")
    t, n, c = x.shape
    if saved.dtype != torch.float32 or tuple(saved.shape) != (t, _lib.HVS_MHC_SAVED_STRIDE) or not saved.is_contiguous():
        raise _lib.HvsError("saved must be contiguous [T, HVS_MHC_SAVED_STRIDE] fp32")
    lib = _lib.load()
    dx = out if out is not None else torch.empty_like(x)
    dphi = torch.empty_like(phi)
    dbias = torch.empty_like(bias)
    dalpha = torch.empty_like(alpha)
    dscale = torch.empty_like(scale)
    ws_bytes = int(lib.hvs_mhc_stream_bwd_saved_workspace(t, n, c))
    ws = workspace if workspace is not None else torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=x.device)
    check(lib.hvs_mhc_stream_bwd_saved(_ptr(x), _ptr(dy), _ptr(saved), _ptr(phi), _ptr(bias), _ptr(alpha), _ptr(scale),
                                       _ptr(dx), _ptr(dphi), _ptr(dbias), _ptr(dalpha), _ptr(dscale), t, n, c, sk_iters,
                                       eps_rms, eps_sk, _lib.HVS_MHC_ADAPTIVE_ITERS if adaptive else 0, _ptr(ws), ws.numel(), _stream()),
          "hvs_mhc_stream_bwd_saved")
    return {"dx": dx, "dphi": dphi, "dbias": dbias, "dalpha": dalpha, "dscale": dscale}


@_on_device
def mhc_stream_post(x: torch.Tensor, coeffs: torch.Tensor, fu: torch.Tensor) -> torch.Tensor:
    _need_cuda(x, coeffs, fu)
    t, n, c = x.shape
    if fu.dtype != torch.bfloat16 or tuple(fu.shape) != (t, c) or not fu.is_contiguous():
        raise _lib.HvsError("fu must be contiguous [T,C] bf16")
    if coeffs.dtype != torch.float32 or tuple(coeffs.shape) != (t, n * n + 2 * n) or not coeffs.is_contiguous():
        raise _lib.HvsError("coeffs must be contiguous [T,n*n+2n] fp32")
    y = torch.empty_like(x)
    check(_lib.load().hvs_mhc_stream_post(_ptr(x), _ptr(coeffs), _ptr(fu), _ptr(y), t, n, c, _stream()),
          "hvs_mhc_stream_post")
    return y


@_on_device
def mhc_stream_bwd(x: torch.Tensor, dy: torch.Tensor, phi: torch.Tensor, bias: torch.Tensor, alpha: torch.Tensor,
                   scale: torch.Tensor, sk_iters: int = 20, eps_rms: float = 1e-8, eps_sk: float = 1e-8,
                   split_phi: bool = False) -> Dict[str, torch.Tensor]:
    """Returns dict(dx bf16 [T,n,C], dphi, dbias, dalpha, dscale fp32)."""
    _need_cuda(x, dy, phi, bias, alpha, scale)
    if dy.dtype != torch.bfloat16 or dy.shape != x.shape or not dy.is_contiguous() or not x.is_contiguous():
        raise _lib.HvsError("dy must be contiguous bf16 with x's shape")
    t, n, c = x.shape
    lib = _lib.load()
    dx = torch.empty_like(x)
    dphi = torch.empty_like(phi)
    dbias = torch.empty_like(bias)
    dalpha = torch.empty_like(alpha)
    dscale = torch.empty_like(scale)
    ws_bytes = int(lib.hvs_mhc_stream_bwd_workspace(t, n, c))
    ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=x.device)
    flags = _lib.HVS_MHC_SPLIT_PHI if split_phi else 0
    check(lib.hvs_mhc_stream_bwd(_ptr(x), _ptr(dy), _ptr(phi), _ptr(bias), _ptr(alpha), _ptr(scale), _ptr(dx),
                                 _ptr(dphi), _ptr(dbias), _ptr(dalpha), _ptr(dscale), t, n, c, sk_iters, eps_rms,
                                 eps_sk, flags, _ptr(ws), ws_bytes, _stream()), "hvs_mhc_stream_bwd")
    return {"dx": dx, "dphi": dphi, "dbias": dbias, "dalpha": dalpha, "dscale": dscale}


# ----------------------------------------------------------------------------- Sinkhorn / static coefficients
@_on_device
def sinkhorn(matrix: torch.Tensor, iters: int = 20, eps: float = 1e-8, tau: float = 1.0,
             history: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need_cuda(matrix, history)
    if matrix.dtype != torch.float32:
        raise _lib.HvsError("sinkhorn expects fp32")
    m = matrix.contiguous()
    if m.dim() == 2 and m.shape[0] == m.shape[1] and m.shape[0] > 32 and tau == 1.0:
        # one large square matrix (a layer's D x D H_res_raw): the multi-CTA slab kernel, not one CTA walking it
        out = torch.empty_like(m)
        static_coeffs([{"h_pre_raw": m.new_zeros((m.shape[0], 1)), "h_res_raw": m, "h_res": out, "convergence": history}], iters, eps)
        return out
    if m.dim() == 2:
        batch, n, mm = 1, m.shape[0], m.shape[1]
    else:
        n, mm = m.shape[-2], m.shape[-1]
        batch = m.numel() // max(n * mm, 1)
    out = torch.empty_like(m)
    check(_lib.load().hvs_sinkhorn(_ptr(m), _ptr(out), batch, n, mm, iters, eps, tau, _ptr(history), _stream()),
          "hvs_sinkhorn")
    return out


@_on_device
def constrained_matrices(h_pre_raw: torch.Tensor, h_post_raw: torch.Tensor, h_res_raw: torch.Tensor,
                         iters: int = 20, eps: float = 1e-8, history: Optional[torch.Tensor] = None):
    _need_cuda(h_pre_raw, h_post_raw, h_res_raw)
    d, hidden = h_pre_raw.shape
    h_pre = torch.empty_like(h_pre_raw)
    h_post = torch.empty_like(h_post_raw)
    h_res = torch.empty_like(h_res_raw)
    check(_lib.load().hvs_mhc_constrained_matrices(
        _ptr(h_pre_raw.contiguous()), _ptr(h_post_raw.contiguous()), _ptr(h_res_raw.contiguous()), _ptr(h_pre),
        _ptr(h_post), _ptr(h_res), d, hidden, iters, eps, _ptr(history), _stream()), "hvs_mhc_constrained_matrices")
    return h_pre, h_post, h_res


# ----------------------------------------------------------------------------- row norms
_NORM_DTYPES = {torch.float32: _lib.HVS_DTYPE_F32, torch.bfloat16: _lib.HVS_DTYPE_BF16}


@_on_device
def rmsnorm_fwd(x: torch.Tensor, scale: torch.Tensor, eps: float = 1e-8, out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """RMSNorm over the last axis (manifold_layers.py:449-456); x fp32 / bf16, any leading shape."""
    _need_cuda(x, scale)
    if x.dtype not in _NORM_DTYPES or scale.dtype != torch.float32:
        raise _lib.HvsError("rmsnorm expects fp32 / bf16 data and an fp32 scale")
    dim = x.shape[-1]
    xc = x.contiguous()
    out = torch.empty(x.shape, dtype=out_dtype or x.dtype, device=x.device)
    check(_lib.load().hvs_rmsnorm_fwd(_ptr(xc), _NORM_DTYPES[xc.dtype], _ptr(scale.contiguous()), _ptr(out),
                                      _NORM_DTYPES[out.dtype], xc.numel() // max(dim, 1), dim, eps, _stream()), "hvs_rmsnorm_fwd")
    return out


@_on_device
def rmsnorm_bwd(x: torch.Tensor, scale: torch.Tensor, dy: torch.Tensor, eps: float = 1e-8) -> Tuple[torch.Tensor, torch.Tensor]:
    _need_cuda(x, scale, dy)
    if x.dtype not in _NORM_DTYPES or dy.dtype != x.dtype:
        raise _lib.HvsError("rmsnorm_bwd expects x and dy of the same fp32 / bf16 dtype")
    dim = x.shape[-1]
    rows = x.numel() // max(dim, 1)
    xc, dyc = x.contiguous(), dy.contiguous()
    dx = torch.empty_like(xc)
    dscale = torch.empty(dim, dtype=torch.float32, device=x.device)
    lib = _lib.load()
    nb = int(lib.hvs_rmsnorm_bwd_workspace(rows, dim))
    ws = torch.empty(max(nb, 256), dtype=torch.uint8, device=x.device)
    check(lib.hvs_rmsnorm_bwd(_ptr(xc), _NORM_DTYPES[xc.dtype], _ptr(scale.contiguous()), _ptr(dyc), _ptr(dx), _ptr(dscale),
                              rows, dim, eps, _ptr(ws), ws.numel(), _stream()), "hvs_rmsnorm_bwd")
    return dx, dscale


@_on_device
def layernorm_fwd(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-5,
                  out_dtype: torch.dtype = torch.float32, out_ld: Optional[int] = None, want_copy: bool = False,
                  copy_ld: Optional[int] = None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """LayerNorm over the last axis of x [rows, dim] -> out [rows, out_ld] (pad columns zero); optionally bf16(x)."""
    _need_cuda(x, weight, bias)
    rows, dim = x.shape
    out_ld = out_ld or dim
    copy_ld = copy_ld or out_ld
    out = torch.empty((rows, out_ld), dtype=out_dtype, device=x.device)
    cp = torch.empty((rows, copy_ld), dtype=torch.bfloat16, device=x.device) if want_copy else None
    xc = x.contiguous()
    check(_lib.load().hvs_layernorm_fwd(_ptr(xc), _NORM_DTYPES[xc.dtype], _ptr(weight), _ptr(bias), _ptr(out),
                                        _NORM_DTYPES[out_dtype], _ptr(cp), rows, dim, out_ld, copy_ld, eps, _stream()),
          "hvs_layernorm_fwd")
    return out, cp


LN_BWD_DIMS = (32, 64, 128, 256, 512, 1024, 1792)     # tuned small-row kernel up to 512, the wide-row kernels beyond (any dim % 8 == 0)


@_on_device
def layernorm_bwd(x: torch.Tensor, weight: torch.Tensor, dy: torch.Tensor, eps: float = 1e-5):
    """Backward of LayerNorm over the last axis of x [rows, dim] (dim in LN_BWD_DIMS): (dx [x's dtype], dweight, dbias fp32)."""
    _need_cuda(x, weight, dy)
    rows, dim = x.shape
    xc, dyc = x.contiguous(), dy.contiguous()
    dx = torch.empty_like(xc)
    dw = torch.empty(dim, dtype=torch.float32, device=x.device)
    db = torch.empty(dim, dtype=torch.float32, device=x.device)
    lib = _lib.load()
    nb = int(lib.hvs_layernorm_bwd_workspace(rows, dim))
    ws = torch.empty(max(nb, 256), dtype=torch.uint8, device=x.device)
    check(lib.hvs_layernorm_bwd(_ptr(xc), _NORM_DTYPES[xc.dtype], _ptr(weight), _ptr(dyc), _NORM_DTYPES[dyc.dtype], _ptr(dx), _ptr(dw),
                                _ptr(db), rows, dim, eps, _ptr(ws), ws.numel(), _stream()), "hvs_layernorm_bwd")
    return dx, dw, db


# ----------------------------------------------------------------------------- K2: batched static coefficients, GEMMs
def _coeff_arrays(jobs: Sequence[Dict[str, Optional[torch.Tensor]]]):
    arr = (_lib.CoeffJob * len(jobs))()
    for a, j in zip(arr, jobs):
        for k in ("h_pre_raw", "h_post_raw", "h_res_raw", "h_pre", "h_post", "h_res", "h_pre_t", "h_post_t", "h_res_t",
                  "uv_history", "convergence"):
            setattr(a, k, _ptr(j.get(k)))
        a.D, a.H = j["h_res_raw"].shape[0], j["h_pre_raw"].shape[1]
        a.Dp = j["h_res_t"].shape[1] if j.get("h_res_t") is not None else (j["h_pre_t"].shape[1] if j.get("h_pre_t") is not None else a.D)
    return arr


@_on_device
def static_coeffs(jobs: Sequence[Dict[str, Optional[torch.Tensor]]], iters: int = 20, eps: float = 1e-8) -> None:
    """constrained_matrices of MANY layers in one launch.  Each job is a dict of tensors named after
    struct hvs_coeff_job's fields; outputs are written in place."""
    if not jobs:
        return
    for j in jobs:
        _need_cuda(*[t for t in j.values() if isinstance(t, torch.Tensor)])
    lib = _lib.load()
    arr = _coeff_arrays(jobs)
    dev = jobs[0]["h_res_raw"].device
    nb = int(lib.hvs_mhc_static_coeffs_workspace(arr, len(jobs), iters, 0))
    if nb == 0:
        raise _lib.HvsError("hvs_mhc_static_coeffs: unsupported job list")
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    check(lib.hvs_mhc_static_coeffs(arr, len(jobs), iters, eps, _ptr(ws), nb, _stream()), "hvs_mhc_static_coeffs")


@_on_device
def static_coeffs_bwd(jobs: Sequence[Dict[str, Optional[torch.Tensor]]], grads: Sequence[Dict[str, Optional[torch.Tensor]]],
                      iters: int = 20, eps: float = 1e-8) -> None:
    """Backward of static_coeffs: grads[i] has d_h_pre / d_h_post / d_h_res (inputs, may be None) and
    d_h_pre_raw / d_h_post_raw / d_h_res_raw (outputs)."""
    if not jobs:
        return
    lib = _lib.load()
    arr = _coeff_arrays(jobs)
    garr = (_lib.CoeffGrad * len(jobs))()
    for g, d in zip(garr, grads):
        for k in ("d_h_pre", "d_h_post", "d_h_res", "d_h_pre_raw", "d_h_post_raw", "d_h_res_raw"):
            t = d.get(k)
            if t is not None:
                _need_cuda(t)
                if t.dtype != torch.float32 or not t.is_contiguous():
                    raise _lib.HvsError(f"{k} must be contiguous fp32")
            setattr(g, k, _ptr(t))
    dev = jobs[0]["h_res_raw"].device
    nb = int(lib.hvs_mhc_static_coeffs_workspace(arr, len(jobs), iters, 1))
    if nb == 0:
        raise _lib.HvsError("hvs_mhc_static_coeffs_bwd: unsupported job list")
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    check(lib.hvs_mhc_static_coeffs_bwd(arr, garr, len(jobs), iters, eps, _ptr(ws), nb, _stream()), "hvs_mhc_static_coeffs_bwd")


@_on_device
def gemm_bf16(a0: torch.Tensor, b0: torch.Tensor, a1: Optional[torch.Tensor] = None, b1: Optional[torch.Tensor] = None,
              bias: Optional[torch.Tensor] = None, ln_weight: Optional[torch.Tensor] = None,
              ln_bias: Optional[torch.Tensor] = None, ln_eps: float = 1e-5, epilogue: int = _lib.HVS_GEMM_EPI_NONE,
              out_dtype: torch.dtype = torch.bfloat16, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[M,N] = epilogue(a0 @ b0^T (+ a1 @ b1^T)); a*: [M,K*] bf16 (row stride may exceed K*), b*: [N,K*] bf16."""
    _need_cuda(a0, b0, a1, b1, bias, ln_weight, ln_bias, out)
    for t in (a0, b0, a1, b1):
        if t is not None and (t.dtype != torch.bfloat16 or t.dim() != 2 or t.stride(1) != 1):
            raise _lib.HvsError("gemm operands must be 2-D bf16 with unit inner stride")
    m, k0 = a0.shape
    n = b0.shape[0]
    if b0.shape[1] != k0 or not b0.is_contiguous():
        raise _lib.HvsError("b0 must be contiguous [N, K0]")
    k1 = 0
    if a1 is not None:
        k1 = a1.shape[1]
        if b1 is None or tuple(b1.shape) != (n, k1) or not b1.is_contiguous() or a1.shape[0] != m:
            raise _lib.HvsError("second operand pair must be a1 [M,K1], b1 [N,K1]")
    if out is None:
        out = torch.empty((m, n), dtype=out_dtype, device=a0.device)
    check(_lib.load().hvs_gemm_bf16(_ptr(a0), a0.stride(0), _ptr(b0), k0, _ptr(a1), a1.stride(0) if a1 is not None else 0,
                                    _ptr(b1), k1, _ptr(bias), _ptr(ln_weight), _ptr(ln_bias), ln_eps, _ptr(out),
                                    _NORM_DTYPES[out.dtype], out.stride(0), m, n, epilogue, _stream()), "hvs_gemm_bf16")
    return out


def _check_bf16_2d(*tensors):
    for t in tensors:
        if t is not None and (t.dtype != torch.bfloat16 or t.dim() != 2 or t.stride(1) != 1):
            raise _lib.HvsError("gemm operands must be 2-D bf16 with unit inner stride")


@_on_device
def gemm_bf16_ex(a: torch.Tensor, b: torch.Tensor, a_mn: bool = False, b_mn: bool = False, bias: Optional[torch.Tensor] = None,
                 epilogue: int = _lib.HVS_GEMM_EPI_NONE, aux: Optional[torch.Tensor] = None, out_dtype: torch.dtype = torch.bfloat16,
                 dropout_p: float = 0.0, dropout_seed: int = 0, dropout_seed_dev: Optional[torch.Tensor] = None,
                 want_colsum: bool = False):
    """The training form of the K2 GEMM kernel (hvs_gemm_bf16_ex):  out[M, N] = epilogue(op(a) op(b)^T).
    a_mn / b_mn: the operand is given as its transpose in place ([K, M] / [K, N] row-major).
    HVS_GEMM_EPI_BIAS_GELU_SAVE returns (out, z); HVS_GEMM_EPI_DGELU with want_colsum returns (out, column sums of out [N] fp32:
    the bias gradient, summed in the epilogue per 32-row group and finished by hvs_colsum_f32); else out."""
    _need_cuda(a, b, bias, aux)
    _check_bf16_2d(a, b, aux)
    k, m = (a.shape[0], a.shape[1]) if a_mn else (a.shape[1], a.shape[0])
    kb, n = (b.shape[0], b.shape[1]) if b_mn else (b.shape[1], b.shape[0])
    if k != kb:
        raise _lib.HvsError(f"contraction extents differ: {k} vs {kb}")
    out = torch.empty((m, n), dtype=out_dtype, device=a.device)
    z = torch.empty((m, n), dtype=torch.bfloat16, device=a.device) if epilogue == _lib.HVS_GEMM_EPI_BIAS_GELU_SAVE else None
    g = _lib.GemmArgs()
    g.a0, g.lda0, g.b0, g.ldb0, g.K0 = _ptr(a), a.stride(0), _ptr(b), b.stride(0), k
    g.a_mn_major, g.b_mn_major = int(a_mn), int(b_mn)
    g.bias = _ptr(bias)
    if aux is not None:
        g.aux, g.ld_aux = _ptr(aux), aux.stride(0)
    g.out, g.out_dtype, g.ldo = _ptr(out), _NORM_DTYPES[out_dtype], out.stride(0)
    if z is not None:
        g.out2, g.ldo2 = _ptr(z), z.stride(0)
    g.M, g.N, g.epilogue = m, n, epilogue
    g.dropout_p, g.dropout_seed = float(dropout_p), int(dropout_seed) & 0xFFFFFFFF
    if dropout_seed_dev is not None:
        _need_cuda(dropout_seed_dev)
        if dropout_seed_dev.dtype not in (torch.int32, torch.int64) or dropout_seed_dev.numel() < 1:
            raise _lib.HvsError("dropout_seed_dev must be an int32 / int64 device tensor")
        g.dropout_seed_dev = _ptr(dropout_seed_dev)
    g.split_k = 1
    part = None
    if want_colsum:
        if epilogue != _lib.HVS_GEMM_EPI_DGELU:
            raise _lib.HvsError("want_colsum goes with HVS_GEMM_EPI_DGELU")
        part = torch.empty((4 * ((m + 127) // 128), n), dtype=torch.float32, device=a.device)
        g.colsum_partials = _ptr(part)
    lib = _lib.load()
    check(lib.hvs_gemm_bf16_ex(ctypes.byref(g), _stream()), "hvs_gemm_bf16_ex")
    if part is not None:
        db = torch.empty(n, dtype=torch.float32, device=a.device)
        nb = int(lib.hvs_colsum_f32_workspace(part.shape[0], n))
        ws = torch.empty(max(nb, 256), dtype=torch.uint8, device=a.device)
        check(lib.hvs_colsum_f32(_ptr(part), part.shape[0], n, _ptr(db), _ptr(ws), ws.numel(), _stream()), "hvs_colsum_f32")
        return out, db
    return (out, z) if z is not None else out


@_on_device
def gemm_wgrad(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Weight gradient  out[Fa, Fb] = a^T b  over the tokens: a [T, Fa], b [T, Fb] bf16 as they lie in memory (both operands
    MN-major), fp32 result.  Tiny outputs over millions of tokens are cut along T (split-K) into fp32 partials that
    hvs_reduce_partials sums in a fixed order."""
    _need_cuda(a, b, out)
    _check_bf16_2d(a, b)
    t, fa = a.shape
    if b.shape[0] != t:
        raise _lib.HvsError("wgrad operands must share the token axis")
    fb = b.shape[1]
    lib = _lib.load()
    splits = int(lib.hvs_gemm_choose_split(fa, fb, t))
    if out is None:
        out = torch.empty((fa, fb), dtype=torch.float32, device=a.device)
    dst = out if splits == 1 else torch.empty((splits, fa, fb), dtype=torch.float32, device=a.device)
    g = _lib.GemmArgs()
    g.a0, g.lda0, g.b0, g.ldb0, g.K0 = _ptr(a), a.stride(0), _ptr(b), b.stride(0), t
    g.a_mn_major, g.b_mn_major = 1, 1
    g.out, g.out_dtype, g.ldo = _ptr(dst), _lib.HVS_DTYPE_F32, fb
    g.M, g.N, g.epilogue = fa, fb, _lib.HVS_GEMM_EPI_NONE
    g.split_k, g.split_stride = splits, fa * fb
    check(lib.hvs_gemm_bf16_ex(ctypes.byref(g), _stream()), "hvs_gemm_bf16_ex(wgrad)")
    if splits > 1:
        check(lib.hvs_reduce_partials(_ptr(dst), splits, fa * fb, fa * fb, _ptr(out), _stream()), "hvs_reduce_partials")
    return out


@_on_device
def colsum_bf16(x: torch.Tensor) -> torch.Tensor:
    """Column sums of a bf16 [rows, cols] matrix in fp32 (bias gradients), fixed summation order."""
    _need_cuda(x)
    _check_bf16_2d(x)
    rows, cols = x.shape
    lib = _lib.load()
    out = torch.empty(cols, dtype=torch.float32, device=x.device)
    nb = int(lib.hvs_colsum_bf16_workspace(rows, cols))
    ws = torch.empty(max(nb, 256), dtype=torch.uint8, device=x.device)
    check(lib.hvs_colsum_bf16(_ptr(x), x.stride(0), rows, cols, _ptr(out), _ptr(ws), ws.numel(), _stream()), "hvs_colsum_bf16")
    return out


@_on_device
def signal_ratio(out: torch.Tensor, x: torch.Tensor, dst: torch.Tensor) -> None:
    """dst[0] = mean row norm of out / (mean row norm of x + 1e-8)  (manifold_layers.py:295-303); dst: a 1-element fp32 view
    on the device (e.g. signal_ratio_history[i:i+1])."""
    _need_cuda(out, x, dst)
    if out.shape != x.shape or out.dim() != 2 or not out.is_contiguous() or not x.is_contiguous():
        raise _lib.HvsError("signal_ratio expects two contiguous [rows, dim] tensors of the same shape")
    if dst.dtype != torch.float32 or dst.numel() < 1:
        raise _lib.HvsError("dst must be an fp32 device tensor")
    rows, dim = out.shape
    if rows == 0:
        return
    lib = _lib.load()
    nb = int(lib.hvs_signal_ratio_workspace(rows))
    ws = torch.empty(max(nb, 256), dtype=torch.uint8, device=out.device)
    check(lib.hvs_signal_ratio(_ptr(out), _NORM_DTYPES[out.dtype], _ptr(x), _NORM_DTYPES[x.dtype], rows, dim, _ptr(dst), _ptr(ws),
                               ws.numel(), _stream()), "hvs_signal_ratio")


def mhc_module_fwd_supported(d: int, h: int) -> bool:
    return bool(_lib.load().hvs_mhc_module_fwd_supported(d, h))


@_on_device
def mhc_module_fwd(x: torch.Tensor, h_pre_t: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor,
                   h_post_t: torch.Tensor, h_res_t: torch.Tensor, ln_pre: Tuple[torch.Tensor, torch.Tensor, float],
                   ln_post: Tuple[torch.Tensor, torch.Tensor, float], out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """The whole token path of the module (:248-267) in one kernel; x [T, D] bf16, (D, H) in {(32, 128), (64, 256)}."""
    _need_cuda(x, h_pre_t, w1, b1, w2, b2, h_post_t, h_res_t)
    t, d = x.shape
    h = h_pre_t.shape[0]
    if x.dtype != torch.bfloat16 or not x.is_contiguous():
        raise _lib.HvsError("mhc_module_fwd expects contiguous bf16 x")
    for name, ten, shape in (("h_pre_t", h_pre_t, (h, d)), ("w1", w1, (2 * h, h)), ("w2", w2, (h, 2 * h)), ("h_post_t", h_post_t, (d, h)),
                             ("h_res_t", h_res_t, (d, d))):
        if ten.dtype != torch.bfloat16 or tuple(ten.shape) != shape or not ten.is_contiguous():
            raise _lib.HvsError(f"{name} must be contiguous bf16 {shape}")
    out = torch.empty((t, d), dtype=out_dtype, device=x.device)
    check(_lib.load().hvs_mhc_module_fwd(_ptr(x), _ptr(h_pre_t), _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2), _ptr(h_post_t), _ptr(h_res_t),
                                         _ptr(ln_pre[0]), _ptr(ln_pre[1]), ln_pre[2], _ptr(ln_post[0]), _ptr(ln_post[1]), ln_post[2],
                                         _ptr(out), _NORM_DTYPES[out_dtype], t, d, h, _stream()), "hvs_mhc_module_fwd")
    return out


# ----------------------------------------------------------------------------- edges of the path: grad clipping, frame preprocessing
def is_mhc_parameter(name: str) -> bool:
    """The reference's grouping rule (mhc_trainer.py:357-359)."""
    return "mhc" in name.lower() or "H_" in name


def clip_grad_dual(named_parameters, max_grad_norm: float = 1.0, mhc_max_norm: float = 0.5) -> torch.Tensor:
    """_apply_manifold_gradient_clipping (mhc_trainer.py:342-383) without host synchronisation: clips the mHC
    parameters' gradients to mhc_max_norm and the others to max_grad_norm IN PLACE; returns a device tensor
    [norm_mhc, norm_other, coef_mhc, coef_other] (total norm = hypot of the first two)."""
    items = [(n, p) for n, p in named_parameters if p.grad is not None]
    if not items:
        return torch.zeros(4)
    dev = items[0][1].grad.device
    with torch.cuda.device(dev):
        arr = (_lib.GradTensor * len(items))()
        for a, (n, p) in zip(arr, items):
            g = p.grad
            _need_cuda(g)
            if g.dtype != torch.float32 or not g.is_contiguous():
                raise _lib.HvsError(f"gradient of {n} must be contiguous fp32")
            a.grad, a.numel, a.group = g.data_ptr(), g.numel(), 0 if is_mhc_parameter(n) else 1
        lib = _lib.load()
        nb = int(lib.hvs_grad_clip_dual_workspace(arr, len(items)))
        ws = torch.empty(max(nb, 256), dtype=torch.uint8, device=dev)
        res = torch.empty(4, dtype=torch.float32, device=dev)
        check(lib.hvs_grad_clip_dual(arr, len(items), mhc_max_norm, max_grad_norm, _ptr(res), _ptr(ws), ws.numel(), _stream()),
              "hvs_grad_clip_dual")
    return res


@_on_device
def gate_residual(y: torch.Tensor, gate: torch.Tensor, residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y * gate (+ residual) for channels-last bf16 feature maps [B, C, H, W] and a per-(image, channel) gate [B, C, 1, 1]
    (vision_backbone.py:125-133) in one pass; returns a channels-last tensor."""
    _need_cuda(y, gate, residual)
    b, c, h, w = y.shape
    cl = torch.channels_last
    if y.dtype != torch.bfloat16 or not y.is_contiguous(memory_format=cl) or gate.dtype != torch.bfloat16 or gate.numel() != b * c:
        raise _lib.HvsError("gate_residual expects a channels-last bf16 map and a bf16 [B, C, 1, 1] gate")
    if residual is not None and (residual.dtype != torch.bfloat16 or residual.shape != y.shape or not residual.is_contiguous(memory_format=cl)):
        raise _lib.HvsError("residual must be a channels-last bf16 map of y's shape")
    out = torch.empty_like(y, memory_format=cl)
    g = gate.reshape(b, c).contiguous()
    check(_lib.load().hvs_gate_residual_bf16(_ptr(y), _ptr(g), _ptr(residual), _ptr(out), b, h * w, c, _stream()), "hvs_gate_residual_bf16")
    return out


ACTIVATIONS = {"none": 0, "silu": 1, "relu": 2, "leaky_relu_0.1": 3}


@_on_device
def bias_act(y: torch.Tensor, bias: torch.Tensor, activation: str = "silu") -> torch.Tensor:
    """act(y + bias[c]) for a channels-last bf16 map [B, C, H, W] in place (the bias of a BatchNorm-folded convolution and
    the activation after it in one pass)."""
    _need_cuda(y, bias)
    b, c, h, w = y.shape
    if y.dtype != torch.bfloat16 or not y.is_contiguous(memory_format=torch.channels_last) or bias.dtype != torch.float32 or bias.numel() != c:
        raise _lib.HvsError("bias_act expects a channels-last bf16 map and an fp32 bias [C]")
    check(_lib.load().hvs_bias_act_bf16(_ptr(y), _ptr(bias.contiguous()), _ptr(y), b * h * w, c, ACTIVATIONS[activation], _stream()),
          "hvs_bias_act_bf16")
    return y


@_on_device
def se_gate(y: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor, activation: str = "silu",
            workspace: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """sigmoid(W2 act(W1 mean_hw(y) + b1) + b2) for a channels-last bf16 map [B, C, H, W] in one launch (the squeeze-excite
    gate of ConvMHCLayer, vision_backbone.py:77-83).  w1 [hidden, C(,1,1)], w2 [C, hidden(,1,1)]; parameters are used as bf16
    (the autocast path).  Returns (gate [B, C, 1, 1] bf16, workspace): pass the workspace back in on the next call of the
    same shape -- it is created zeroed and the kernel leaves it zeroed."""
    _need_cuda(y, w1, b1, w2, b2)
    b, c, h, w = y.shape
    hidden = w1.shape[0]
    if (y.dtype != torch.bfloat16 or not y.is_contiguous(memory_format=torch.channels_last) or w1.numel() != hidden * c
            or w2.numel() != c * hidden or b1.numel() != hidden or b2.numel() != c or activation not in ACTIVATIONS):
        raise _lib.HvsError("se_gate expects a channels-last bf16 map, w1 [hidden, C], b1 [hidden], w2 [C, hidden], b2 [C]")
    bf = torch.bfloat16
    w1, b1, w2, b2 = (t.detach().to(bf).contiguous() for t in (w1, b1, w2, b2))
    lib = _lib.load()
    need = int(lib.hvs_se_gate_workspace(b, h * w, c))
    if need == 0:
        raise _lib.HvsError(f"se_gate: unsupported shape (C = {c})")
    key = (b, h * w, c, y.device)
    if workspace is None or getattr(workspace, "_hvs_se_key", None) != key or workspace.numel() != need:
        workspace = torch.zeros(need, dtype=torch.uint8, device=y.device)     # (the layout inside depends on the shape)
        workspace._hvs_se_key = key
    gate = torch.empty((b, c, 1, 1), dtype=bf, device=y.device)
    check(lib.hvs_se_gate_bf16(_ptr(y), _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2), _ptr(gate), b, h * w, c, hidden, ACTIVATIONS[activation],
                               _ptr(workspace), workspace.numel(), _stream()), "hvs_se_gate_bf16")
    return gate, workspace


_PRE_DTYPES = {torch.float32: _lib.HVS_DTYPE_F32, torch.float16: _lib.HVS_DTYPE_F16, torch.bfloat16: _lib.HVS_DTYPE_BF16}
IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


@_on_device
def preprocess_frame(frame_u8: torch.Tensor, height: int, width: int, bgr_to_rgb: bool = True, mean=IMAGENET_MEAN, std=IMAGENET_STD,
                     out_dtype: torch.dtype = torch.float32, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """HWC (or HW) uint8 frame on the device -> normalised CHW tensor [3, height, width] (preprocessing.py:181-273)."""
    _need_cuda(frame_u8, out)
    if frame_u8.dtype != torch.uint8 or frame_u8.dim() not in (2, 3) or not frame_u8.is_contiguous():
        raise _lib.HvsError("frame must be a contiguous uint8 HW or HWC tensor")
    h, w = frame_u8.shape[:2]
    c = 1 if frame_u8.dim() == 2 else frame_u8.shape[2]
    if out is None:
        out = torch.empty((3, height, width), dtype=out_dtype, device=frame_u8.device)
    m3 = (ctypes.c_float * 3)(*mean) if mean is not None else None
    s3 = (ctypes.c_float * 3)(*std) if std is not None else None
    check(_lib.load().hvs_preprocess_u8(_ptr(frame_u8), h, w, c, w * c, _ptr(out), _PRE_DTYPES[out.dtype], height, width,
                                        1 if bgr_to_rgb else 0, m3, s3, _stream()), "hvs_preprocess_u8")
    return out


# ----------------------------------------------------------------------------- detection
_DTYPES = {torch.float32: _lib.HVS_DTYPE_F32, torch.float16: _lib.HVS_DTYPE_F16, torch.bfloat16: _lib.HVS_DTYPE_BF16}


@_on_device
def yolo_decode(pred: torch.Tensor, anchor_wh: torch.Tensor, want_scores: bool = False,
                want_objectness: bool = True) -> Dict[str, torch.Tensor]:
    """pred [B,A,H,W,5+C] (any strides, fp32/fp16/bf16), anchor_wh [A,2] fp32."""
    _need_cuda(pred, anchor_wh)
    if pred.dim() != 5 or pred.dtype not in _DTYPES:
        raise _lib.HvsError("pred must be [B,A,H,W,5+C] fp32/fp16/bf16")
    b, a, h, w, d = pred.shape
    c = d - 5
    dev = pred.device
    boxes = torch.empty((b, a, h, w, 4), dtype=torch.float32, device=dev)
    cs = torch.empty((b, a, h, w), dtype=torch.float32, device=dev)
    ci = torch.empty((b, a, h, w), dtype=torch.int64, device=dev)
    obj = torch.empty((b, a, h, w, 1), dtype=torch.float32, device=dev) if want_objectness else None
    sc = torch.empty((b, a, h, w, c), dtype=torch.float32, device=dev) if want_scores else None
    strides = (ctypes.c_int64 * 5)(*pred.stride())
    awh = anchor_wh.to(device=dev, dtype=torch.float32).contiguous()
    check(_lib.load().hvs_yolo_decode(_ptr(pred), _DTYPES[pred.dtype], strides, _ptr(awh), _ptr(boxes), _ptr(cs),
                                      _ptr(ci), _ptr(obj), _ptr(sc), b, a, h, w, c, _stream()), "hvs_yolo_decode")
    out = {"boxes": boxes, "class_scores": cs, "class_indices": ci}
    if obj is not None:
        out["objectness"] = obj
    if sc is not None:
        out["scores"] = sc
    return out


def yolo_decode_scales(preds: Sequence[torch.Tensor], anchor_whs: Sequence[torch.Tensor],
                       want_objectness: bool = True) -> List[Dict[str, torch.Tensor]]:
    """Every scale of the head in one launch (hvs_yolo_decode_scales): preds[k] [B,A_k,H_k,W_k,5+C] (any strides, one dtype,
    one B and C), anchor_whs[k] [A_k,2].  Outputs per scale as yolo_decode without the per-class score tensor."""
    if len(preds) == 0 or len(preds) != len(anchor_whs):
        raise _lib.HvsError("yolo_decode_scales: one anchor table per prediction tensor")
    _need_cuda(*preds, *anchor_whs)
    p0 = preds[0]
    if any(p.dim() != 5 or p.dtype != p0.dtype or p.shape[0] != p0.shape[0] or p.shape[4] != p0.shape[4] or p.device != p0.device
           for p in preds) or p0.dtype not in _DTYPES:
        raise _lib.HvsError("yolo_decode_scales: preds must be [B,A,H,W,5+C] fp32/fp16/bf16 sharing B, C, dtype and device")
    dev = p0.device
    with torch.cuda.device(dev):
        table = (_lib.DecodeScale * len(preds))()
        outs, keep = [], []
        for k, (p, awh) in enumerate(zip(preds, anchor_whs)):
            b, a, h, w, d = p.shape
            awh = awh.to(device=dev, dtype=torch.float32).contiguous()
            boxes = torch.empty((b, a, h, w, 4), dtype=torch.float32, device=dev)
            cs = torch.empty((b, a, h, w), dtype=torch.float32, device=dev)
            ci = torch.empty((b, a, h, w), dtype=torch.int64, device=dev)
            obj = torch.empty((b, a, h, w, 1), dtype=torch.float32, device=dev) if want_objectness else None
            e = table[k]
            e.pred, e.anchor_wh, e.boxes, e.class_scores, e.class_idx, e.objectness = _ptr(p), _ptr(awh), _ptr(boxes), _ptr(cs), _ptr(ci), _ptr(obj)
            for i, st in enumerate(p.stride()):
                e.pred_stride[i] = st
            e.A, e.H, e.W = a, h, w
            keep.append(awh)
            out = {"boxes": boxes, "class_scores": cs, "class_indices": ci}
            if obj is not None:
                out["objectness"] = obj
            outs.append(out)
        check(_lib.load().hvs_yolo_decode_scales(table, len(preds), _DTYPES[p0.dtype], p0.shape[0], p0.shape[4] - 5, _stream()),
              "hvs_yolo_decode_scales")
    return outs


@_on_device
def head_decode_fused(tokens: torch.Tensor, weight256: torch.Tensor, bias256: torch.Tensor, anchor_wh: torch.Tensor, b: int, h: int,
                      w: int, want_objectness: bool = False) -> Dict[str, torch.Tensor]:
    """1x1 prediction conv + decode in one kernel.  tokens [B*H*W, C_in] bf16 (pixel-major), weight256 [256, C_in] bf16."""
    _need_cuda(tokens, weight256, bias256, anchor_wh)
    if tokens.dtype != torch.bfloat16 or tokens.dim() != 2 or tokens.stride(1) != 1 or tokens.shape[0] != b * h * w:
        raise _lib.HvsError("tokens must be [B*H*W, C_in] bf16 with unit inner stride")
    c_in = tokens.shape[1]
    if weight256.dtype != torch.bfloat16 or tuple(weight256.shape) != (256, c_in) or not weight256.is_contiguous():
        raise _lib.HvsError("weight256 must be contiguous bf16 [256, C_in]")
    dev = tokens.device
    boxes = torch.empty((b, 3, h, w, 4), dtype=torch.float32, device=dev)
    cs = torch.empty((b, 3, h, w), dtype=torch.float32, device=dev)
    ci = torch.empty((b, 3, h, w), dtype=torch.int64, device=dev)
    obj = torch.empty((b, 3, h, w, 1), dtype=torch.float32, device=dev) if want_objectness else None
    awh = anchor_wh.to(device=dev, dtype=torch.float32).contiguous()
    check(_lib.load().hvs_head_decode_fused(_ptr(tokens), tokens.stride(0), _ptr(weight256), _ptr(bias256.contiguous()), _ptr(awh),
                                            _ptr(boxes), _ptr(cs), _ptr(ci), _ptr(obj), b, h, w, c_in, _stream()), "hvs_head_decode_fused")
    out = {"boxes": boxes, "class_scores": cs, "class_indices": ci}
    if obj is not None:
        out["objectness"] = obj
    return out


@_on_device
def nms(boxes: torch.Tensor, scores: torch.Tensor, classes: Optional[torch.Tensor] = None,
        iou_threshold: float = 0.5, max_detections: int = 100, score_threshold: float = float("-inf"),
        class_aware: bool = False, boxes_xyxy: bool = True, offsets: Optional[torch.Tensor] = None,
        max_n: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Greedy NMS on one set (offsets=None) or several ([P+1] int64 device offsets).
    Returns (keep_idx [P,max_det], keep_src [P,max_det], keep_count [P])."""
    _need_cuda(boxes, scores, classes, offsets)
    n = boxes.shape[0]
    dev = boxes.device
    boxes = boxes.to(torch.float32).contiguous()
    scores = scores.to(torch.float32).contiguous()
    if classes is not None:
        classes = classes.to(torch.int64).contiguous()
    if offsets is None:
        offsets = torch.tensor([0, n], dtype=torch.int64, device=dev)
        max_n = n
    elif max_n is None:
        max_n = n
    p = offsets.numel() - 1
    keep_idx = torch.empty((p, max_detections), dtype=torch.int64, device=dev)
    keep_src = torch.empty((p, max_detections), dtype=torch.int64, device=dev)
    keep_cnt = torch.empty((p,), dtype=torch.int32, device=dev)
    mode = _lib.HVS_NMS_CLASS_AWARE if class_aware else _lib.HVS_NMS_AGNOSTIC
    if class_aware and boxes_xyxy:
        mode |= _lib.HVS_NMS_BOXES_XYXY
    check(_lib.load().hvs_nms(_ptr(boxes), _ptr(scores), _ptr(classes), _ptr(offsets), p, max_n, score_threshold,
                              iou_threshold, max_detections, mode, _ptr(keep_idx), _ptr(keep_src), _ptr(keep_cnt),
                              _stream()), "hvs_nms")
    return keep_idx, keep_src, keep_cnt


@_on_device
def post_process(decoded: Sequence[Dict[str, torch.Tensor]], confidence_threshold: float = 0.5,
                 iou_threshold: float = 0.5, max_detections: int = 100):
    """Two-stage multi-scale NMS for a batch.  decoded: per-scale dicts from yolo_decode.
    Returns (det_boxes [B,max_det,4], det_scores [B,max_det], det_labels [B,max_det], det_count [B])."""
    s = len(decoded)
    b = decoded[0]["class_scores"].shape[0]
    dev = decoded[0]["boxes"].device
    keepalive = []
    boxes_p = (ctypes.c_void_p * s)()
    scores_p = (ctypes.c_void_p * s)()
    cls_p = (ctypes.c_void_p * s)()
    n_p = (ctypes.c_int * s)()
    for i, d in enumerate(decoded):
        _need_cuda(d["boxes"], d["class_scores"], d["class_indices"])
        bx = d["boxes"].reshape(b, -1, 4).to(torch.float32).contiguous()
        sc = d["class_scores"].reshape(b, -1).to(torch.float32).contiguous()
        ci = d["class_indices"].reshape(b, -1).to(torch.int64).contiguous()
        keepalive += [bx, sc, ci]
        boxes_p[i], scores_p[i], cls_p[i], n_p[i] = bx.data_ptr(), sc.data_ptr(), ci.data_ptr(), bx.shape[1]
    lib = _lib.load()
    ws_bytes = int(lib.hvs_post_process_workspace(b, s, max_detections))
    ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=dev)
    det_boxes = torch.empty((b, max_detections, 4), dtype=torch.float32, device=dev)
    det_scores = torch.empty((b, max_detections), dtype=torch.float32, device=dev)
    det_labels = torch.empty((b, max_detections), dtype=torch.int64, device=dev)
    det_count = torch.empty((b,), dtype=torch.int32, device=dev)
    check(lib.hvs_post_process(boxes_p, scores_p, cls_p, n_p, s, b, confidence_threshold, iou_threshold,
                               max_detections, _ptr(det_boxes), _ptr(det_scores), _ptr(det_labels), _ptr(det_count),
                               _ptr(ws), ws_bytes, _stream()), "hvs_post_process")
    return det_boxes, det_scores, det_labels, det_count
