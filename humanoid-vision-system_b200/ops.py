"""Functional wrappers: torch CUDA tensors in, C-ABI calls on the current stream, tensors out.

PyTorch is used for device memory and streams only; every computation below runs in
libhvs_b200.so.  CPU tensors are rejected (no fallback).
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import check


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*tensors: Optional[torch.Tensor]):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.HvsError("hvs_b200 ops need CUDA tensors (there is no CPU path)")


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


# ----------------------------------------------------------------------------- K1 stream mHC
def mhc_stream_fwd(x: torch.Tensor, phi: torch.Tensor, bias: torch.Tensor, alpha: torch.Tensor,
                   scale: torch.Tensor, sk_iters: int = 20, eps_rms: float = 1e-8, eps_sk: float = 1e-8,
                   want_y: bool = True, want_u: bool = False, want_coeffs: bool = False,
                   split_phi: bool = False, out: Optional[torch.Tensor] = None,
                   saved: Optional[torch.Tensor] = None
                   ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor], Optional[torch.Tensor]]:
    """x [T,n,C] bf16 -> (y [T,n,C] bf16, u [T,C] bf16, coeffs [T,n*n+2n] fp32).
    `saved` ([T, SAVED_STRIDE] fp32, from `new_saved`) receives the statistics `mhc_stream_bwd_saved` consumes."""
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
    flags = _lib.HVS_MHC_SPLIT_PHI if split_phi else 0
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


def mhc_stream_bwd_saved(x: torch.Tensor, dy: torch.Tensor, saved: torch.Tensor, phi: torch.Tensor, bias: torch.Tensor,
                         alpha: torch.Tensor, scale: torch.Tensor, sk_iters: int = 20, eps_rms: float = 1e-8,
                         eps_sk: float = 1e-8, out: Optional[torch.Tensor] = None,
                         workspace: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """Fused single-pass backward (dx and every parameter gradient) from the statistics saved by the forward."""
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
                                       eps_rms, eps_sk, 0, _ptr(ws), ws.numel(), _stream()), "hvs_mhc_stream_bwd_saved")
    return {"dx": dx, "dphi": dphi, "dbias": dbias, "dalpha": dalpha, "dscale": dscale}


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
def sinkhorn(matrix: torch.Tensor, iters: int = 20, eps: float = 1e-8, tau: float = 1.0,
             history: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need_cuda(matrix, history)
    if matrix.dtype != torch.float32:
        raise _lib.HvsError("sinkhorn expects fp32")
    m = matrix.contiguous()
    if m.dim() == 2:
        batch, n, mm = 1, m.shape[0], m.shape[1]
    else:
        n, mm = m.shape[-2], m.shape[-1]
        batch = m.numel() // max(n * mm, 1)
    out = torch.empty_like(m)
    check(_lib.load().hvs_sinkhorn(_ptr(m), _ptr(out), batch, n, mm, iters, eps, tau, _ptr(history), _stream()),
          "hvs_sinkhorn")
    return out


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


# ----------------------------------------------------------------------------- detection
_DTYPES = {torch.float32: _lib.HVS_DTYPE_F32, torch.float16: _lib.HVS_DTYPE_F16, torch.bfloat16: _lib.HVS_DTYPE_BF16}


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
