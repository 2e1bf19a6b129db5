"""nn.Module surface of the mHC hot path.

* ``StreamMHC``                 K1, the north_star stream layer (n residual streams, per-token Sinkhorn).
* ``SinkhornKnoppProjection``   reference signature (src/models/manifold_layers.py:25-101).
* ``RMSNorm``                   reference signature (:437-456).
* ``ManifoldHyperConnection``   K2, the reference-literal module (:104-346): same constructor, same
                                state_dict keys, same ``constrained_matrices`` / ``get_stability_metrics``.

All coefficient work runs in libhvs_b200.so; PyTorch supplies device memory, streams, autograd
plumbing and (for K2) the plain library GEMMs / LayerNorm.
"""
from __future__ import annotations

import math
from typing import Any, Callable, Dict, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from ._lib import HvsError


# ----------------------------------------------------------------------------- K1
FUSED_BWD_MAX_ITERS = 24      # both backward kernels keep every iteration's scalings in shared memory
FWD_MAX_ITERS = 64


def _check_trainable_shape(n: int, c: int, split_phi: bool = False):
    """n = 4, C = 512 trains on the tuned kernels (forward saving statistics + ONE fused backward kernel); every other
    n in {2, 4}, C % 8 == 0, C <= 1024 on the general three-launch backward (mhc_stream_generic_bwd.cu)."""
    if n not in (2, 4) or c % 8 or not 8 <= c <= 1024 or split_phi:
        raise HvsError("stream mHC training kernels take n_streams in {2, 4}, channels % 8 == 0, channels <= 1024 with the bf16 "
                       f"projection operand; got n = {n}, C = {c}, split_phi = {split_phi}")


def _check_trainable_iters(sk_iters: int):
    """The forward kernel takes up to 64 Sinkhorn iterations; the backward kernels (fused single-pass and the
    two-kernel recompute form alike) differentiate at most 24.  Reject a training call up front instead of
    failing inside autograd after the forward has run."""
    if sk_iters > FUSED_BWD_MAX_ITERS:
        raise HvsError(f"stream mHC training supports sk_iterations <= {FUSED_BWD_MAX_ITERS} "
                       f"(got {sk_iters}); inference (no grad) supports up to {FWD_MAX_ITERS}")


class _StreamMHCFn(torch.autograd.Function):
    """Training path.  The forward additionally writes 112 B/token of statistics (un-normalised projection and
    sum of squares); the backward is then ONE fused kernel (dx and every parameter gradient in a single pass over
    x and dy).  sk_iters <= 24 (checked before the forward runs)."""

    @staticmethod
    def forward(ctx, x, phi, bias, alpha, scale, sk_iters, eps_rms, eps_sk, adaptive=False):
        _check_trainable_iters(sk_iters)
        _check_trainable_shape(x.shape[1], x.shape[2])
        fused = (x.shape[1], x.shape[2]) == (4, 512)      # tuned single-pass backward; other shapes recompute (general kernels)
        adaptive = bool(adaptive) and fused
        saved = ops.new_saved(x) if fused else None
        y, _, _ = ops.mhc_stream_fwd(x, phi, bias, alpha, scale, sk_iters, eps_rms, eps_sk, saved=saved, adaptive=adaptive)
        if fused:
            ctx.save_for_backward(x, phi, bias, alpha, scale, saved)
        else:
            ctx.save_for_backward(x, phi, bias, alpha, scale)
        ctx.cfg = (sk_iters, eps_rms, eps_sk, fused, adaptive)
        return y

    @staticmethod
    def backward(ctx, dy):
        sk_iters, eps_rms, eps_sk, fused, adaptive = ctx.cfg
        if fused:
            x, phi, bias, alpha, scale, saved = ctx.saved_tensors
            g = ops.mhc_stream_bwd_saved(x, dy.contiguous(), saved, phi, bias, alpha, scale, sk_iters, eps_rms, eps_sk, adaptive=adaptive)
        else:
            x, phi, bias, alpha, scale = ctx.saved_tensors
            g = ops.mhc_stream_bwd(x, dy.contiguous(), phi, bias, alpha, scale, sk_iters, eps_rms, eps_sk)
        return g["dx"], g["dphi"], g["dbias"], g["dalpha"], g["dscale"], None, None, None, None


class StreamMHC(nn.Module):
    """Stream mHC residual layer:  y = H_res x + H_post (x) fn(H_pre^T x).

    x: [..., n, C] bf16.  n = 4, C = 512 runs on the tuned TMA / tensor-core kernels (forward AND the fused training
    backward); every other n in {2, 4}, C % 8 == 0, C <= 1024 runs on the general kernels (forward, and a three-launch
    backward whose dW contraction is the tcgen05 GEMM); ``split_phi=True`` (fp32-accurate projection operand,
    HVS_MHC_SPLIT_PHI) is forward-only for all shapes.  ``fn=None`` is the
    identity (one fused kernel).  With a wrapped layer ``fn`` the forward runs the coefficient kernel, ``fn`` on the
    bf16 layer input, and the mixing kernel (inference path).
    """

    def __init__(self, n_streams: int = 4, channels: int = 512, alpha: float = 0.01, sk_iterations: int = 20,
                 eps: float = 1e-8, phi_std: float = 0.02, fn: Optional[Callable[[torch.Tensor], torch.Tensor]] = None,
                 device=None, split_phi: bool = False, adaptive_sinkhorn: bool = False):
        super().__init__()
        n, c = n_streams, channels
        self.split_phi = split_phi
        self.adaptive_sinkhorn = adaptive_sinkhorn         # HVS_MHC_ADAPTIVE_ITERS (n = 4, C = 512): stop at the bitwise fixed point
        k = n * n + 2 * n
        self.n_streams, self.channels, self.sk_iterations, self.eps = n, c, sk_iterations, eps
        self.phi = nn.Parameter(torch.randn(n * c, k, device=device) * phi_std)
        self.bias = nn.Parameter(torch.zeros(k, device=device))
        self.alpha = nn.Parameter(torch.full((3,), float(alpha), device=device))   # reference `alpha` (:134)
        self.rms_scale = nn.Parameter(torch.ones(n * c, device=device))            # RMSNorm.scale (:446)
        self.fn = fn

    def coefficients(self, x: torch.Tensor):
        """(H_pre [T,n], H_post [T,n], H_res [T,n,n]) for x [T,n,C]."""
        t, n = x.shape[0], self.n_streams
        _, _, co = ops.mhc_stream_fwd(x, self.phi, self.bias, self.alpha, self.rms_scale, self.sk_iterations,
                                      self.eps, self.eps, want_y=False, want_coeffs=True, split_phi=self.split_phi)
        return co[:, :n], co[:, n:2 * n], co[:, 2 * n:].reshape(t, n, n)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        shape = x.shape
        xf = x.reshape(-1, self.n_streams, self.channels).contiguous()
        needs_grad = torch.is_grad_enabled() and (xf.requires_grad or any(p.requires_grad for p in self.parameters()))
        if self.fn is None and not needs_grad:
            y, _, _ = ops.mhc_stream_fwd(xf, self.phi, self.bias, self.alpha, self.rms_scale, self.sk_iterations,
                                         self.eps, self.eps, split_phi=self.split_phi,
                                         adaptive=self.adaptive_sinkhorn and not self.split_phi and (self.n_streams, self.channels) == (4, 512))
        elif self.fn is None:
            _check_trainable_shape(self.n_streams, self.channels, self.split_phi)
            y = _StreamMHCFn.apply(xf, self.phi, self.bias, self.alpha, self.rms_scale, self.sk_iterations,
                                   self.eps, self.eps, self.adaptive_sinkhorn)
        else:
            if torch.is_grad_enabled() and (xf.requires_grad or self.phi.requires_grad):
                raise HvsError("StreamMHC with a wrapped fn is forward-only in this build (fused backward covers fn=None)")
            _, u, co = ops.mhc_stream_fwd(xf, self.phi, self.bias, self.alpha, self.rms_scale, self.sk_iterations,
                                          self.eps, self.eps, want_y=False, want_u=True, want_coeffs=True, split_phi=self.split_phi)
            fu = self.fn(u).to(torch.bfloat16).contiguous()
            y = ops.mhc_stream_post(xf, co, fu)
        return y.reshape(shape)


def stream_mhc_fwd_bwd_host(x_host: torch.Tensor, dy_host: torch.Tensor, layer: StreamMHC,
                            y_host: torch.Tensor, dx_host: torch.Tensor, chunk_tokens: int = 1 << 15,
                            device: Optional[torch.device] = None) -> Dict[str, torch.Tensor]:
    """Host-buffer entry: x, dy (pinned host, bf16 [T,n,C]) -> y, dx written to pinned host buffers,
    parameter gradients returned on the host.  Chunks are pipelined over three streams (H2D, compute,
    D2H) with double-buffered device staging, so copies overlap the kernels.  32 k-token chunks (268 MB per tensor)
    measured best on a B200 box: 181 ms for 2^20 tokens = 47.5 GB/s each way, against 49.9 GB/s for bare concurrent
    H2D + D2H copies of the same buffers (tools/e2e_sweep.py)."""
    dev = device or layer.phi.device
    _check_trainable_iters(layer.sk_iterations)
    t = x_host.shape[0]
    n, c = layer.n_streams, layer.channels
    nchunks = (t + chunk_tokens - 1) // chunk_tokens
    cur = torch.cuda.current_stream(dev)
    s_in, s_cmp, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    # the caller may just have written the parameters (optimizer step) or be using memory the staging buffers
    # will be carved from on its own stream: order the three side streams after it
    for side in (s_in, s_cmp, s_out):
        side.wait_stream(cur)
    bufs = [{k: torch.empty((chunk_tokens, n, c), dtype=torch.bfloat16, device=dev) for k in ("x", "dy", "y", "dx")}
            for _ in range(2)]
    fused = True
    saved = [ops.new_saved(bufs[0]["x"]) for _ in range(2)] if fused else None
    ws = None
    if fused:
        from . import _lib
        ws = torch.empty(int(_lib.load().hvs_mhc_stream_bwd_saved_workspace(chunk_tokens, n, c)), dtype=torch.uint8, device=dev)
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_cmp = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]
    acc = None
    params = (layer.phi.detach(), layer.bias.detach(), layer.alpha.detach(), layer.rms_scale.detach())
    for i in range(nchunks):
        lo, hi = i * chunk_tokens, min(t, (i + 1) * chunk_tokens)
        m = hi - lo
        b = bufs[i % 2]
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_out[i % 2])                       # staging buffers drained by chunk i-2
            b["x"][:m].copy_(x_host[lo:hi], non_blocking=True)
            b["dy"][:m].copy_(dy_host[lo:hi], non_blocking=True)
            ev_in[i % 2].record(s_in)
        with torch.cuda.stream(s_cmp):
            s_cmp.wait_event(ev_in[i % 2])
            if fused:
                sv = saved[i % 2][:m]
                ops.mhc_stream_fwd(b["x"][:m], *params, layer.sk_iterations, layer.eps, layer.eps, out=b["y"][:m], saved=sv)
                g = ops.mhc_stream_bwd_saved(b["x"][:m], b["dy"][:m], sv, *params, layer.sk_iterations, layer.eps,
                                             layer.eps, out=b["dx"][:m], workspace=ws)
            else:
                ops.mhc_stream_fwd(b["x"][:m], *params, layer.sk_iterations, layer.eps, layer.eps, out=b["y"][:m])
                g = ops.mhc_stream_bwd(b["x"][:m], b["dy"][:m], *params, layer.sk_iterations, layer.eps, layer.eps)
                b["dx"][:m].copy_(g["dx"])
            if acc is None:
                acc = {k: g[k].clone() for k in ("dphi", "dbias", "dalpha", "dscale")}
            else:
                for k in acc:
                    acc[k] += g[k]
            ev_cmp[i % 2].record(s_cmp)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_cmp[i % 2])
            y_host[lo:hi].copy_(b["y"][:m], non_blocking=True)
            dx_host[lo:hi].copy_(b["dx"][:m], non_blocking=True)
            ev_out[i % 2].record(s_out)
    out = {}
    with torch.cuda.stream(s_out):
        s_out.wait_stream(s_cmp)
        for k, v in (acc or {}).items():
            out[k] = v.to("cpu", non_blocking=False)
    s_out.synchronize()
    cur.wait_stream(s_out)                                   # staging buffers return to the caller's pool in order
    return out


# ----------------------------------------------------------------------------- reference-signature modules
def _pad64(v: int) -> int:
    """K extent of the bf16 operand copies.  The GEMM's tensor maps carry the true extents and TMA zero-fills a partial
    64-element stage, so no padding is needed any more; kept as the single place that decides it."""
    return v


class _SquareSinkhornFn(torch.autograd.Function):
    """Differentiable D x D projection: forward and backward are the batched static-coefficient kernels
    (hvs_mhc_static_coeffs / _bwd) with a single job; no unrolled autograd graph."""

    @staticmethod
    def forward(ctx, raw, iters, eps, history):
        d = raw.shape[0]
        raw = raw.contiguous()
        out = torch.empty_like(raw)
        uv = torch.empty((iters + 1, 2, d), dtype=torch.float32, device=raw.device)
        dummy = raw.new_zeros((d, 1))
        ops.static_coeffs([{"h_pre_raw": dummy, "h_post_raw": None, "h_res_raw": raw, "h_res": out, "uv_history": uv,
                            "convergence": history}], iters, eps)
        ctx.save_for_backward(raw, uv)
        ctx.cfg = (iters, eps)
        return out

    @staticmethod
    def backward(ctx, g):
        raw, uv = ctx.saved_tensors
        iters, eps = ctx.cfg
        d_raw = torch.empty_like(raw)
        dummy = raw.new_zeros((raw.shape[0], 1))
        ops.static_coeffs_bwd([{"h_pre_raw": dummy, "h_res_raw": raw, "h_res": d_raw, "uv_history": uv}],
                              [{"d_h_res": g.contiguous().float(), "d_h_res_raw": d_raw}], iters, eps)
        return d_raw, None, None, None


class SinkhornKnoppProjection(nn.Module):
    """SinkhornKnoppProjection(num_iterations=20, epsilon=1e-8, tau=1.0)  (manifold_layers.py:25-101).
    ``forward`` runs hvs_sinkhorn; the ``convergence_history`` buffer is filled on the device without
    the reference's 3 host synchronisations per iteration.  A square 2-D input that requires grad goes through
    the static-coefficient kernels (forward + exact reverse sweep); only batched / non-square inputs that require
    grad -- which no model in the reference has -- use the torch-op restatement below."""

    def __init__(self, num_iterations: int = 20, epsilon: float = 1e-8, tau: float = 1.0):
        super().__init__()
        self.num_iterations, self.epsilon, self.tau = num_iterations, epsilon, tau
        self.register_buffer("convergence_history", torch.zeros(num_iterations))

    def forward(self, matrix: torch.Tensor, return_history: bool = False):
        if torch.is_grad_enabled() and matrix.requires_grad:
            if matrix.dim() == 2 and matrix.shape[0] == matrix.shape[1] and self.tau == 1.0 and matrix.dtype == torch.float32:
                out = _SquareSinkhornFn.apply(matrix, self.num_iterations, self.epsilon, self.convergence_history)
            else:
                out = _sinkhorn_autograd(matrix, self.num_iterations, self.epsilon, self.tau)
        else:
            out = ops.sinkhorn(matrix.detach().float(), self.num_iterations, self.epsilon, self.tau,
                               history=self.convergence_history)
        if return_history:
            hist = self.convergence_history.detach().cpu()
            last = float(hist[-1]) if len(hist) else 0.0
            return out, {"row_sums": (hist + 1.0).tolist(), "col_sums": [1.0] * len(hist),
                         "final_row_error": last, "final_col_error": 0.0}
        return out

    def get_convergence_metrics(self) -> Dict[str, Any]:
        h = self.convergence_history
        return {"mean_convergence": h.mean().item(), "max_convergence": h.max().item(),
                "final_convergence": h[-1].item()}


def _sinkhorn_autograd(matrix, iters, eps, tau):
    """Batched / non-square projection that requires grad (not on any model's path): the reference arithmetic in
    torch device ops."""
    squeeze = matrix.dim() == 2
    p = matrix.unsqueeze(0) if squeeze else matrix
    m = p.shape[-1]
    p = torch.softmax(p / tau, dim=-1) * m
    for _ in range(iters):
        p = p / (p.sum(dim=-1, keepdim=True) + eps)
        p = p / (p.sum(dim=-2, keepdim=True) + eps)
    return p.squeeze(0) if squeeze else p


class _RMSNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scale, eps):
        ctx.save_for_backward(x, scale)
        ctx.eps = eps
        return ops.rmsnorm_fwd(x, scale, eps)

    @staticmethod
    def backward(ctx, dy):
        x, scale = ctx.saved_tensors
        dx, dscale = ops.rmsnorm_bwd(x, scale, dy.to(x.dtype), ctx.eps)
        return dx, dscale, None


class RMSNorm(nn.Module):
    """RMSNorm(dim, eps=1e-8) (manifold_layers.py:437-456); hvs_rmsnorm_fwd / _bwd (fp32 or bf16 data, fp32 stats)."""

    def __init__(self, dim: int, eps: float = 1e-8):
        super().__init__()
        self.scale = nn.Parameter(torch.ones(dim))
        self.eps = eps

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.float()
        if torch.is_grad_enabled() and (x.requires_grad or self.scale.requires_grad):
            return _RMSNormFn.apply(x, self.scale, self.eps)
        return ops.rmsnorm_fwd(x, self.scale.detach(), self.eps)


class _LayerNormFn(torch.autograd.Function):
    """nn.LayerNorm over the last axis on this library's row kernels (training of the module's norm_pre / norm_post; the
    ATen kernels spend 40 ms per 16-image step on the backbone's 32 / 64-wide rows).  fp32 statistics; the output dtype
    is chosen by the caller: bf16 where the only consumer is an autocast matmul (one rounding of the fp32 result either
    way), fp32 for the module output."""

    @staticmethod
    def forward(ctx, x2, weight, bias, eps, out_dtype):
        x2 = x2.contiguous()
        out, _ = ops.layernorm_fwd(x2, weight.detach(), bias.detach(), eps, out_dtype=out_dtype)
        ctx.save_for_backward(x2, weight)
        ctx.eps = eps
        return out

    @staticmethod
    def backward(ctx, dy):
        x2, weight = ctx.saved_tensors
        if dy.dtype not in (torch.float32, torch.bfloat16):
            dy = dy.float()
        dx, dw, db = ops.layernorm_bwd(x2, weight.detach(), dy, ctx.eps)
        return dx, dw, db, None, None


def _layer_norm(mod: nn.LayerNorm, x: torch.Tensor, out_dtype: torch.dtype) -> torch.Tensor:
    """mod(x) through _LayerNormFn when the kernels take the shape, else the module itself."""
    d = x.shape[-1]
    if x.is_cuda and d in ops.LN_BWD_DIMS and x.dtype in (torch.float32, torch.bfloat16) and mod.weight.dtype == torch.float32:
        return _LayerNormFn.apply(x.reshape(-1, d), mod.weight, mod.bias, mod.eps, out_dtype).reshape(x.shape)
    return mod(x)


class _CoeffState:
    """Per-module device buffers of the static coefficient path (fp32 matrices, their bf16 transposed copies for the
    GEMM B operands, the scaling history the backward consumes, bf16 copies of the MLP weights)."""

    def __init__(self, mod: "ManifoldHyperConnection"):
        d, h, dev = mod.input_dim, mod.hidden_dim, mod.H_res_raw.device
        dp = _pad64(d)
        it = mod.sinkhorn.num_iterations
        f32, bf = torch.float32, torch.bfloat16
        self.h_pre = torch.empty((d, h), dtype=f32, device=dev)
        self.h_post = torch.empty((h, d), dtype=f32, device=dev)
        self.h_res = torch.empty((d, d), dtype=f32, device=dev)
        self.h_pre_t = torch.zeros((h, dp), dtype=bf, device=dev)
        self.h_post_t = torch.zeros((d, h), dtype=bf, device=dev)
        self.h_res_t = torch.zeros((d, dp), dtype=bf, device=dev)
        self.uv_history = torch.empty((it + 1, 2, d), dtype=f32, device=dev)
        self.w1 = self.w2 = None
        self.key = None
        self.wkey = None

    def job(self, mod) -> Dict[str, Optional[torch.Tensor]]:
        return {"h_pre_raw": mod.H_pre_raw.detach(), "h_post_raw": mod.H_post_raw.detach(), "h_res_raw": mod.H_res_raw.detach(),
                "h_pre": self.h_pre, "h_post": self.h_post, "h_res": self.h_res, "h_pre_t": self.h_pre_t,
                "h_post_t": self.h_post_t, "h_res_t": self.h_res_t, "uv_history": self.uv_history,
                "convergence": mod.sinkhorn.convergence_history}


def refresh_static_coefficients(model: nn.Module, force: bool = False) -> int:
    """constrained_matrices (:205-221) of EVERY ManifoldHyperConnection under `model` whose parameters changed since
    the last refresh, in ONE kernel launch (the reference recomputes each layer's, with 60 host syncs, on every
    forward).  Returns the number of layers refreshed.  Called by the hybrid_vision harness before a forward / at the
    top of a training step; a module whose cache is stale when its own forward runs refreshes itself."""
    mods = [m for m in model.modules() if isinstance(m, ManifoldHyperConnection) and m.H_res_raw.is_cuda]
    # a training step that is being captured into a CUDA graph must contain the refresh: the host-side cache check does
    # not run when the graph is replayed, the parameters change every step
    force = force or (torch.is_grad_enabled() and torch.cuda.is_current_stream_capturing())
    todo = [m for m in mods if force or m._state is None or m._state.key != m._key()]
    groups: Dict[Tuple[Any, int, float], list] = {}
    for m in todo:
        groups.setdefault((m.H_res_raw.device, m.sinkhorn.num_iterations, m.sinkhorn.epsilon), []).append(m)
    for (_, iters, eps), ms in groups.items():
        jobs = []
        for m in ms:
            if m._state is None:
                m._state = _CoeffState(m)
            jobs.append(m._state.job(m))
        ops.static_coeffs(jobs, iters, eps)
        for m in ms:
            m._state.key = m._key()
    return len(todo)


class _CoeffFn(torch.autograd.Function):
    """(H_pre_raw, H_post_raw, H_res_raw) -> (H_pre, H_post, H_res) for training: forward reads the module's
    (batched-refreshed) state, backward is hvs_mhc_static_coeffs_bwd."""

    @staticmethod
    def forward(ctx, mod, pre_raw, post_raw, res_raw):
        st = mod._fresh_state()
        ctx.mod = mod
        ctx.uv = st.uv_history.clone()                 # the state may be refreshed (optimizer step) before backward runs
        ctx.save_for_backward(pre_raw, post_raw, res_raw)
        return st.h_pre.clone(), st.h_post.clone(), st.h_res.clone()

    @staticmethod
    def backward(ctx, d_pre, d_post, d_res):
        pre_raw, post_raw, res_raw = ctx.saved_tensors
        mod = ctx.mod
        g = {"d_h_pre": None, "d_h_post": None, "d_h_res": None, "d_h_pre_raw": None, "d_h_post_raw": None, "d_h_res_raw": None}
        outs = [None, None, None]
        for i, (name, d, raw) in enumerate((("pre", d_pre, pre_raw), ("post", d_post, post_raw), ("res", d_res, res_raw))):
            if d is not None and ctx.needs_input_grad[i + 1]:
                g[f"d_h_{name}"] = d.contiguous().float()
                outs[i] = torch.empty_like(raw, memory_format=torch.contiguous_format)
                g[f"d_h_{name}_raw"] = outs[i]
        job = {"h_pre_raw": pre_raw.detach().contiguous(), "h_post_raw": post_raw.detach().contiguous(),
               "h_res_raw": res_raw.detach().contiguous(), "h_res": mod._state.h_res, "uv_history": ctx.uv}
        ops.static_coeffs_bwd([job], [g], mod.sinkhorn.num_iterations, mod.sinkhorn.epsilon)
        return None, outs[0], outs[1], outs[2]


_DROPOUT_CALLS = [0]
_DROPOUT_STEP: Dict[Any, torch.Tensor] = {}


def _next_dropout_seed() -> int:
    """A fresh 32-bit seed per training forward, reproducible under torch.manual_seed."""
    _DROPOUT_CALLS[0] += 1
    return (torch.initial_seed() * 0x9E3779B1 + _DROPOUT_CALLS[0] * 0x85EBCA6B) & 0xFFFFFFFF


def dropout_step_counter(device) -> torch.Tensor:
    """Device-resident step counter mixed into every dropout seed of the module kernels on `device`.  A training step
    captured in a CUDA graph bakes the host-side seeds into the graph; advancing this counter INSIDE the captured step
    (``dropout_step_counter(dev).add_(1)``) gives every replay new masks, the same in its forward and its backward."""
    dev = torch.device(device)
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    if key not in _DROPOUT_STEP:
        _DROPOUT_STEP[key] = torch.zeros(1, dtype=torch.int32, device=dev)
    return _DROPOUT_STEP[key]


class _K2TokenPathFn(torch.autograd.Function):
    """Training form of the module's token path (manifold_layers.py:248-267) on this library's kernels: the forward is
    LayerNorm -> four tcgen05 GEMM launches (bias + GELU + dropout epilogues that also keep the pre-activations) ->
    LayerNorm; the backward is the two LayerNorm backward kernels, five data-gradient GEMMs (GELU' x dropout mask in the
    epilogue, which also sums the bias gradients per 32-row group; the weights are read as they lie, MN-major) and five
    weight-gradient GEMMs over the token axis (both activations MN-major, split-K with a fixed-order reduction).
    bf16 operands, fp32 accumulation, bf16 activations between the GEMMs: the reference's own CUDA-autocast arithmetic."""

    @staticmethod
    def forward(ctx, mod, x2, h_pre, h_post, h_res, w1, b1, w2, b2, g_pre, be_pre, g_post, be_post):
        from . import _lib
        st = mod._fresh_state()
        w1b, w2b = mod._mlp_bf16()
        p = float(mod.mlp[2].p) if mod.training else 0.0
        seed1 = _next_dropout_seed() if p > 0 else 0
        seed2 = seed1 ^ 0x5BD1E995
        sdev = dropout_step_counter(x2.device) if p > 0 else None
        x2 = x2.contiguous()
        bf = torch.bfloat16
        xn, xb = ops.layernorm_fwd(x2, g_pre.detach(), be_pre.detach(), mod.norm_pre.eps, out_dtype=bf, want_copy=x2.dtype != bf)
        if xb is None:
            xb = x2
        h0 = ops.gemm_bf16(xn, st.h_pre_t)                                                               # :253
        a1, z1 = ops.gemm_bf16_ex(h0, w1b, bias=b1.detach(), epilogue=_lib.HVS_GEMM_EPI_BIAS_GELU_SAVE, dropout_p=p, dropout_seed=seed1,
                                      dropout_seed_dev=sdev)
        a2, z2 = ops.gemm_bf16_ex(a1, w2b, bias=b2.detach(), epilogue=_lib.HVS_GEMM_EPI_BIAS_GELU_SAVE, dropout_p=p, dropout_seed=seed2,
                                      dropout_seed_dev=sdev)
        pre = ops.gemm_bf16(a2, st.h_post_t, xb, st.h_res_t, out_dtype=bf)                               # :259-263
        out, _ = ops.layernorm_fwd(pre, g_post.detach(), be_post.detach(), mod.norm_post.eps, out_dtype=mod.output_dtype or torch.float32)
        ctx.mod, ctx.cfg, ctx.key = mod, (p, seed1, seed2, sdev), st.key
        ctx.save_for_backward(x2, xn, xb, h0, z1, a1, z2, a2, pre, w1b, w2b, g_pre, g_post)
        return out

    @staticmethod
    def backward(ctx, dout):
        from . import _lib
        mod = ctx.mod
        st = mod._state
        if st is None or st.key != ctx.key:
            raise HvsError("static coefficients were refreshed (parameters changed) between the forward and its backward")
        p, seed1, seed2, sdev = ctx.cfg
        x2, xn, xb, h0, z1, a1, z2, a2, pre, w1b, w2b, g_pre, g_post = ctx.saved_tensors
        dge = _lib.HVS_GEMM_EPI_DGELU
        if dout.dtype not in (torch.float32, torch.bfloat16):
            dout = dout.float()
        d_pre, dg_post, dbe_post = ops.layernorm_bwd(pre, g_post.detach(), dout.contiguous(), mod.norm_post.eps)       # bf16 [T, D]
        dz2, d_b2 = ops.gemm_bf16_ex(d_pre, st.h_post_t, b_mn=True, epilogue=dge, aux=z2, dropout_p=p, dropout_seed=seed2,
                                     dropout_seed_dev=sdev, want_colsum=True)                                                 # [T, H], [H]
        dx_res = ops.gemm_bf16_ex(d_pre, st.h_res_t, b_mn=True)                                                        # [T, D]
        d_h_post = ops.gemm_wgrad(a2, d_pre)                                                                           # [H, D]
        d_h_res = ops.gemm_wgrad(xb, d_pre)                                                                            # [D, D]
        d_w2 = ops.gemm_wgrad(dz2, a1)                                                                                 # [H, 2H]
        dz1, d_b1 = ops.gemm_bf16_ex(dz2, w2b, b_mn=True, epilogue=dge, aux=z1, dropout_p=p, dropout_seed=seed1,
                                     dropout_seed_dev=sdev, want_colsum=True)                                                 # [T, 2H], [2H]
        d_w1 = ops.gemm_wgrad(dz1, h0)                                                                                 # [2H, H]
        dh0 = ops.gemm_bf16_ex(dz1, w1b, b_mn=True)                                                                    # [T, H]
        d_h_pre = ops.gemm_wgrad(xn, dh0)                                                                              # [D, H]
        dxn = ops.gemm_bf16_ex(dh0, st.h_pre_t, b_mn=True)                                                             # [T, D]
        dx, dg_pre, dbe_pre = ops.layernorm_bwd(x2, g_pre.detach(), dxn, mod.norm_pre.eps)
        dx = dx + dx_res.to(dx.dtype)
        return None, dx, d_h_pre, d_h_post, d_h_res, d_w1, d_b1, d_w2, d_b2, dg_pre, dbe_pre, dg_post, dbe_post


class _AllCoeffsFn(torch.autograd.Function):
    """constrained_matrices of EVERY layer of a model as ONE autograd node: the forward is the single batched launch of
    refresh_static_coefficients, the backward ONE hvs_mhc_static_coeffs_bwd launch for all layers (autograd runs it when
    every layer's dH has been accumulated, at the end of the backward pass) instead of one cooperative launch per layer
    (76 launches, 20 ms of a 180 ms training step)."""

    @staticmethod
    def forward(ctx, mods, *raws):
        ctx.mods = mods
        ctx.keys = [m._state.key for m in mods]
        outs = []
        for m in mods:
            st = m._state
            outs += [st.h_pre.detach(), st.h_post.detach(), st.h_res.detach()]   # aliases of the state buffers (no copy)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *douts):
        mods = ctx.mods
        jobs, grads, outs = [], [], []
        for i, m in enumerate(mods):
            if m._state.key != ctx.keys[i]:
                raise HvsError("static coefficients were refreshed (parameters changed) between the forward and its backward")
            st = m._state
            g: Dict[str, Optional[torch.Tensor]] = {}
            any_grad = False
            for j, (name, raw) in enumerate((("pre", m.H_pre_raw), ("post", m.H_post_raw), ("res", m.H_res_raw))):
                d = douts[3 * i + j]
                if d is not None and ctx.needs_input_grad[1 + 3 * i + j]:
                    g[f"d_h_{name}"] = d.contiguous().float()
                    o = torch.empty_like(raw, memory_format=torch.contiguous_format)
                    g[f"d_h_{name}_raw"] = o
                    outs.append(o)
                    any_grad = True
                else:
                    outs.append(None)
            if any_grad:
                jobs.append({"h_pre_raw": m.H_pre_raw.detach(), "h_post_raw": m.H_post_raw.detach(), "h_res_raw": m.H_res_raw.detach(),
                             "h_res": st.h_res, "uv_history": st.uv_history})
                grads.append(g)
        # one launch per (iterations, eps) group -- a model built from the reference has exactly one
        by_cfg: Dict[Tuple[int, float], Tuple[list, list]] = {}
        ji = 0
        for i, m in enumerate(mods):
            if any(outs[3 * i + j] is not None for j in range(3)):
                cfg = (m.sinkhorn.num_iterations, m.sinkhorn.epsilon)
                by_cfg.setdefault(cfg, ([], []))
                by_cfg[cfg][0].append(jobs[ji])
                by_cfg[cfg][1].append(grads[ji])
                ji += 1
        for (iters, eps), (js, gs) in by_cfg.items():
            ops.static_coeffs_bwd(js, gs, iters, eps)
        return (None, *outs)


def batched_training_coefficients(model: nn.Module) -> int:
    """Training counterpart of refresh_static_coefficients: refreshes every layer's coefficients in one launch and hangs
    them on ONE autograd node, so that the coefficient backward of all layers is one launch too.  Each module's
    constrained_matrices() returns its slice until clear_training_coefficients(model) is called (the host model does both
    around its forward).  Returns the number of layers covered."""
    mods = [m for m in model.modules() if isinstance(m, ManifoldHyperConnection) and m.H_res_raw.is_cuda
            and any(p.requires_grad for p in (m.H_pre_raw, m.H_post_raw, m.H_res_raw))]
    if not mods:
        return 0
    refresh_static_coefficients(model)
    raws = []
    for m in mods:
        raws += [m.H_pre_raw, m.H_post_raw, m.H_res_raw]
    outs = _AllCoeffsFn.apply(mods, *raws)
    for i, m in enumerate(mods):
        m._train_coeffs = (outs[3 * i], outs[3 * i + 1], outs[3 * i + 2])
    return len(mods)


def clear_training_coefficients(model: nn.Module) -> None:
    for m in model.modules():
        if isinstance(m, ManifoldHyperConnection):
            m._train_coeffs = None


class ManifoldHyperConnection(nn.Module):
    """Drop-in for the reference ManifoldHyperConnection (manifold_layers.py:104-346).

    Same constructor, parameters, buffers and state_dict keys (H_pre_raw, H_post_raw, H_res_raw,
    gradient_norms, eigenvalues, signal_ratio_history, sinkhorn.convergence_history, mlp.{0,3}.*,
    norm_pre.*, norm_post.*).  What runs where:
      * constrained_matrices: hvs_mhc_static_coeffs (all layers of a model in one launch through
        ``refresh_static_coefficients``), cached until a parameter changes; its backward is
        hvs_mhc_static_coeffs_bwd (no unrolled Sinkhorn graph);
      * inference forward (no grad): hvs_layernorm_fwd -> four hvs_gemm_bf16 launches (tcgen05, fused bias+GELU and
        residual + LayerNorm epilogues) under the reference's CUDA-autocast convention (bf16 operands, fp32
        accumulation, fp32 output);
      * training forward (grad): the same token path in torch ops so autograd sees it (the K2 backward GEMMs are
        library calls), with the coefficient kernels above;
      * stability monitoring (:282-316) is computed on demand in get_stability_metrics instead of an eigvalsh inside
        every training forward.
    """

    def __init__(self, input_dim: int, expansion_rate: int = 4, hidden_dim: Optional[int] = None, alpha: float = 0.01,
                 sk_iterations: int = 20, use_mixed_precision: bool = True, dropout_rate: float = 0.1):
        super().__init__()
        self.input_dim = input_dim
        self.expansion_rate = expansion_rate
        self.hidden_dim = hidden_dim or (input_dim * expansion_rate)
        self.alpha, self.use_mixed_precision, self.dropout_rate = alpha, use_mixed_precision, dropout_rate
        self.H_pre_raw = nn.Parameter(torch.empty(input_dim, self.hidden_dim))
        self.H_post_raw = nn.Parameter(torch.empty(self.hidden_dim, input_dim))
        self.H_res_raw = nn.Parameter(torch.empty(input_dim, input_dim))
        self.sinkhorn = SinkhornKnoppProjection(sk_iterations)
        self.mlp = nn.Sequential(nn.Linear(self.hidden_dim, self.hidden_dim * 2), nn.GELU(), nn.Dropout(dropout_rate),
                                 nn.Linear(self.hidden_dim * 2, self.hidden_dim), nn.GELU(), nn.Dropout(dropout_rate))
        self.norm_pre = nn.LayerNorm(input_dim)
        self.norm_post = nn.LayerNorm(input_dim)
        self.dropout = nn.Dropout(dropout_rate)
        self.register_buffer("gradient_norms", torch.zeros(3))
        self.register_buffer("eigenvalues", torch.zeros(input_dim))
        self.register_buffer("signal_ratio_history", torch.zeros(1000))
        self.signal_ratio_idx = 0
        self.dtype = torch.bfloat16 if use_mixed_precision else torch.float32
        self._state: Optional[_CoeffState] = None
        self._train_coeffs: Optional[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = None   # batched_training_coefficients
        self.monitor_signal_ratio = True
        self.output_dtype: Optional[torch.dtype] = None      # None: fp32 like the reference (its last op is a LayerNorm under autocast)
        self.use_chain_kernel = True                          # (D, H) = (32, 128) / (64, 256), bf16 input: hvs_mhc_module_fwd
        self.use_training_kernels = True                      # grad / train mode: _K2TokenPathFn instead of torch ops + library GEMMs
        self._initialize_weights()

    def _initialize_weights(self):                       # :191-203
        for w in (self.H_pre_raw, self.H_post_raw, self.H_res_raw):
            nn.init.xavier_uniform_(w, gain=0.1)
        for layer in self.mlp:
            if isinstance(layer, nn.Linear):
                nn.init.xavier_uniform_(layer.weight, gain=math.sqrt(2))
                nn.init.zeros_(layer.bias)

    def _key(self):
        return tuple((p.data_ptr(), p._version, p.device) for p in (self.H_pre_raw, self.H_post_raw, self.H_res_raw))

    def _fresh_state(self) -> _CoeffState:
        if self._state is None or self._state.key != self._key():
            refresh_static_coefficients(self)
        return self._state

    def constrained_matrices(self) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:      # :205-221
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in (self.H_pre_raw, self.H_post_raw, self.H_res_raw))
        if needs_grad:
            if self._train_coeffs is not None:               # one autograd node for all layers of the host model
                return self._train_coeffs
            return _CoeffFn.apply(self, self.H_pre_raw, self.H_post_raw, self.H_res_raw)
        st = self._fresh_state()
        return st.h_pre, st.h_post, st.h_res

    # ------------------------------------------------------------------ inference token path (K2 kernels)
    def fused_supported(self) -> bool:
        d, h = self.input_dim, self.hidden_dim
        return (self.use_mixed_precision and d % 32 == 0 and h % 64 == 0 and (d <= 256 or d % 256 == 0)
                and (h <= 256 or h % 256 == 0) and (2 * h <= 256 or (2 * h) % 256 == 0))

    def _mlp_bf16(self):
        st = self._state
        w1, w2 = self.mlp[0].weight, self.mlp[3].weight
        key = (w1.data_ptr(), w1._version, w2.data_ptr(), w2._version)
        if st.wkey != key or (torch.is_grad_enabled() and torch.cuda.is_current_stream_capturing()):
            st.w1 = w1.detach().to(torch.bfloat16).contiguous()
            st.w2 = w2.detach().to(torch.bfloat16).contiguous()
            st.wkey = key
        return st.w1, st.w2

    def _forward_fused(self, x2: torch.Tensor, out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
        """x2 [T, D] fp32 / bf16 -> [T, D] (:248-267, eval mode: dropout is the identity)."""
        from . import _lib
        out_dtype = out_dtype or self.output_dtype or torch.float32
        st = self._fresh_state()
        w1, w2 = self._mlp_bf16()
        d = self.input_dim
        dp = _pad64(d)
        x2 = x2.contiguous()
        if self.use_chain_kernel and x2.dtype == torch.bfloat16 and ops.mhc_module_fwd_supported(d, self.hidden_dim):
            # small widths: the whole path in ONE kernel per token tile, no intermediate in HBM
            return ops.mhc_module_fwd(x2, st.h_pre_t, w1, self.mlp[0].bias.detach(), w2, self.mlp[3].bias.detach(), st.h_post_t,
                                      st.h_res_t, (self.norm_pre.weight.detach(), self.norm_pre.bias.detach(), self.norm_pre.eps),
                                      (self.norm_post.weight.detach(), self.norm_post.bias.detach(), self.norm_post.eps), out_dtype)
        reuse = x2.dtype == torch.bfloat16 and dp == d           # bf16 activations: x itself is the A operand of x @ H_res
        xn, xb = ops.layernorm_fwd(x2, self.norm_pre.weight.detach(), self.norm_pre.bias.detach(), self.norm_pre.eps,
                                   out_dtype=torch.bfloat16, out_ld=dp, want_copy=not reuse, copy_ld=dp)   # :250
        if reuse:
            xb = x2
        z = ops.gemm_bf16(xn, st.h_pre_t)                                                                 # :253
        z = ops.gemm_bf16(z, w1, bias=self.mlp[0].bias.detach(), epilogue=_lib.HVS_GEMM_EPI_BIAS_GELU)     # :164-165
        z = ops.gemm_bf16(z, w2, bias=self.mlp[3].bias.detach(), epilogue=_lib.HVS_GEMM_EPI_BIAS_GELU)     # :167-168
        if d <= 512:
            return ops.gemm_bf16(z, st.h_post_t, xb, st.h_res_t, ln_weight=self.norm_post.weight.detach(),
                                 ln_bias=self.norm_post.bias.detach(), ln_eps=self.norm_post.eps,
                                 epilogue=_lib.HVS_GEMM_EPI_LAYERNORM, out_dtype=out_dtype)                # :259-267
        pre = ops.gemm_bf16(z, st.h_post_t, xb, st.h_res_t, out_dtype=torch.float32)
        out, _ = ops.layernorm_fwd(pre, self.norm_post.weight.detach(), self.norm_post.bias.detach(), self.norm_post.eps,
                                   out_dtype=out_dtype)
        return out

    def forward(self, x: torch.Tensor) -> torch.Tensor:  # :223-280
        shape = x.shape
        needs_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))
        if x.is_cuda and not needs_grad and not self.training and self.fused_supported() and x.dtype in (torch.float32, torch.bfloat16):
            return self._forward_fused(x.reshape(-1, shape[-1])).reshape(shape)
        if x.is_cuda and self.use_training_kernels and self.fused_supported() and x.dtype in (torch.float32, torch.bfloat16):
            return self._forward_training(x.reshape(-1, shape[-1])).reshape(shape)
        return self.forward_library(x)

    def _forward_training(self, x2: torch.Tensor) -> torch.Tensor:
        """Grad-enabled / train-mode forward on the library's own kernels (_K2TokenPathFn); dropout of the output and the
        signal-ratio monitor as in forward_library."""
        h_pre, h_post, h_res = self.constrained_matrices()
        out = _K2TokenPathFn.apply(self, x2, h_pre, h_post, h_res, self.mlp[0].weight, self.mlp[0].bias, self.mlp[3].weight,
                                   self.mlp[3].bias, self.norm_pre.weight, self.norm_pre.bias, self.norm_post.weight, self.norm_post.bias)
        out = self.dropout(out)
        if self.training and self.monitor_signal_ratio:
            with torch.no_grad():                        # :295-303, without the per-call eigvalsh: one fused kernel pair
                i = self.signal_ratio_idx % 1000
                ops.signal_ratio(out.detach(), x2.detach().contiguous(), self.signal_ratio_history[i:i + 1])
                self.signal_ratio_idx += 1
        return out

    def forward_library(self, x: torch.Tensor) -> torch.Tensor:
        """The token path in torch ops (library GEMMs under bf16 autocast, exactly the reference's CUDA execution):
        the training path (autograd sees it), the path for shapes the kernels do not take, and the baseline the fused
        path is benchmarked against."""
        shape = x.shape
        if x.dim() > 2:
            x = x.reshape(shape[0], -1, shape[-1])
        x_in = x
        h_pre, h_post, h_res = self.constrained_matrices()
        mixed = self.use_mixed_precision and x.is_cuda
        with torch.autocast("cuda", enabled=mixed, dtype=self.dtype):
            z = _layer_norm(self.norm_pre, x, torch.bfloat16 if mixed else torch.float32)
            z = torch.matmul(z, h_pre)
            z = self.mlp(z)
            z = torch.matmul(z, h_post)
            out = torch.matmul(x_in, h_res) + z
            out = _layer_norm(self.norm_post, out, torch.float32)
            out = self.dropout(out)
        if self.training and self.monitor_signal_ratio:
            with torch.no_grad():                        # :295-303, without the per-call eigvalsh
                ratio = torch.norm(out.float(), dim=-1).mean() / (torch.norm(x_in.float(), dim=-1).mean() + 1e-8)
                self.signal_ratio_history[self.signal_ratio_idx % 1000] = ratio
                self.signal_ratio_idx += 1
        return out.reshape(shape)

    def record_gradient_norms(self):
        """Fill the reference's ``gradient_norms`` buffer (:153) from the current .grad of the three raw matrices
        (call after backward; one small device op, no host sync)."""
        with torch.no_grad():
            for i, p in enumerate((self.H_pre_raw, self.H_post_raw, self.H_res_raw)):
                if p.grad is not None:
                    self.gradient_norms[i] = p.grad.norm()

    def get_stability_metrics(self) -> Dict[str, Any]:   # :318-341 (values as _monitor_stability :282-316 computes them)
        with torch.no_grad():
            st = self._fresh_state()
            h = st.h_res.detach().float()
            eig = torch.linalg.eigvalsh((h + h.T) / 2)
            self.eigenvalues.copy_(eig)
            metrics = {"max_eigenvalue": eig.max().item(), "min_eigenvalue": eig.min().item(),
                       "eigenvalue_range": (eig.max() - eig.min()).item(),
                       "sk_convergence": self.sinkhorn.get_convergence_metrics(),
                       "row_sum_error": (h.sum(1).mean() - 1).abs().item(),
                       "col_sum_error": (h.sum(0).mean() - 1).abs().item()}
            if self.signal_ratio_idx > 0:
                v = self.signal_ratio_history[:min(self.signal_ratio_idx, 1000)]
                metrics.update({"signal_ratio_mean": v.mean().item(), "signal_ratio_std": v.std().item() if len(v) > 1 else 0.0,
                                "signal_ratio_min": v.min().item(), "signal_ratio_max": v.max().item(),
                                "signal_ratio": v[-1].item()})
        return metrics

    def extra_repr(self) -> str:
        return (f"input_dim={self.input_dim}, hidden_dim={self.hidden_dim}, expansion={self.expansion_rate}, "
                f"alpha={self.alpha}")
