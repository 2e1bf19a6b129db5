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
FUSED_BWD_MAX_ITERS = 24      # the single-pass backward keeps every iteration's scalings in shared memory


class _StreamMHCFn(torch.autograd.Function):
    """Training path.  The forward additionally writes 112 B/token of statistics (un-normalised projection and
    sum of squares); the backward is then ONE fused kernel (dx and every parameter gradient in a single pass over
    x and dy).  With more than 24 Sinkhorn iterations the two-kernel backward that recomputes everything is used."""

    @staticmethod
    def forward(ctx, x, phi, bias, alpha, scale, sk_iters, eps_rms, eps_sk):
        fused = sk_iters <= FUSED_BWD_MAX_ITERS
        saved = ops.new_saved(x) if fused else None
        y, _, _ = ops.mhc_stream_fwd(x, phi, bias, alpha, scale, sk_iters, eps_rms, eps_sk, saved=saved)
        if fused:
            ctx.save_for_backward(x, phi, bias, alpha, scale, saved)
        else:
            ctx.save_for_backward(x, phi, bias, alpha, scale)
        ctx.cfg = (sk_iters, eps_rms, eps_sk, fused)
        return y

    @staticmethod
    def backward(ctx, dy):
        sk_iters, eps_rms, eps_sk, fused = ctx.cfg
        if fused:
            x, phi, bias, alpha, scale, saved = ctx.saved_tensors
            g = ops.mhc_stream_bwd_saved(x, dy.contiguous(), saved, phi, bias, alpha, scale, sk_iters, eps_rms, eps_sk)
        else:
            x, phi, bias, alpha, scale = ctx.saved_tensors
            g = ops.mhc_stream_bwd(x, dy.contiguous(), phi, bias, alpha, scale, sk_iters, eps_rms, eps_sk)
        return g["dx"], g["dphi"], g["dbias"], g["dalpha"], g["dscale"], None, None, None


class StreamMHC(nn.Module):
    """Stream mHC residual layer:  y = H_res x + H_post (x) fn(H_pre^T x).

    x: [..., n, C] bf16 (n = 4, C = 512 in this build).  ``fn=None`` is the identity (one fused kernel,
    differentiable through the fused backward).  With a wrapped layer ``fn`` the forward runs the
    coefficient kernel, ``fn`` on the bf16 layer input, and the mixing kernel (inference path).
    """

    def __init__(self, n_streams: int = 4, channels: int = 512, alpha: float = 0.01, sk_iterations: int = 20,
                 eps: float = 1e-8, phi_std: float = 0.02, fn: Optional[Callable[[torch.Tensor], torch.Tensor]] = None,
                 device=None):
        super().__init__()
        n, c = n_streams, channels
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
                                      self.eps, self.eps, want_y=False, want_coeffs=True)
        return co[:, :n], co[:, n:2 * n], co[:, 2 * n:].reshape(t, n, n)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        shape = x.shape
        xf = x.reshape(-1, self.n_streams, self.channels).contiguous()
        if self.fn is None:
            y = _StreamMHCFn.apply(xf, self.phi, self.bias, self.alpha, self.rms_scale, self.sk_iterations,
                                   self.eps, self.eps)
        else:
            if torch.is_grad_enabled() and (xf.requires_grad or self.phi.requires_grad):
                raise HvsError("StreamMHC with a wrapped fn is forward-only in this build (fused backward covers fn=None)")
            _, u, co = ops.mhc_stream_fwd(xf, self.phi, self.bias, self.alpha, self.rms_scale, self.sk_iterations,
                                          self.eps, self.eps, want_y=False, want_u=True, want_coeffs=True)
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
    t = x_host.shape[0]
    n, c = layer.n_streams, layer.channels
    nchunks = (t + chunk_tokens - 1) // chunk_tokens
    s_in, s_cmp, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    bufs = [{k: torch.empty((chunk_tokens, n, c), dtype=torch.bfloat16, device=dev) for k in ("x", "dy", "y", "dx")}
            for _ in range(2)]
    fused = layer.sk_iterations <= FUSED_BWD_MAX_ITERS
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
    return out


# ----------------------------------------------------------------------------- reference-signature modules
class SinkhornKnoppProjection(nn.Module):
    """SinkhornKnoppProjection(num_iterations=20, epsilon=1e-8, tau=1.0)  (manifold_layers.py:25-101).
    ``forward`` runs hvs_sinkhorn; the ``convergence_history`` buffer is filled on the device without
    the reference's 3 host synchronisations per iteration."""

    def __init__(self, num_iterations: int = 20, epsilon: float = 1e-8, tau: float = 1.0):
        super().__init__()
        self.num_iterations, self.epsilon, self.tau = num_iterations, epsilon, tau
        self.register_buffer("convergence_history", torch.zeros(num_iterations))

    def forward(self, matrix: torch.Tensor, return_history: bool = False):
        if torch.is_grad_enabled() and matrix.requires_grad:
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
    """Differentiable D x D projection for TRAINING of the reference-literal module: same arithmetic in
    torch device ops so autograd can unroll it.  (The fused CUDA backward exists for the per-token K1
    path; a D x D backward kernel is listed as next work in DESIGN.md.)"""
    squeeze = matrix.dim() == 2
    p = matrix.unsqueeze(0) if squeeze else matrix
    m = p.shape[-1]
    p = torch.softmax(p / tau, dim=-1) * m
    for _ in range(iters):
        p = p / (p.sum(dim=-1, keepdim=True) + eps)
        p = p / (p.sum(dim=-2, keepdim=True) + eps)
    return p.squeeze(0) if squeeze else p


class RMSNorm(nn.Module):
    """RMSNorm(dim, eps=1e-8) (manifold_layers.py:437-456)."""

    def __init__(self, dim: int, eps: float = 1e-8):
        super().__init__()
        self.scale = nn.Parameter(torch.ones(dim))
        self.eps = eps

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        rms = torch.sqrt(torch.mean(x.pow(2), dim=-1, keepdim=True) + self.eps)
        return x / rms * self.scale


class ManifoldHyperConnection(nn.Module):
    """Drop-in for the reference ManifoldHyperConnection (manifold_layers.py:104-346).

    Same constructor, parameters, buffers and state_dict keys (H_pre_raw, H_post_raw, H_res_raw,
    gradient_norms, eigenvalues, signal_ratio_history, sinkhorn.convergence_history, mlp.{0,3}.*,
    norm_pre.*, norm_post.*).  Differences, all behind the same interface:
      * the constrained matrices come from hvs_mhc_constrained_matrices and are cached until a
        parameter changes (the reference recomputes them, with 60 host syncs, on every forward);
      * stability monitoring (:282-316) is computed on demand in get_stability_metrics instead of an
        eigvalsh inside every training forward.
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
        self._cache = None
        self._cache_key = None
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

    def constrained_matrices(self) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:      # :205-221
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in (self.H_pre_raw, self.H_post_raw, self.H_res_raw))
        if needs_grad:
            return (torch.sigmoid(self.H_pre_raw), 2 * torch.sigmoid(self.H_post_raw), self.sinkhorn(self.H_res_raw))
        key = self._key()
        if self._cache is None or self._cache_key != key:
            self._cache = ops.constrained_matrices(self.H_pre_raw.detach(), self.H_post_raw.detach(), self.H_res_raw.detach(),
                                                   self.sinkhorn.num_iterations, self.sinkhorn.epsilon,
                                                   history=self.sinkhorn.convergence_history)
            self._cache_key = key
        return self._cache

    def forward(self, x: torch.Tensor) -> torch.Tensor:  # :223-280
        shape = x.shape
        if x.dim() > 2:
            x = x.reshape(shape[0], -1, shape[-1])
        x_in = x
        h_pre, h_post, h_res = self.constrained_matrices()
        with torch.autocast("cuda", enabled=self.use_mixed_precision and x.is_cuda, dtype=self.dtype):
            z = self.norm_pre(x)
            z = torch.matmul(z, h_pre)
            z = self.mlp(z)
            z = torch.matmul(z, h_post)
            out = torch.matmul(x_in, h_res) + z
            out = self.norm_post(out)
            out = self.dropout(out)
        if self.training:
            with torch.no_grad():                        # :295-303, without the per-call eigvalsh
                ratio = torch.norm(out.float(), dim=-1).mean() / (torch.norm(x_in.float(), dim=-1).mean() + 1e-8)
                self.signal_ratio_history[self.signal_ratio_idx % 1000] = ratio
                self.signal_ratio_idx += 1
        return out.reshape(shape)

    def get_stability_metrics(self) -> Dict[str, Any]:   # :318-341
        with torch.no_grad():
            _, _, h_res = self.constrained_matrices()
            h = h_res.detach().float()
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
