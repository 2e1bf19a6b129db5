"""CPU restatement of the mHC arithmetic (TEST INFRASTRUCTURE, see oracle/__init__.py).

Every function names the reference lines it follows
(/root/reference/src/models/manifold_layers.py unless stated otherwise).
Plain torch fp32 ops on CPU tensors; runs on the GPU box without the
reference checkout.

Two contracts:

* K1 "stream mHC" -- ``stream_mhc_forward``: n residual streams of C channels
  per token, RMSNorm over the flattened n*C row, a small projection to
  n + n + n*n logits, sigmoid / 2*sigmoid gates, per-token Sinkhorn-Knopp on
  the n x n block, mixing of the streams.  Composed only of reference
  primitives: RMSNorm (:449-456), the gates (:213, :216) and the batched
  branch of SinkhornKnoppProjection.forward (:46-49, :56-83).
* K2 "module" -- ``mhc_module_forward``: the reference-literal
  ManifoldHyperConnection.forward (:223-280) in eval mode.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional, Tuple

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------
def sinkhorn_knopp(matrix: torch.Tensor, num_iterations: int = 20,
                   epsilon: float = 1e-8, tau: float = 1.0,
                   return_history: bool = False):
    """SinkhornKnoppProjection.forward (:32-93).

    ``[B,n,m]`` or ``[n,m]`` -> same shape.  Start from ``softmax(M/tau, -1) * m``
    (:56-57); then ``num_iterations`` times: divide every row by
    ``(row_sum + eps)`` (:66-67), divide every column by ``(col_sum + eps)``
    (:71-72).  The last operation is the column normalisation.  A 2-D input is
    treated as a batch of one (repair R1: the shipped 2-D branch leaves ``m``
    unbound, :50-57).
    ``return_history`` additionally returns the per-iteration
    ``|mean(row_sum) - 1|`` that the reference writes to its
    ``convergence_history`` buffer (:76-77).
    """
    squeeze = matrix.dim() == 2
    p = matrix.unsqueeze(0) if squeeze else matrix
    shape = p.shape
    n, m = shape[-2], shape[-1]
    p = p.reshape(-1, n, m)
    p = torch.softmax(p / tau, dim=-1) * m
    hist = []
    for _ in range(num_iterations):
        rs = p.sum(dim=2, keepdim=True)
        p = p / (rs + epsilon)
        cs = p.sum(dim=1, keepdim=True)
        p = p / (cs + epsilon)
        if return_history:
            hist.append((rs.mean() - 1.0).abs())
    p = p.reshape(shape)
    if squeeze:
        p = p.squeeze(0)
    if return_history:
        return p, torch.stack(hist) if hist else torch.zeros(0)
    return p


def rms_norm(x: torch.Tensor, scale: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """RMSNorm.forward (:449-456): ``x / sqrt(mean(x^2, -1) + eps) * scale``."""
    rms = torch.sqrt(torch.mean(x * x, dim=-1, keepdim=True) + eps)
    return x / rms * scale


def constrained_matrices(h_pre_raw: torch.Tensor, h_post_raw: torch.Tensor,
                         h_res_raw: torch.Tensor, sk_iterations: int = 20
                         ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """ManifoldHyperConnection.constrained_matrices (:205-221)."""
    return (torch.sigmoid(h_pre_raw),
            2.0 * torch.sigmoid(h_post_raw),
            sinkhorn_knopp(h_res_raw, sk_iterations))


# --------------------------------------------------------------------------
# K2: reference-literal module forward (eval mode: dropout is the identity)
# --------------------------------------------------------------------------
def mhc_module_forward(x: torch.Tensor, p: Dict[str, torch.Tensor],
                       sk_iterations: int = 20) -> torch.Tensor:
    """ManifoldHyperConnection.forward (:223-280), eval mode, fp32.

    ``p`` uses the reference's state_dict keys: H_pre_raw, H_post_raw,
    H_res_raw, mlp.0.{weight,bias}, mlp.3.{weight,bias}, norm_pre.{weight,bias},
    norm_post.{weight,bias}.  Input ``[T,D]`` or ``[B,*,D]`` (flattened to
    ``[B,-1,D]`` and restored, :233-239, :277-278).
    """
    shape = x.shape
    d = shape[-1]
    if x.dim() > 2:
        x = x.reshape(shape[0], -1, d)
    h_pre, h_post, h_res = constrained_matrices(
        p["H_pre_raw"], p["H_post_raw"], p["H_res_raw"], sk_iterations)
    z = F.layer_norm(x, (d,), p["norm_pre.weight"], p["norm_pre.bias"])      # :250
    z = z @ h_pre                                                            # :253
    z = F.gelu(F.linear(z, p["mlp.0.weight"], p["mlp.0.bias"]))              # :164-165
    z = F.gelu(F.linear(z, p["mlp.3.weight"], p["mlp.3.bias"]))              # :167-168
    z = z @ h_post                                                           # :259
    out = x @ h_res + z                                                      # :263-264
    out = F.layer_norm(out, (d,), p["norm_post.weight"], p["norm_post.bias"])  # :267
    return out.reshape(shape)


# --------------------------------------------------------------------------
# K1: stream mHC
# --------------------------------------------------------------------------
def bf16_round(t: torch.Tensor) -> torch.Tensor:
    """Round-to-nearest-even to bf16, returned as fp32."""
    return t.to(torch.bfloat16).to(torch.float32)


def stream_mhc_coeffs(x: torch.Tensor, phi: torch.Tensor, bias: torch.Tensor,
                      alpha: torch.Tensor, rms_scale: torch.Tensor,
                      sk_iterations: int = 20, eps: float = 1e-8,
                      split_phi: bool = False, dtype: torch.dtype = torch.float32):
    """Coefficient stage of the stream kernel.

    x          [T,n,C] bf16 (or fp32 holding bf16 values)
    phi        [n*C, n*n+2n] fp32    projection
    bias       [n*n+2n] fp32
    alpha      [3] fp32              per-group logit scale (pre, post, res);
                                     init 0.01 = the reference's ``alpha`` (:134)
    rms_scale  [n*C] fp32            RMSNorm gain (:446)

    Numeric convention (DESIGN.md "operand convention"): the input row is
    exact bf16; RMSNorm statistics, gates, Sinkhorn and accumulation are fp32;
    the projection operand ``rms_scale * phi`` is rounded to bf16 -- the
    reference runs every matmul of the layer under
    ``torch.cuda.amp.autocast(dtype=bfloat16)`` (:248), i.e. with bf16 operands
    and fp32 accumulation.  ``split_phi=True`` keeps the fp32 operand instead
    (what the kernel's two-term hi+lo mode reproduces).

    Returns (H_pre [T,n], H_post [T,n], H_res [T,n,n], logits [T,n*n+2n]), fp32.
    """
    t, n, c = x.shape
    xf = x.to(dtype).reshape(t, n * c)
    inv_rms = 1.0 / torch.sqrt(torch.mean(xf * xf, dim=-1, keepdim=True) + eps)  # :451
    w = rms_scale.to(dtype)[:, None] * phi.to(dtype)
    if not split_phi:
        # bf16 rounding of the operand with a straight-through gradient kept in `dtype` (a plain
        # .to(bf16).to(dtype) would also round the GRADIENT to bf16 on its way back).  r - w is exact
        # (Sterbenz), so w + (r - w) == r bit for bit.
        w = w + (w.to(torch.bfloat16).to(dtype) - w).detach()
    raw = (xf @ w) * inv_rms                                  # == rms_norm(x) @ phi
    alpha = alpha.to(dtype)
    a = torch.cat([alpha[0].expand(n), alpha[1].expand(n), alpha[2].expand(n * n)])
    logits = raw * a + bias.to(dtype)
    h_pre = torch.sigmoid(logits[:, :n])                       # :213
    h_post = 2.0 * torch.sigmoid(logits[:, n:2 * n])           # :216
    h_res = sinkhorn_knopp(logits[:, 2 * n:].reshape(t, n, n), sk_iterations)  # :219, batched branch
    return h_pre, h_post, h_res, logits


def stream_mhc_forward(x: torch.Tensor, phi: torch.Tensor, bias: torch.Tensor,
                       alpha: torch.Tensor, rms_scale: torch.Tensor,
                       sk_iterations: int = 20, eps: float = 1e-8,
                       fn: Optional[Callable[[torch.Tensor], torch.Tensor]] = None,
                       split_phi: bool = False, round_output: bool = True,
                       dtype: torch.dtype = torch.float32):
    """Full stream-mHC layer:  y = H_res x + H_post (x) fn(H_pre^T x).

    ``fn=None`` is the identity (the microbenchmark of BASELINE.json config 2);
    then the layer input ``u`` is consumed in fp32.  With a real ``fn`` the
    layer input is rounded to bf16 first (it crosses HBM as bf16) and so is
    ``fn``'s result.  ``y`` is rounded once to bf16 (``round_output``).

    Returns dict(y [T,n,C], u [T,C] fp32, H_pre, H_post, H_res).
    """
    t, n, c = x.shape
    h_pre, h_post, h_res, logits = stream_mhc_coeffs(
        x, phi, bias, alpha, rms_scale, sk_iterations, eps, split_phi, dtype)
    xs = x.to(dtype)
    u = torch.einsum("tj,tjc->tc", h_pre, xs)
    if fn is None:
        fu = u
    else:
        fu = fn(u.to(torch.bfloat16)).to(dtype)
    y = torch.einsum("tij,tjc->tic", h_res, xs) + h_post[:, :, None] * fu[:, None, :]
    if round_output:
        y = y.to(torch.bfloat16)
    return {"y": y, "u": u, "H_pre": h_pre, "H_post": h_post, "H_res": h_res,
            "logits": logits}


def stream_mhc_backward(x: torch.Tensor, dy: torch.Tensor, phi: torch.Tensor,
                        bias: torch.Tensor, alpha: torch.Tensor,
                        rms_scale: torch.Tensor, sk_iterations: int = 20,
                        eps: float = 1e-8, split_phi: bool = False):
    """Gradients of ``stream_mhc_forward`` (fn = identity) by autograd through
    the fp32 restatement.  Returns dict(dx [T,n,C] fp32, dphi, dbias, dalpha,
    dscale)."""
    xl = x.to(torch.float32).detach().requires_grad_(True)
    phil = phi.detach().clone().requires_grad_(True)
    bl = bias.detach().clone().requires_grad_(True)
    al = alpha.detach().clone().requires_grad_(True)
    sl = rms_scale.detach().clone().requires_grad_(True)
    out = stream_mhc_forward(xl, phil, bl, al, sl, sk_iterations, eps,
                             None, split_phi, round_output=False)
    out["y"].backward(dy.to(torch.float32))
    return {"dx": xl.grad, "dphi": phil.grad, "dbias": bl.grad,
            "dalpha": al.grad, "dscale": sl.grad}


def bf16_ulp(mag: torch.Tensor) -> torch.Tensor:
    """Size of one bf16 unit in the last place at magnitude ``mag`` (fp32)."""
    mag = mag.abs().clamp_min(torch.finfo(torch.float32).tiny)
    e = torch.floor(torch.log2(mag))
    return torch.pow(2.0, e - 7.0)


def mixing_condition_magnitude(x: torch.Tensor, h_pre, h_post, h_res) -> torch.Tensor:
    """sum_j |M_ij| |x_j| with M = H_res + H_post H_pre^T -- the magnitude at
    which the 2-bf16-ulp output bound is measured (raw ulp
    counts explode where a mean of zero-mean streams cancels to ~0)."""
    m = h_res + h_post[:, :, None] * h_pre[:, None, :]
    return torch.einsum("tij,tjc->tic", m.abs(), x.to(torch.float32).abs())
