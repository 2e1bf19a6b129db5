"""The algebra both K1 kernels rely on, pinned on the CPU against the oracle (reference arithmetic,
manifold_layers.py:56-77):

* forward: iterating the SCALINGS  u_k = 1 / (K v_{k-1} + eps),  v_k = 1 / (K^T u_k + eps),  P = diag(u) K diag(v)
  with K = m * softmax(logits)  equals the reference's  P <- P / (row_sum + eps);  P <- P / (col_sum + eps);
* backward: the reverse sweep over that scaling form (what `mhc_stream_bwd_fused_kernel`'s coefficient warps run:
  no reciprocals, no reconstruction of P) equals autograd through the reference iteration."""
import numpy as np
import pytest
import torch

from oracle import mhc_ref


def scaling_forward(logits: np.ndarray, iters: int, eps: float):
    """[B,4,4] float64 -> P, K, u history [iters+1,B,4], v history (slot 0 = ones)."""
    z = logits - logits.max(-1, keepdims=True)
    e = np.exp(z)
    k = 4.0 * e / e.sum(-1, keepdims=True)
    b = logits.shape[0]
    us, vs = [np.ones((b, 4))], [np.ones((b, 4))]
    for _ in range(iters):
        u = 1.0 / (np.einsum("bij,bj->bi", k, vs[-1]) + eps)
        v = 1.0 / (np.einsum("bij,bi->bj", k, u) + eps)
        us.append(u); vs.append(v)
    p = us[-1][:, :, None] * k * vs[-1][:, None, :]
    return p, k, np.stack(us), np.stack(vs)


def reverse_sweep(g: np.ndarray, k: np.ndarray, us: np.ndarray, vs: np.ndarray):
    """dL/dK for dL/dP = g, in the sign convention of the kernel (tn = vb v^2 = -tb, ubn = -ub)."""
    u, v = us[-1], vs[-1]
    gk = g * k
    ubn = -np.einsum("bij,bj->bi", gk, v)
    vb = np.einsum("bij,bi->bj", gk, u)
    dk = g * u[:, :, None] * v[:, None, :]
    for it in range(us.shape[0] - 1, 0, -1):
        u, v, vprev = us[it], vs[it], vs[it - 1]
        tn = vb * v * v
        ubn = ubn + np.einsum("bij,bj->bi", k, tn)
        dk = dk - u[:, :, None] * tn[:, None, :]
        sb = ubn * u * u
        ubn = np.zeros_like(ubn)
        vb = np.einsum("bij,bi->bj", k, sb)
        dk = dk + sb[:, :, None] * vprev[:, None, :]
    return dk


@pytest.mark.parametrize("iters", [0, 1, 5, 20])
@pytest.mark.parametrize("std", [0.01, 0.5, 2.0])
def test_scaling_form_equals_reference_iteration(iters, std):
    g = torch.Generator().manual_seed(iters * 7 + int(std * 100))
    logits = torch.randn(257, 4, 4, generator=g, dtype=torch.float64) * std
    ref = mhc_ref.sinkhorn_knopp(logits, iters, 1e-8).numpy()
    p, _, _, _ = scaling_forward(logits.numpy(), iters, 1e-8)
    # the two differ only in where eps enters (eps vs eps / u): ~1e-8 for the benchmark's logits, < 1e-6 for hot ones
    # (small u), an order of magnitude inside north_star's 1e-5
    assert np.abs(p - ref).max() <= (1e-6 if std >= 2.0 else 1e-7) * np.abs(ref).max()


@pytest.mark.parametrize("iters", [0, 1, 5, 20])
def test_reverse_sweep_equals_autograd(iters):
    g = torch.Generator().manual_seed(100 + iters)
    logits = (torch.randn(129, 4, 4, generator=g, dtype=torch.float64) * 0.7).requires_grad_(True)
    up = torch.randn(129, 4, 4, generator=g, dtype=torch.float64)
    (mhc_ref.sinkhorn_knopp(logits, iters, 1e-8) * up).sum().backward()
    want = logits.grad.numpy()
    p, k, us, vs = scaling_forward(logits.detach().numpy(), iters, 1e-8)
    dk = reverse_sweep(up.numpy(), k, us, vs)
    # K = 4 softmax(logits): d logits = K (dK - sum_j(dK K) / 4), row by row
    dl = k * (dk - (dk * k).sum(-1, keepdims=True) / 4.0)
    assert np.abs(dl - want).max() <= 1e-6 * max(np.abs(want).max(), 1e-30)
