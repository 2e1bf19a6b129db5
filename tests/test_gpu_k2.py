"""K2, the reference-literal ManifoldHyperConnection (src/models/manifold_layers.py:104-346) on the CUDA kernels:
batched static coefficients (+ their backward), row norms, the tcgen05 GEMMs with fused epilogues, and the module's
fused inference path -- each through the C ABI, against the CPU oracle / golden vectors.

Tolerances: fp32 coefficient work 1e-5 relative (north_star); the GEMM path follows the reference's CUDA-autocast
convention (bf16 operands, fp32 accumulate), so it is compared (i) tightly with a torch fp32 reference fed the SAME
bf16-rounded operands and (ii) with the fp32 oracle at a bf16-operand tolerance (2^-7 of the LayerNorm-ed output scale)."""
import numpy as np
import pytest
import torch

from oracle import mhc_ref

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rand(*shape, seed=0, std=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * std


# ----------------------------------------------------------------------------- row norms
def test_rmsnorm_kernel_golden_and_backward(golden):
    import hvs_b200
    g = golden("rmsnorm")
    x, scale, want = (torch.from_numpy(g[k]) for k in ("x", "scale", "out"))
    mod = hvs_b200.RMSNorm(2048).to(DEV)
    with torch.no_grad():
        mod.scale.copy_(scale)
        got = mod(x.to(DEV))
    assert ((got.cpu() - want).abs() / want.abs().clamp_min(1e-3)).max() < 1e-5
    # bf16 data, fp32 statistics
    xb = x.to(torch.bfloat16)
    with torch.no_grad():
        gb = mod(xb.to(DEV))
    ref = mhc_ref.rms_norm(xb.float(), scale)
    assert gb.dtype == torch.bfloat16 and (gb.cpu().float() - ref).abs().max() <= 2.0 ** -7 * ref.abs().max()
    # backward vs autograd through the oracle; deterministic
    for rows, dim in ((37, 256), (1000, 2048), (3, 96)):
        xx = _rand(rows, dim, seed=rows).requires_grad_(True)
        sc = (1 + 0.1 * _rand(dim, seed=dim)).requires_grad_(True)
        dy = _rand(rows, dim, seed=7)
        mhc_ref.rms_norm(xx, sc).backward(dy)
        m2 = hvs_b200.RMSNorm(dim).to(DEV)
        with torch.no_grad():
            m2.scale.copy_(sc.detach())
        xg = xx.detach().to(DEV).requires_grad_(True)
        m2(xg).backward(dy.to(DEV))
        assert torch.allclose(xg.grad.cpu(), xx.grad, rtol=1e-4, atol=1e-5)
        assert torch.allclose(m2.scale.grad.cpu(), sc.grad, rtol=1e-4, atol=1e-4)
        dx2, ds2 = hvs_b200.ops.rmsnorm_bwd(xx.detach().to(DEV), sc.detach().to(DEV), dy.to(DEV))
        assert torch.equal(dx2, xg.grad) and torch.equal(ds2, m2.scale.grad)


def test_layernorm_kernel_with_padding_and_copy():
    import hvs_b200
    for rows, dim, ld in ((100, 32, 64), (7, 512, 512), (33, 1792, 1792), (5, 96, 128)):
        x = _rand(rows, dim, seed=dim) * 3 + 1.5
        w, b = 1 + 0.1 * _rand(dim, seed=1), 0.1 * _rand(dim, seed=2)
        want = torch.nn.functional.layer_norm(x, (dim,), w, b)
        out, cp = hvs_b200.ops.layernorm_fwd(x.to(DEV), w.to(DEV), b.to(DEV), 1e-5, torch.float32, ld, True, ld)
        assert torch.allclose(out[:, :dim].cpu(), want, rtol=1e-5, atol=2e-6)
        assert (out[:, dim:] == 0).all() and (cp[:, dim:] == 0).all()
        assert torch.equal(cp[:, :dim].cpu(), x.to(torch.bfloat16))
        ob, _ = hvs_b200.ops.layernorm_fwd(x.to(DEV), w.to(DEV), b.to(DEV), 1e-5, torch.bfloat16, ld)
        assert torch.equal(ob[:, :dim].cpu(), out[:, :dim].cpu().to(torch.bfloat16))


@pytest.mark.parametrize("rows,dim", [(4099, 32), (1000, 64), (333, 128), (77, 256), (130, 512), (1, 32)])
def test_layernorm_small_row_kernel(rows, dim):
    """The vectorised D/8-lanes-per-row kernel (dense bf16 output, D in {32...512}) for bf16 and fp32 inputs."""
    import hvs_b200
    w, b = (1 + 0.1 * _rand(dim, seed=1)).to(DEV), (0.1 * _rand(dim, seed=2)).to(DEV)
    for dt in (torch.bfloat16, torch.float32):
        x = (_rand(rows, dim, seed=rows) * 2 + 0.7).to(dt)
        want = torch.nn.functional.layer_norm(x.float(), (dim,), w.cpu(), b.cpu())
        out, cp = hvs_b200.ops.layernorm_fwd(x.to(DEV), w, b, 1e-5, torch.bfloat16, dim, True, dim)
        assert (out.cpu().float() - want).abs().max() <= 2.0 ** -8 * want.abs().max() + 1e-6
        assert torch.equal(out.cpu(), want.to(torch.bfloat16)) or (out.cpu().float() - want).abs().max() < 0.02
        assert torch.equal(cp.cpu(), x.to(torch.bfloat16))
        out2, none = hvs_b200.ops.layernorm_fwd(x.to(DEV), w, b, 1e-5, torch.bfloat16, dim)
        assert none is None and torch.equal(out2, out)


@pytest.mark.parametrize("rows,dim", [(4099, 32), (1000, 64), (333, 128), (77, 256), (130, 512)])
def test_layernorm_backward_kernel(rows, dim):
    import hvs_b200
    for xdt, gdt in ((torch.float32, torch.float32), (torch.bfloat16, torch.float32), (torch.bfloat16, torch.bfloat16)):
        x = (_rand(rows, dim, seed=rows) * 2 + 0.7).to(xdt)
        w, dy = 1 + 0.1 * _rand(dim, seed=1), _rand(rows, dim, seed=3).to(gdt)
        xl, wl, bl = x.float().requires_grad_(True), w.clone().requires_grad_(True), torch.zeros(dim, requires_grad=True)
        torch.nn.functional.layer_norm(xl, (dim,), wl, bl).backward(dy.float())
        dx, dw, db = hvs_b200.ops.layernorm_bwd(x.to(DEV), w.to(DEV), dy.to(DEV))
        tol = 2.0 ** -7 if xdt == torch.bfloat16 else 1e-5
        assert (dx.cpu().float() - xl.grad).abs().max() <= tol * xl.grad.abs().max() + 1e-6
        assert torch.allclose(dw.cpu(), wl.grad, rtol=1e-4, atol=1e-3) and torch.allclose(db.cpu(), bl.grad, rtol=1e-4, atol=1e-3)
        dx2, dw2, _ = hvs_b200.ops.layernorm_bwd(x.to(DEV), w.to(DEV), dy.to(DEV))
        assert torch.equal(dx, dx2) and torch.equal(dw, dw2)
    # fp32 output from the small-row forward kernel
    out, _ = hvs_b200.ops.layernorm_fwd(x.to(DEV), w.to(DEV), torch.zeros(dim, device=DEV), 1e-5, torch.float32, dim)
    assert torch.allclose(out.cpu(), torch.nn.functional.layer_norm(x.float(), (dim,), w, torch.zeros(dim)), rtol=1e-5, atol=2e-6)


def test_gemm_partial_k_stage_is_zero_filled():
    """K = 32 / 96 / 160 (not multiples of the 64-element stage): TMA zero-fills the rest of the box on both operands."""
    import hvs_b200
    for m, n, k in ((300, 64, 32), (129, 96, 96), (1000, 256, 160)):
        a, b = _rand(m, k, seed=k).to(torch.bfloat16), (_rand(n, k, seed=n) / k ** 0.5).to(torch.bfloat16)
        out = hvs_b200.ops.gemm_bf16(a.to(DEV), b.to(DEV), out_dtype=torch.float32).cpu()
        ref = a.double() @ b.double().t()
        assert (out.double() - ref).abs().max() <= 2e-5 * ref.abs().max() + 1e-6, (m, n, k)


# ----------------------------------------------------------------------------- batched static coefficients
def _jobs(dims, iters=20, hidden_mult=2, seed=0, std=0.3):
    import hvs_b200
    jobs, raws = [], []
    for i, d in enumerate(dims):
        h = max(64, d * hidden_mult)
        dp = (d + 63) // 64 * 64
        raw = [_rand(d, h, seed=seed + 3 * i, std=std), _rand(h, d, seed=seed + 3 * i + 1, std=std), _rand(d, d, seed=seed + 3 * i + 2, std=std)]
        raws.append(raw)
        f32, bf = dict(dtype=torch.float32, device=DEV), dict(dtype=torch.bfloat16, device=DEV)
        jobs.append({"h_pre_raw": raw[0].to(DEV), "h_post_raw": raw[1].to(DEV), "h_res_raw": raw[2].to(DEV),
                     "h_pre": torch.empty(d, h, **f32), "h_post": torch.empty(h, d, **f32), "h_res": torch.empty(d, d, **f32),
                     "h_pre_t": torch.full((h, dp), 7.0, **bf), "h_post_t": torch.empty(d, h, **bf), "h_res_t": torch.full((d, dp), 7.0, **bf),
                     "uv_history": torch.empty(iters + 1, 2, d, **f32), "convergence": torch.empty(iters, **f32)})
    return jobs, raws


def test_static_coeffs_all_layers_one_launch_vs_oracle():
    """Every D the model has, including the 1024 and 1792 the single-CTA path never tested, in ONE
    launch: 1e-5 relative to the oracle, row / column sums, transposed bf16 copies, convergence history."""
    import hvs_b200
    dims = [32, 64, 128, 256, 256, 512, 1024, 1792, 3, 48]
    jobs, raws = _jobs(dims)
    before = hvs_b200._lib.launch_count()
    hvs_b200.ops.static_coeffs(jobs, 20, 1e-8)
    torch.cuda.synchronize()
    assert hvs_b200._lib.launch_count() == before + 1
    for d, j, raw in zip(dims, jobs, raws):
        hp, hq, hr = mhc_ref.constrained_matrices(*raw)
        want, hist = mhc_ref.sinkhorn_knopp(raw[2], 20, return_history=True)
        got = j["h_res"].cpu()
        assert ((got - hr).abs() / hr.abs()).max() < 1e-5, d
        assert (got.sum(0) - 1).abs().max() < 1e-4 and (got.sum(1) - 1).abs().max() < 1e-4, d
        assert ((j["h_pre"].cpu() - hp).abs() / hp.abs()).max() < 1e-5 and ((j["h_post"].cpu() - hq).abs() / hq.abs()).max() < 1e-5
        assert torch.allclose(j["convergence"].cpu(), hist, atol=3e-6), d
        assert torch.equal(j["h_res_t"][:, :d].cpu(), got.t().to(torch.bfloat16))
        assert torch.equal(j["h_pre_t"][:, :d].cpu(), j["h_pre"].cpu().t().to(torch.bfloat16))
        assert torch.equal(j["h_post_t"].cpu(), j["h_post"].cpu().t().to(torch.bfloat16))
        assert (j["h_res_t"][:, d:] == 0).all() and (j["h_pre_t"][:, d:] == 0).all()
        # the scalings reproduce the projection: P = diag(u_n) K diag(v_n)
        k = torch.softmax(raw[2], -1) * d
        uv = j["uv_history"].cpu()
        assert torch.allclose(uv[20, 0][:, None] * k * uv[20, 1][None, :], got, rtol=1e-5, atol=1e-9)
    jobs2, _ = _jobs(dims)
    hvs_b200.ops.static_coeffs(jobs2, 20, 1e-8)
    for a, b in zip(jobs, jobs2):
        assert torch.equal(a["h_res"], b["h_res"]) and torch.equal(a["convergence"], b["convergence"])


def test_static_coeffs_hot_logits_and_iteration_counts():
    import hvs_b200
    for iters in (0, 1, 5, 20):
        jobs, raws = _jobs([64, 200], iters=iters, std=1.0, seed=11)
        hvs_b200.ops.static_coeffs(jobs, iters, 1e-8)
        for j, raw in zip(jobs, raws):
            want = mhc_ref.sinkhorn_knopp(raw[2], iters)
            assert ((j["h_res"].cpu() - want).abs() / want.abs()).max() < 1e-5, iters


@pytest.mark.parametrize("dims", [[16, 64], [256], [1024, 32]])
def test_static_coeffs_backward_vs_oracle_autograd(dims):
    import hvs_b200
    jobs, raws = _jobs(dims, seed=5, std=0.5)
    hvs_b200.ops.static_coeffs(jobs, 20, 1e-8)
    grads, want = [], []
    for d, j, raw in zip(dims, jobs, raws):
        h = raw[0].shape[1]
        g = [_rand(d, h, seed=d + 1), _rand(h, d, seed=d + 2), _rand(d, d, seed=d + 3)]
        leaf = [r.clone().requires_grad_(True) for r in raw]
        hp, hq, hr = mhc_ref.constrained_matrices(*leaf)
        (hp * g[0]).sum().backward(); (hq * g[1]).sum().backward(); (hr * g[2]).sum().backward()
        want.append([t.grad for t in leaf])
        grads.append({"d_h_pre": g[0].to(DEV), "d_h_post": g[1].to(DEV), "d_h_res": g[2].to(DEV),
                      "d_h_pre_raw": torch.empty(d, h, device=DEV), "d_h_post_raw": torch.empty(h, d, device=DEV),
                      "d_h_res_raw": torch.empty(d, d, device=DEV)})
    hvs_b200.ops.static_coeffs_bwd(jobs, grads, 20, 1e-8)
    for d, gr, w in zip(dims, grads, want):
        for name, ref in zip(("d_h_pre_raw", "d_h_post_raw", "d_h_res_raw"), w):
            got = gr[name].cpu().double()
            rel = ((got - ref.double()).norm() / ref.double().norm()).item()
            assert rel < 1e-5, (d, name, rel)
        assert (gr["d_h_res_raw"].cpu() - w[2]).abs().max() <= 1e-4 * w[2].abs().max() + 1e-9


def test_sinkhorn_module_square_grad_uses_kernel_backward():
    import hvs_b200
    sk = hvs_b200.SinkhornKnoppProjection(20).to(DEV)
    w = torch.nn.Parameter(_rand(48, 48, seed=9, std=0.4).to(DEV))
    tgt = _rand(48, 48, seed=10).to(DEV)
    before = hvs_b200._lib.launch_count()
    (sk(w) * tgt).sum().backward()
    assert hvs_b200._lib.launch_count() == before + 2              # one forward, one backward launch: nothing unrolled
    leaf = w.detach().cpu().clone().requires_grad_(True)
    (mhc_ref.sinkhorn_knopp(leaf, 20) * tgt.cpu()).sum().backward()
    assert ((w.grad.cpu() - leaf.grad).norm() / leaf.grad.norm()) < 1e-5


# ----------------------------------------------------------------------------- tcgen05 GEMM + epilogues
def _gemm_ref(a0, b0, a1=None, b1=None, bias=None):
    acc = a0.double() @ b0.double().t()
    if a1 is not None:
        acc = acc + a1.double() @ b1.double().t()
    if bias is not None:
        acc = acc + bias.double()
    return acc


@pytest.mark.parametrize("m,n,k", [(128, 256, 64), (1, 32, 64), (300, 96, 128), (1000, 512, 2048), (4099, 1792, 256),
                                   (257, 2048, 512), (129, 64, 4096)])
def test_gemm_plain_and_gelu(m, n, k):
    import hvs_b200
    from hvs_b200 import _lib
    a = _rand(m, k, seed=m).to(torch.bfloat16)
    b = (_rand(n, k, seed=n) / k ** 0.5).to(torch.bfloat16)
    bias = 0.1 * _rand(n, seed=3)
    out = hvs_b200.ops.gemm_bf16(a.to(DEV), b.to(DEV), out_dtype=torch.float32).cpu()
    ref = _gemm_ref(a, b)
    assert (out.double() - ref).abs().max() <= 2e-5 * ref.abs().max() + 1e-6        # fp32 accumulation of exact bf16 products
    outb = hvs_b200.ops.gemm_bf16(a.to(DEV), b.to(DEV)).cpu()
    assert outb.dtype == torch.bfloat16 and torch.equal(outb, out.to(torch.bfloat16))
    g = hvs_b200.ops.gemm_bf16(a.to(DEV), b.to(DEV), bias=bias.to(DEV), epilogue=_lib.HVS_GEMM_EPI_BIAS_GELU, out_dtype=torch.float32).cpu()
    gref = torch.nn.functional.gelu(_gemm_ref(a, b, bias=bias))
    assert (g.double() - gref).abs().max() <= 2e-5 * gref.abs().max() + 2e-6    # fp32-output path: erff
    gb = hvs_b200.ops.gemm_bf16(a.to(DEV), b.to(DEV), bias=bias.to(DEV), epilogue=_lib.HVS_GEMM_EPI_BIAS_GELU).cpu()
    assert (gb.float().double() - gref).abs().max() <= 2.0 ** -8 * gref.abs().max() + 4e-6   # bf16-output path: packed A&S erf (6e-7) + one rounding


@pytest.mark.parametrize("m,n,k0,k1", [(500, 64, 256, 64), (128, 512, 2048, 512), (1000, 256, 512, 256), (77, 32, 128, 64),
                                       (260, 96, 192, 128)])
def test_gemm_two_operand_pairs_layernorm_epilogue(m, n, k0, k1):
    import hvs_b200
    from hvs_b200 import _lib
    a0, b0 = _rand(m, k0, seed=1).to(torch.bfloat16), (_rand(n, k0, seed=2) / k0 ** 0.5).to(torch.bfloat16)
    a1, b1 = (_rand(m, k1, seed=3) + 0.5).to(torch.bfloat16), (_rand(n, k1, seed=4) / k1 ** 0.5).to(torch.bfloat16)
    w, b = 1 + 0.1 * _rand(n, seed=5), 0.1 * _rand(n, seed=6)
    pre = _gemm_ref(a0, b0, a1, b1)
    ref = torch.nn.functional.layer_norm(pre, (n,), w.double(), b.double())
    # padded A rows (row stride > K) exercise the lda path
    a1p = torch.zeros(m, k1 + 64, dtype=torch.bfloat16)
    a1p[:, :k1] = a1
    out = hvs_b200.ops.gemm_bf16(a0.to(DEV), b0.to(DEV), a1p.to(DEV)[:, :k1], b1.to(DEV), ln_weight=w.to(DEV), ln_bias=b.to(DEV),
                                 epilogue=_lib.HVS_GEMM_EPI_LAYERNORM, out_dtype=torch.float32).cpu()
    assert (out.double() - ref).abs().max() < 5e-5
    raw = hvs_b200.ops.gemm_bf16(a0.to(DEV), b0.to(DEV), a1.to(DEV), b1.to(DEV), out_dtype=torch.float32).cpu()
    assert (raw.double() - pre).abs().max() <= 2e-5 * pre.abs().max() + 1e-6
    again = hvs_b200.ops.gemm_bf16(a0.to(DEV), b0.to(DEV), a1p.to(DEV)[:, :k1], b1.to(DEV), ln_weight=w.to(DEV), ln_bias=b.to(DEV),
                                   epilogue=_lib.HVS_GEMM_EPI_LAYERNORM, out_dtype=torch.float32).cpu()
    assert torch.equal(out, again)


def test_gemm_rejects_bad_shapes():
    import hvs_b200
    from hvs_b200._lib import HvsError
    a, b = torch.zeros(8, 44, dtype=torch.bfloat16, device=DEV), torch.zeros(32, 44, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(HvsError):
        hvs_b200.ops.gemm_bf16(a, b)                                # K not a multiple of 8 (rows must be 16-byte multiples)
    with pytest.raises(HvsError):
        hvs_b200.ops.gemm_bf16(torch.zeros(8, 64, dtype=torch.bfloat16, device=DEV), torch.zeros(40, 64, dtype=torch.bfloat16, device=DEV))
    assert hvs_b200.ops.gemm_bf16(torch.zeros(0, 64, dtype=torch.bfloat16, device=DEV), torch.zeros(32, 64, dtype=torch.bfloat16, device=DEV)).shape == (0, 32)


# ----------------------------------------------------------------------------- the module's fused inference path
# Conditioning.  At the reference's initialisation (xavier, gain 0.1: H_pre ~ 0.5, H_post ~ 1, H_res ~ 1/D everywhere) the
# input of norm_post is a per-token CONSTANT plus a variation of relative size ~1e-2...1e-4, and LayerNorm divides by
# that variation.  Under the reference's own CUDA convention (torch.cuda.amp.autocast(bfloat16), :248) the matrices are
# rounded to bf16 (relative step 2^-8), which is as large as the variation: the module's output is then dominated by
# rounding and ANY two correct bf16 implementations (this one, cuBLAS under autocast) differ by O(1) after the norm.
# So parity is asserted (a) on the LayerNorm INPUT at bf16-operand accuracy for the reference's init, (b) on the module
# output for well-conditioned ("trained-like", raw std 1) coefficients, and (c) on the output for the reference's init
# with the bound scaled by the row's condition number max|pre| / std(pre).
def _module_pair(d, n, hidden=None, seed=0, raw_std=None):
    import hvs_b200
    torch.manual_seed(seed)
    mod = hvs_b200.ManifoldHyperConnection(d, expansion_rate=n, hidden_dim=hidden)
    with torch.no_grad():                                           # non-trivial norms / biases
        for p in (mod.norm_pre.weight, mod.norm_post.weight):
            p.add_(0.1 * torch.randn_like(p))
        for p in (mod.norm_pre.bias, mod.norm_post.bias, mod.mlp[0].bias, mod.mlp[3].bias):
            p.add_(0.05 * torch.randn_like(p))
        if raw_std is not None:
            for p in (mod.H_pre_raw, mod.H_post_raw, mod.H_res_raw):
                p.normal_(0, raw_std)
    params = {k: v.detach().clone() for k, v in mod.state_dict().items()}
    return mod.to(DEV).eval(), params


def _oracle_pre_and_out(x, p, bf16_operands):
    """LayerNorm input and module output of the reference arithmetic (:248-267), in fp64; with bf16_operands the
    reference's CUDA-autocast convention (LN fp32 -> bf16, every matmul operand and intermediate activation bf16)."""
    r = (lambda t: t.to(torch.bfloat16).double()) if bf16_operands else (lambda t: t.double())
    hp, hq, hr = mhc_ref.constrained_matrices(p["H_pre_raw"], p["H_post_raw"], p["H_res_raw"])
    d = x.shape[-1]
    xn = r(torch.nn.functional.layer_norm(x.float(), (d,), p["norm_pre.weight"], p["norm_pre.bias"]))
    z = r((xn @ r(hp)).float())
    z = r(torch.nn.functional.gelu(z @ r(p["mlp.0.weight"]).t() + p["mlp.0.bias"].double()).float())
    z = r(torch.nn.functional.gelu(z @ r(p["mlp.3.weight"]).t() + p["mlp.3.bias"].double()).float())
    pre = z @ r(hq) + r(x) @ r(hr)
    return pre, torch.nn.functional.layer_norm(pre, (d,), p["norm_post.weight"].double(), p["norm_post.bias"].double())


def _fused_pre(mod, x):
    """The fused path's LayerNorm input (same launches as _forward_fused with the last epilogue switched off)."""
    import hvs_b200
    from hvs_b200 import _lib
    from hvs_b200.mhc import _pad64
    st = mod._fresh_state()
    w1, w2 = mod._mlp_bf16()
    dp = _pad64(mod.input_dim)
    xn, xb = hvs_b200.ops.layernorm_fwd(x, mod.norm_pre.weight.detach(), mod.norm_pre.bias.detach(), mod.norm_pre.eps,
                                        out_dtype=torch.bfloat16, out_ld=dp, want_copy=True, copy_ld=dp)
    z = hvs_b200.ops.gemm_bf16(xn, st.h_pre_t)
    z = hvs_b200.ops.gemm_bf16(z, w1, bias=mod.mlp[0].bias.detach(), epilogue=_lib.HVS_GEMM_EPI_BIAS_GELU)
    z = hvs_b200.ops.gemm_bf16(z, w2, bias=mod.mlp[3].bias.detach(), epilogue=_lib.HVS_GEMM_EPI_BIAS_GELU)
    return hvs_b200.ops.gemm_bf16(z, st.h_post_t, xb, st.h_res_t, out_dtype=torch.float32)


SHAPES = [(64, 4, None, 1000), (32, 4, None, 4099), (256, 2, None, 401), (512, 4, None, 300), (256, 2, 1024, 257),
          (1024, 2, None, 130), (1792, 2, None, 1), (128, 4, None, 77)]


@pytest.mark.parametrize("d,n,hidden,t", SHAPES)
def test_module_fused_forward_trained_like_coefficients(d, n, hidden, t):
    import hvs_b200
    mod, p = _module_pair(d, n, hidden, seed=d + n, raw_std=1.0)
    assert mod.fused_supported()
    x = _rand(t, d, seed=t) * 1.5 + 0.2
    before = hvs_b200._lib.launch_count()
    with torch.no_grad():
        y = mod(x.to(DEV))
    launches = hvs_b200._lib.launch_count() - before
    assert launches == (7 if d > 512 else 6)                        # coefficients (1, first call only) + LN + 4 GEMMs (+ LN)
    assert y.dtype == torch.float32 and y.shape == x.shape
    _, tight = _oracle_pre_and_out(x, p, True)
    _, exact = _oracle_pre_and_out(x, p, False)
    e_tight = (y.cpu().double() - tight).abs()
    e_exact = (y.cpu().double() - exact).abs()
    print(f"[k2 trained-like d={d} n={n} t={t}] vs bf16-convention ref max {e_tight.max():.2e} mean {e_tight.mean():.2e}; "
          f"vs fp32 oracle max {e_exact.max():.2e} mean {e_exact.mean():.2e}")
    assert e_tight.max() < 3e-2 and e_tight.mean() < 2e-3          # same operands: only accumulation order / 1-ulp re-rounding
    assert e_exact.max() < 2.0 ** -3 and e_exact.mean() < 2.0 ** -6   # bf16 operands vs the fp32 oracle, output scale ~1
    with torch.no_grad():
        y2 = mod(x.to(DEV))
    assert torch.equal(y, y2)
    assert hvs_b200._lib.launch_count() - before == launches + launches - 1      # cached coefficients: no second refresh


def _condition_magnitude(x, p):
    """sum_k |a_k| |b_k| of the two contractions that feed norm_post (z @ H_post + x @ H_res), bf16 convention: the
    magnitude at which operand rounding acts (the K2 analogue of K1's condition magnitude)."""
    r = lambda t: t.to(torch.bfloat16).double()
    hp, hq, hr = mhc_ref.constrained_matrices(p["H_pre_raw"], p["H_post_raw"], p["H_res_raw"])
    d = x.shape[-1]
    xn = r(torch.nn.functional.layer_norm(x.float(), (d,), p["norm_pre.weight"], p["norm_pre.bias"]))
    z = r((xn @ r(hp)).float())
    z = r(torch.nn.functional.gelu(z @ r(p["mlp.0.weight"]).t() + p["mlp.0.bias"].double()).float())
    z = r(torch.nn.functional.gelu(z @ r(p["mlp.3.weight"]).t() + p["mlp.3.bias"].double()).float())
    return z.abs() @ r(hq).abs() + r(x).abs() @ r(hr).abs()


@pytest.mark.parametrize("d,n,hidden,t", SHAPES)
def test_module_fused_forward_reference_init(d, n, hidden, t):
    """The reference's own initialisation (ill-conditioned, see the note above): the LayerNorm INPUT agrees with the
    bf16-convention arithmetic to 2^-6 of the condition magnitude (intermediate activations re-round to bf16, a 1-ulp
    flip there is 2^-9 relative), and the output within the bound scaled by the row's condition number."""
    mod, p = _module_pair(d, n, hidden, seed=d + n)
    x = _rand(t, d, seed=t) * 1.5 + 0.2
    with torch.no_grad():
        y = mod(x.to(DEV)).cpu().double()
        pre = _fused_pre(mod, x.to(DEV)).cpu().double()
    pre_c, out_c = _oracle_pre_and_out(x, p, True)
    pre_o, _ = _oracle_pre_and_out(x, p, False)
    cmag = _condition_magnitude(x, p)
    e_c = ((pre - pre_c).abs() / cmag).max().item()
    e_o = ((pre - pre_o).abs() / cmag).max().item()
    cond = pre_c.abs().amax(-1, keepdim=True) / pre_c.std(-1, keepdim=True).clamp_min(1e-30)
    out_err = ((y - out_c).abs() / cond).max().item()
    print(f"[k2 reference-init d={d} t={t}] LN input err / condition magnitude: vs convention {e_c:.2e}, vs fp32 oracle {e_o:.2e}; "
          f"median row condition {cond.median().item():.0f}; output err / condition {out_err:.2e}")
    assert e_c < 2.0 ** -6
    assert e_o < 0.25                                               # what bf16 operands cost at this init (H_pre = 0.5 +- 0.01 has ~3 significant bits of signal)
    assert ((y - out_c).abs() <= 6e-2 + 1e-4 * cond).all()


def test_module_fused_forward_golden_and_shapes(golden):
    import hvs_b200
    g = golden("mhc_module")
    for tag, (d, n) in {"d64n4": (64, 4), "d32n2": (32, 2)}.items():
        mod = hvs_b200.ManifoldHyperConnection(d, expansion_rate=n).to(DEV).eval()
        keys = {k.split("/p/")[1] for k in g.files if k.startswith(tag + "/p/")}
        p = {k: torch.from_numpy(g[f"{tag}/p/{k}"]) for k in keys}
        mod.load_state_dict(p)
        x = torch.from_numpy(g[f"{tag}/x"])
        x2 = x.reshape(-1, d)
        with torch.no_grad():
            y = mod(x.to(DEV))
            pre = _fused_pre(mod, x2.to(DEV)).cpu().double()
        assert y.shape == x.shape
        pre_o, out_o = _oracle_pre_and_out(x2, p, False)
        assert (out_o.float().reshape(x.shape) - torch.from_numpy(g[f"{tag}/y"])).abs().max() < 1e-4     # the fixture is this arithmetic
        pre_c, out_c = _oracle_pre_and_out(x2, p, True)
        cmag = _condition_magnitude(x2, p)
        assert ((pre - pre_c).abs() / cmag).max() < 2.0 ** -6 and ((pre - pre_o).abs() / cmag).max() < 0.25
        cond = pre_c.abs().amax(-1, keepdim=True) / pre_c.std(-1, keepdim=True)
        excess = ((y.cpu().double().reshape(-1, d) - out_c).abs() - 1e-4 * cond).max().item()
        assert excess < 0.15, excess                # bf16 re-rounding floor (one flipped ulp of an intermediate, amplified by the norm) + conditioning
        hr = mod.constrained_matrices()[2].cpu()
        assert ((hr - torch.from_numpy(g[f"{tag}/H_res"])).abs() / torch.from_numpy(g[f"{tag}/H_res"])).max() < 1e-5
    # bf16 input, empty input, bf16 output on request
    mod, p = _module_pair(64, 4, raw_std=1.0)
    with torch.no_grad():
        xb = (_rand(50, 64, seed=1)).to(torch.bfloat16)
        yb = mod(xb.to(DEV))
        assert yb.dtype == torch.float32 and (yb.cpu() - mhc_ref.mhc_module_forward(xb.float(), p)).abs().max() < 2.0 ** -3
        assert mod(torch.zeros(0, 64, device=DEV)).shape == (0, 64)
        mod.output_dtype = torch.bfloat16
        y16 = mod(xb.to(DEV))
        assert y16.dtype == torch.bfloat16 and torch.equal(y16, yb.to(torch.bfloat16))
        lib = mod.forward_library(xb.to(DEV))                       # the reference's own CUDA execution (library GEMMs under autocast)
        assert (lib.float() - yb).abs().max() < 0.3 and (lib.float() - yb).abs().mean() < 5e-2   # two bf16 pipelines: isolated 1-ulp flips, amplified by the norm


@pytest.mark.parametrize("d,n", [(32, 4), (64, 4)])
def test_module_single_kernel_chain_path(d, n):
    """hvs_mhc_module_fwd: the whole token path in ONE launch (no intermediate in HBM) for (D, H) = (32, 128) / (64, 256)
    and bf16 input -- against the five-launch path on the same input (same operands, same K order: differences only
    from the LayerNorm statistics' summation order) and against the oracle at the bf16-operand tolerance."""
    import hvs_b200
    mod, p = _module_pair(d, n, seed=7 + d, raw_std=1.0)
    assert hvs_b200.ops.mhc_module_fwd_supported(d, mod.hidden_dim)
    for t in (1, 127, 128, 129, 1000, 40000):
        x = (_rand(t, d, seed=t) * 1.5 + 0.2).to(torch.bfloat16)
        with torch.no_grad():
            mod.use_chain_kernel = True
            mod(x[:1].to(DEV))
            before = hvs_b200._lib.launch_count()
            y = mod(x.to(DEV))
            assert hvs_b200._lib.launch_count() - before == 1           # one kernel for the whole module
            mod.use_chain_kernel = False
            y5 = mod(x.to(DEV))
            mod.use_chain_kernel = True
            y_again = mod(x.to(DEV))
        assert y.dtype == torch.float32 and y.shape == (t, d)
        assert torch.equal(y, y_again)
        diff = (y - y5).abs()
        assert diff.max() < 2e-2 and diff.mean() < 1e-4, (t, diff.max().item(), diff.mean().item())
        exact = mhc_ref.mhc_module_forward(x.float(), p)
        err = (y.cpu() - exact).abs()
        assert err.max() < 2.0 ** -3 and err.mean() < 2.0 ** -6, (t, err.max().item(), err.mean().item())
    with torch.no_grad():
        mod.output_dtype = torch.bfloat16
        yb = mod(x.to(DEV))
        assert yb.dtype == torch.bfloat16 and torch.equal(yb, y.to(torch.bfloat16))


def test_refresh_static_coefficients_batches_all_modules():
    import hvs_b200
    from hvs_b200.mhc import refresh_static_coefficients
    torch.manual_seed(0)
    net = torch.nn.ModuleList([hvs_b200.ManifoldHyperConnection(d, expansion_rate=2) for d in (32, 64, 256, 512, 1792)]).to(DEV).eval()
    before = hvs_b200._lib.launch_count()
    assert refresh_static_coefficients(net) == 5
    assert hvs_b200._lib.launch_count() == before + 1               # five layers, one launch
    assert refresh_static_coefficients(net) == 0
    with torch.no_grad():
        net[1].H_res_raw.mul_(1.5)
    assert refresh_static_coefficients(net) == 1
    for m in net:
        hp, hq, hr = mhc_ref.constrained_matrices(m.H_pre_raw.detach().cpu(), m.H_post_raw.detach().cpu(), m.H_res_raw.detach().cpu())
        got = m.constrained_matrices()
        assert ((got[2].cpu() - hr).abs() / hr).max() < 1e-5 and ((got[0].cpu() - hp).abs() / hp).max() < 1e-5


def test_module_training_path_gradients_vs_oracle():
    """The LIBRARY form of the training path (use_training_kernels = False, and every fp32 module): coefficient forward /
    backward on the kernels (hvs_mhc_static_coeffs[_bwd]), token path in torch ops; the kernel form is covered by
    tests/test_gpu_k2_train.py.  fp32 token path (use_mixed_precision=False) against autograd through the fp32 oracle at 1e-3; the bf16
    autocast path on trained-like coefficients at the bf16-operand tolerance."""
    import hvs_b200
    for mixed, raw_std, tol in ((False, 1.0, 5e-3), (True, 1.0, 1e-1)):
        mod, p = _module_pair(64, 4, seed=3, raw_std=raw_std)
        mod.use_mixed_precision = mixed
        mod.use_training_kernels = False
        mod.train()
        mod.dropout.p = 0.0
        for m in mod.mlp:
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
        x = _rand(200, 64, seed=8)
        dy = _rand(200, 64, seed=9)
        xg = x.to(DEV).requires_grad_(True)
        before = hvs_b200._lib.launch_count()
        mod(xg).backward(dy.to(DEV))
        assert hvs_b200._lib.launch_count() - before == 2 + 2 + 4   # coefficients fwd + bwd (nothing unrolled), 2 LayerNorm fwd, 2 x 2 bwd
        grad_keys = ("H_pre_raw", "H_post_raw", "H_res_raw", "mlp.0.weight", "mlp.0.bias", "mlp.3.weight", "mlp.3.bias",
                     "norm_pre.weight", "norm_pre.bias", "norm_post.weight", "norm_post.bias")
        leaf = {k: (v.clone().requires_grad_(True) if k in grad_keys else v) for k, v in p.items()}
        xl = x.clone().requires_grad_(True)
        mhc_ref.mhc_module_forward(xl, leaf).backward(dy)

        def rel(a, b):
            return ((a.cpu().double() - b.double()).norm() / b.double().norm()).item()
        got = dict(mod.named_parameters())
        errs = {k: rel(got[k].grad, leaf[k].grad) for k in grad_keys}
        errs["x"] = rel(xg.grad, xl.grad)
        print(f"[k2 training mixed={mixed} raw_std={raw_std}] " + " ".join(f"{k}:{v:.1e}" for k, v in errs.items()))
        assert max(errs.values()) < tol, errs
        mod.record_gradient_norms()
        assert torch.allclose(mod.gradient_norms.cpu(), torch.stack([mod.H_pre_raw.grad.norm(), mod.H_post_raw.grad.norm(), mod.H_res_raw.grad.norm()]).cpu())
