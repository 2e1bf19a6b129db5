"""Parity of the K1 stream-mHC CUDA path (through the C ABI) against the CPU oracle and the golden
vectors.  Tolerances are north_star's: fp32 coefficients within 1e-5 relative, row/column sums within
1e-4 of 1, bf16 outputs within 2 bf16 ulp (measured at the condition magnitude sum_j |M_ij||x_j|)."""
import numpy as np
import pytest
import torch

from oracle import mhc_ref

pytestmark = pytest.mark.gpu

COEF_RTOL = 1e-5
ULP_BOUND = 2.0


def bits_to_bf16(a):
    return torch.from_numpy(a.astype(np.int16)).view(torch.bfloat16)


def make_inputs(t, seed=0, alpha=0.01, phistd=0.02, bstd=0.0, n=4, c=512):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(t, n, c, generator=g).to(torch.bfloat16)
    k = n * n + 2 * n
    phi = torch.randn(n * c, k, generator=g) * phistd
    bias = torch.randn(k, generator=g) * bstd
    al = torch.full((3,), alpha)
    scale = 1.0 + 0.05 * torch.randn(n * c, generator=g)
    return x, phi, bias, al, scale


def run_gpu(x, phi, bias, al, scale, **kw):
    import hvs_b200
    dev = "cuda:0"
    y, u, co = hvs_b200.ops.mhc_stream_fwd(x.to(dev), phi.to(dev), bias.to(dev), al.to(dev), scale.to(dev),
                                          want_y=True, want_u=True, want_coeffs=True, **kw)
    torch.cuda.synchronize()
    return y.cpu(), u.cpu(), co.cpu()


def check_against_oracle(x, phi, bias, al, scale, y, u, co, ref=None):
    t, n, c = x.shape
    ref = ref or mhc_ref.stream_mhc_forward(x, phi, bias, al, scale)
    h_pre, h_post, h_res = co[:, :n], co[:, n:2 * n], co[:, 2 * n:].reshape(t, n, n)
    for name, got, want in (("H_pre", h_pre, ref["H_pre"]), ("H_post", h_post, ref["H_post"]), ("H_res", h_res, ref["H_res"])):
        rel = ((got - want).abs() / want.abs()).max().item()
        assert rel < COEF_RTOL, f"{name} rel err {rel:.2e}"
    assert (h_res.sum(-1) - 1).abs().max() < 1e-4 and (h_res.sum(-2) - 1).abs().max() < 1e-4
    assert (h_res >= 0).all()
    mag = mhc_ref.mixing_condition_magnitude(x, ref["H_pre"], ref["H_post"], ref["H_res"])
    y32 = torch.einsum("tij,tjc->tic", ref["H_res"], x.float()) + ref["H_post"][:, :, None] * ref["u"][:, None, :]
    ulps = ((y.float() - y32).abs() / mhc_ref.bf16_ulp(mag)).max().item()
    assert ulps <= ULP_BOUND, f"y off by {ulps:.2f} bf16 ulp at condition magnitude"
    umag = torch.einsum("tj,tjc->tc", ref["H_pre"].abs(), x.float().abs())
    uulps = ((u.float() - ref["u"]).abs() / mhc_ref.bf16_ulp(umag)).max().item()
    assert uulps <= ULP_BOUND, f"u off by {uulps:.2f} bf16 ulp"
    return ulps


@pytest.mark.parametrize("t", [1, 15, 16, 17, 96, 1000, 4099])
def test_forward_matches_oracle(t):
    inp = make_inputs(t, seed=t)
    y, u, co = run_gpu(*inp)
    check_against_oracle(*inp, y, u, co)


def test_forward_hot_logits():
    # larger logits: the Sinkhorn iteration has real work to do; row sums are reported, not assumed
    inp = make_inputs(257, seed=5, alpha=0.3, phistd=0.05, bstd=0.2)
    y, u, co = run_gpu(*inp)
    check_against_oracle(*inp, y, u, co)


def test_forward_golden(golden):
    g = golden("stream_mhc")
    for tag in ("n4c512", "n4c512_hot"):
        x = bits_to_bf16(g[f"{tag}/x_bits"])
        phi, bias = torch.from_numpy(g[f"{tag}/phi"]), torch.from_numpy(g[f"{tag}/bias"])
        al, scale = torch.from_numpy(g[f"{tag}/alpha"]), torch.from_numpy(g[f"{tag}/scale"])
        y, u, co = run_gpu(x, phi, bias, al, scale)
        ref = {"H_pre": torch.from_numpy(g[f"{tag}/H_pre"]), "H_post": torch.from_numpy(g[f"{tag}/H_post"]),
               "H_res": torch.from_numpy(g[f"{tag}/H_res"]), "u": torch.from_numpy(g[f"{tag}/u"])}
        check_against_oracle(x, phi, bias, al, scale, y, u, co, ref=ref)


def test_zero_iterations_and_empty():
    import hvs_b200
    inp = make_inputs(33, seed=9)
    y, u, co = run_gpu(*inp, sk_iters=0)
    ref = mhc_ref.stream_mhc_forward(*inp, sk_iterations=0)
    assert ((co[:, 8:].reshape(33, 4, 4) - ref["H_res"]).abs() / ref["H_res"]).max() < COEF_RTOL
    x = torch.empty(0, 4, 512, dtype=torch.bfloat16, device="cuda:0")
    y0, _, _ = hvs_b200.ops.mhc_stream_fwd(x, inp[1].cuda(), inp[2].cuda(), inp[3].cuda(), inp[4].cuda())
    assert y0.shape == (0, 4, 512)


@pytest.mark.parametrize("n,c", [(2, 64), (2, 256), (2, 1024), (4, 64), (4, 128), (4, 256), (4, 1024), (4, 512), (2, 40)])
@pytest.mark.parametrize("split", [False, True])
def test_forward_general_shapes_and_split_phi(n, c, split):
    """Every stream shape the model's call sites have (n in {2, 4}, C from 32 to 1024) on the general
    forward kernel, and HVS_MHC_SPLIT_PHI (fp32-accurate projection operand) for all shapes incl. n = 4, C = 512."""
    if (n, c) == (4, 512) and not split:
        pytest.skip("the tuned kernel: covered above")
    import hvs_b200
    t = 333
    x, phi, bias, al, scale = make_inputs(t, seed=n * c + split, alpha=0.3, phistd=0.03, bstd=0.1, n=n, c=c)
    y, u, co = run_gpu(x, phi, bias, al, scale, split_phi=split)
    ref = mhc_ref.stream_mhc_forward(x, phi, bias, al, scale, split_phi=split)
    check_against_oracle(x, phi, bias, al, scale, y, u, co, ref=ref)
    # wrapped-layer path: coefficients + mixing kernel with fu = u reproduces y
    y2 = hvs_b200.ops.mhc_stream_post(x.cuda(), co.cuda(), u.cuda())
    mag = mhc_ref.mixing_condition_magnitude(x, ref["H_pre"], ref["H_post"], ref["H_res"])
    assert ((y2.cpu().float() - y.float()).abs() <= 2 * mhc_ref.bf16_ulp(mag)).all()
    layer = hvs_b200.StreamMHC(n_streams=n, channels=c, split_phi=split, device="cuda")
    with torch.no_grad():
        assert layer(x.cuda()).shape == x.shape
    if split:                                           # the fp32-accurate operand is forward-only
        with pytest.raises(hvs_b200.HvsError, match="training kernels"):
            layer(x.cuda().requires_grad_(True))


def test_deterministic_and_unsupported_shape():
    import hvs_b200
    inp = make_inputs(2048, seed=3)
    a = run_gpu(*inp)
    b = run_gpu(*inp)
    assert torch.equal(a[0].view(torch.int16), b[0].view(torch.int16)) and torch.equal(a[2], b[2])
    x = torch.zeros(4, 3, 256, dtype=torch.bfloat16, device="cuda:0")          # n = 3: neither kernel takes it
    with pytest.raises(hvs_b200.HvsError):
        hvs_b200.ops.mhc_stream_fwd(x, torch.zeros(768, 15, device="cuda:0"), torch.zeros(15, device="cuda:0"),
                                    torch.zeros(3, device="cuda:0"), torch.ones(768, device="cuda:0"))
    x = torch.zeros(4, 4, 2048, dtype=torch.bfloat16, device="cuda:0")         # C > 1024
    with pytest.raises(hvs_b200.HvsError):
        hvs_b200.ops.mhc_stream_fwd(x, torch.zeros(8192, 24, device="cuda:0"), torch.zeros(24, device="cuda:0"),
                                    torch.zeros(3, device="cuda:0"), torch.ones(8192, device="cuda:0"))


def test_split_pre_post_equals_fused():
    """coefficients + post kernel with fu = bf16(u) reproduces the wrapped-layer form."""
    import hvs_b200
    inp = make_inputs(300, seed=11)
    dev = "cuda:0"
    xd = inp[0].to(dev)
    _, u, co = hvs_b200.ops.mhc_stream_fwd(xd, *(p.to(dev) for p in inp[1:]), want_y=False, want_u=True, want_coeffs=True)
    y = hvs_b200.ops.mhc_stream_post(xd, co, u).cpu()
    ref = mhc_ref.stream_mhc_forward(*inp, fn=lambda v: v)
    mag = mhc_ref.mixing_condition_magnitude(inp[0], ref["H_pre"], ref["H_post"], ref["H_res"])
    ulps = ((y.float() - ref["y"].float()).abs() / mhc_ref.bf16_ulp(mag)).max().item()
    assert ulps <= ULP_BOUND + 1.0   # fu itself carries one bf16 rounding


def test_large_linearity_property():
    """BASELINE config size (2^20 tokens would need 8.6 GB; 2^18 here): the layer is linear in x for fixed
    coefficients, so post(x, coeffs, 0) on a stream permutation equals the permuted output."""
    import hvs_b200
    t = 1 << 18
    g = torch.Generator(device="cuda:0").manual_seed(1)
    x = torch.randn(t, 4, 512, generator=g, device="cuda:0", dtype=torch.bfloat16)
    _, phi, bias, al, scale = make_inputs(1)
    y, _, co = hvs_b200.ops.mhc_stream_fwd(x, phi.cuda(), bias.cuda(), al.cuda(), scale.cuda(), want_coeffs=True)
    hres = co[:, 8:].reshape(t, 4, 4)
    assert (hres.sum(-1) - 1).abs().max() < 1e-4 and (hres.sum(-2) - 1).abs().max() < 1e-4
    # spot-check 64 random tokens against the oracle
    idx = torch.randint(0, t, (64,), generator=torch.Generator().manual_seed(2))
    xs = x[idx.cuda()].cpu()
    ref = mhc_ref.stream_mhc_forward(xs, phi, bias, al, scale)
    mag = mhc_ref.mixing_condition_magnitude(xs, ref["H_pre"], ref["H_post"], ref["H_res"])
    y32 = torch.einsum("tij,tjc->tic", ref["H_res"], xs.float()) + ref["H_post"][:, :, None] * ref["u"][:, None, :]
    assert ((y[idx.cuda()].cpu().float() - y32).abs() <= ULP_BOUND * mhc_ref.bf16_ulp(mag)).all()


# ----------------------------------------------------------------------------- backward
def run_bwd(x, dy, phi, bias, al, scale, **kw):
    import hvs_b200
    dev = "cuda:0"
    out = hvs_b200.ops.mhc_stream_bwd(x.to(dev), dy.to(dev), phi.to(dev), bias.to(dev), al.to(dev), scale.to(dev), **kw)
    torch.cuda.synchronize()
    return {k: v.cpu() for k, v in out.items()}


def check_bwd(inp, dy, got, tag=""):
    x, phi, bias, al, scale = inp
    ref = mhc_ref.stream_mhc_backward(x, dy, phi, bias, al, scale)
    fwd = mhc_ref.stream_mhc_forward(x, phi, bias, al, scale)
    # dx: bf16, error measured at the condition magnitude of the dominant term  M^T dy
    m = fwd["H_res"] + fwd["H_post"][:, :, None] * fwd["H_pre"][:, None, :]
    mag = torch.einsum("tij,tic->tjc", m.abs(), dy.float().abs()) + ref["dx"].abs()
    ulps = ((got["dx"].float() - ref["dx"]).abs() / mhc_ref.bf16_ulp(mag)).max().item()
    assert ulps <= ULP_BOUND, f"{tag} dx off by {ulps:.2f} bf16 ulp"
    for name in ("dphi", "dbias", "dalpha", "dscale"):
        a, b = got[name].double(), ref[name].double()
        rel = ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
        assert rel < 3e-4, f"{tag} {name} relative error {rel:.2e}"
    return ulps


@pytest.mark.parametrize("t", [1, 7, 8, 9, 64, 1000, 2051])
def test_backward_matches_oracle(t):
    inp = make_inputs(t, seed=100 + t, alpha=0.3, phistd=0.03, bstd=0.1)
    dy = torch.randn(t, 4, 512, generator=torch.Generator().manual_seed(t)).to(torch.bfloat16)
    got = run_bwd(*inp[:1], dy, *inp[1:])
    check_bwd(inp, dy, got, f"T={t}")


@pytest.mark.parametrize("n,c", [(2, 64), (2, 256), (2, 1024), (4, 64), (4, 128), (4, 256), (4, 1024), (2, 40), (4, 8)])
def test_backward_general_shapes_match_oracle(n, c):
    """Every stream shape but (4, 512) trains on the general backward (mhc_stream_generic_bwd.cu: per-token kernel, the
    tcgen05 GEMM for dW = x^T E, finalize): same bounds as the tuned kernels, bitwise reproducible, and the nn.Module's
    autograd path goes through it."""
    import hvs_b200
    t = 777
    inp = make_inputs(t, seed=1000 * n + c, alpha=0.3, phistd=0.03, bstd=0.1, n=n, c=c)
    dy = torch.randn(t, n, c, generator=torch.Generator().manual_seed(c)).to(torch.bfloat16)
    got = run_bwd(*inp[:1], dy, *inp[1:])
    check_bwd(inp, dy, got, f"n={n} C={c}")
    again = run_bwd(*inp[:1], dy, *inp[1:])
    assert all(torch.equal(got[k], again[k]) for k in got)
    x, phi, bias, al, scale = inp
    layer = hvs_b200.StreamMHC(n_streams=n, channels=c, device="cuda:0")
    with torch.no_grad():
        layer.phi.copy_(phi); layer.bias.copy_(bias); layer.alpha.copy_(al); layer.rms_scale.copy_(scale)
    xg = x.to("cuda:0").requires_grad_(True)
    layer(xg).backward(dy.to("cuda:0"))
    assert torch.equal(xg.grad.cpu(), got["dx"]) and torch.equal(layer.phi.grad.cpu(), got["dphi"])
    assert torch.equal(layer.rms_scale.grad.cpu(), got["dscale"]) and torch.equal(layer.alpha.grad.cpu(), got["dalpha"])


def test_backward_general_shape_edges():
    import hvs_b200
    for t in (1, 5, 4099):
        inp = make_inputs(t, seed=t, alpha=0.2, n=2, c=128)
        dy = torch.randn(t, 2, 128, generator=torch.Generator().manual_seed(t)).to(torch.bfloat16)
        check_bwd(inp, dy, run_bwd(*inp[:1], dy, *inp[1:]), f"T={t}")
    inp = make_inputs(300, seed=3, alpha=0.2, n=4, c=256)
    dy = torch.randn(300, 4, 256, generator=torch.Generator().manual_seed(9)).to(torch.bfloat16)
    for iters in (0, 1, 24):
        got = run_bwd(*inp[:1], dy, *inp[1:], sk_iters=iters)
        ref = mhc_ref.stream_mhc_backward(inp[0], dy, *inp[1:], sk_iterations=iters)
        assert ((got["dphi"].double() - ref["dphi"].double()).norm() / ref["dphi"].double().norm()).item() < 3e-4
    with pytest.raises(hvs_b200.HvsError):
        run_bwd(*inp[:1], dy, *inp[1:], sk_iters=25)
    z = hvs_b200.ops.mhc_stream_bwd(torch.zeros(0, 2, 64, dtype=torch.bfloat16, device="cuda:0"), torch.zeros(0, 2, 64, dtype=torch.bfloat16, device="cuda:0"),
                                    torch.zeros(128, 8, device="cuda:0"), torch.zeros(8, device="cuda:0"), torch.zeros(3, device="cuda:0"), torch.ones(128, device="cuda:0"))
    assert z["dx"].shape == (0, 2, 64) and float(z["dphi"].abs().max()) == 0.0 and float(z["dbias"].abs().max()) == 0.0


def test_backward_init_scale_logits():
    inp = make_inputs(515, seed=21)          # alpha = 0.01, the microbenchmark's configuration
    dy = torch.randn(515, 4, 512, generator=torch.Generator().manual_seed(5)).to(torch.bfloat16)
    got = run_bwd(*inp[:1], dy, *inp[1:])
    check_bwd(inp, dy, got, "init")


def test_backward_deterministic_and_linear_in_dy():
    inp = make_inputs(3000, seed=33, alpha=0.2)
    g = torch.Generator().manual_seed(6)
    dy = torch.randn(3000, 4, 512, generator=g).to(torch.bfloat16)
    a = run_bwd(*inp[:1], dy, *inp[1:])
    b = run_bwd(*inp[:1], dy, *inp[1:])
    for k in a:
        assert torch.equal(a[k].view(torch.int16) if a[k].dtype == torch.bfloat16 else a[k], b[k].view(torch.int16) if b[k].dtype == torch.bfloat16 else b[k]), k
    # parameter gradients are linear in dy: doubling dy (exact in bf16) doubles them
    c = run_bwd(*inp[:1], (dy.float() * 2).to(torch.bfloat16), *inp[1:])
    for k in ("dphi", "dbias", "dalpha", "dscale"):
        assert torch.allclose(c[k], 2 * a[k], rtol=1e-4, atol=1e-6 * a[k].abs().max().item()), k


# ----------------------------------------------------------------------------- fused backward (saved statistics)
def run_bwd_saved(x, dy, phi, bias, al, scale, sk_iters=20):
    import hvs_b200
    dev = "cuda:0"
    xd = x.to(dev)
    P = [p.to(dev) for p in (phi, bias, al, scale)]
    saved = hvs_b200.ops.new_saved(xd)
    y, _, _ = hvs_b200.ops.mhc_stream_fwd(xd, *P, sk_iters=sk_iters, saved=saved)
    out = hvs_b200.ops.mhc_stream_bwd_saved(xd, dy.to(dev), saved, *P, sk_iters=sk_iters)
    torch.cuda.synchronize()
    res = {k: v.cpu() for k, v in out.items()}
    res["saved"] = saved.cpu()
    res["y"] = y.cpu()
    return res


def test_forward_saved_statistics():
    """The training forward writes raw = x . bf16(scale*phi) and sum x^2 (what the fused backward consumes)."""
    inp = make_inputs(300, seed=7, alpha=0.3, phistd=0.03, bstd=0.1)
    x, phi, bias, al, scale = inp
    got = run_bwd_saved(x, torch.zeros_like(x), phi, bias, al, scale)
    w = (phi * scale[:, None]).to(torch.bfloat16).double()
    raw = x.reshape(300, -1).double() @ w
    ss = (x.double() ** 2).sum((1, 2))
    assert torch.allclose(got["saved"][:, :24].double(), raw, rtol=1e-5, atol=1e-4)
    assert torch.allclose(got["saved"][:, 24].double(), ss, rtol=1e-5)
    assert (got["saved"][:, 25:] == 0).all()
    # the saved variant computes the same y as the plain forward
    import hvs_b200
    y0, _, _ = hvs_b200.ops.mhc_stream_fwd(x.cuda(), phi.cuda(), bias.cuda(), al.cuda(), scale.cuda())
    assert torch.equal(y0.cpu().view(torch.int16), got["y"].view(torch.int16))


@pytest.mark.parametrize("t", [1, 7, 8, 9, 24, 25, 64, 1000, 2051])
def test_fused_backward_matches_oracle(t):
    inp = make_inputs(t, seed=200 + t, alpha=0.3, phistd=0.03, bstd=0.1)
    dy = torch.randn(t, 4, 512, generator=torch.Generator().manual_seed(t)).to(torch.bfloat16)
    got = run_bwd_saved(*inp[:1], dy, *inp[1:])
    check_bwd(inp, dy, got, f"fused T={t}")


def test_fused_backward_warm_and_init_2ulp_hot_logits_looser_norm_bound():
    """Warm logits (std ~0.7: Sinkhorn still converges to 5e-7, the scaling-form reverse sweep is exercised far from
    the uniform matrix) and the microbenchmark's alpha = 0.01 configuration against the oracle; hot logits (std > 2,
    20 iterations leave a 3e-2 row error) against the oracle for the parameter gradients and against the two-kernel
    backward for dx -- there the bf16 rounding of e in the W e MMA, common to both kernels, exceeds the 2-ulp
    metric, which only accounts for the mixing term."""
    inp = make_inputs(777, seed=41, alpha=0.5, phistd=0.03, bstd=0.3)
    dy = torch.randn(777, 4, 512, generator=torch.Generator().manual_seed(9)).to(torch.bfloat16)
    check_bwd(inp, dy, run_bwd_saved(*inp[:1], dy, *inp[1:]), "fused warm")
    inp = make_inputs(515, seed=21)
    dy = torch.randn(515, 4, 512, generator=torch.Generator().manual_seed(5)).to(torch.bfloat16)
    check_bwd(inp, dy, run_bwd_saved(*inp[:1], dy, *inp[1:]), "fused init")
    inp = make_inputs(777, seed=41, alpha=1.0, phistd=0.05, bstd=0.5)
    x, phi, bias, al, scale = inp
    dy = torch.randn(777, 4, 512, generator=torch.Generator().manual_seed(9)).to(torch.bfloat16)
    got, two = run_bwd_saved(x, dy, phi, bias, al, scale), run_bwd(x, dy, phi, bias, al, scale)
    ref = mhc_ref.stream_mhc_backward(x, dy, phi, bias, al, scale)
    for name in ("dphi", "dbias", "dalpha", "dscale"):
        a, b = got[name].double(), ref[name].double()
        assert ((a - b).norm() / b.norm()).item() < 3e-4, name
    d = got["dx"].float() - two["dx"].float()
    assert (d.norm() / two["dx"].float().norm()).item() < 2.0 ** -8            # bf16 output rounding level
    assert d.abs().max().item() <= 2.0 ** -6 * two["dx"].float().abs().max().item()
    assert ((got["dx"].float() - ref["dx"]).norm() / ref["dx"].norm()).item() < 2.0 ** -7


@pytest.mark.parametrize("iters", [0, 1, 5, 20, 24])
def test_fused_backward_iteration_counts(iters):
    import hvs_b200
    inp = make_inputs(130, seed=50 + iters, alpha=0.4, phistd=0.03, bstd=0.2)
    x, phi, bias, al, scale = inp
    dy = torch.randn(130, 4, 512, generator=torch.Generator().manual_seed(3)).to(torch.bfloat16)
    got = run_bwd_saved(x, dy, phi, bias, al, scale, sk_iters=iters)
    old = hvs_b200.ops.mhc_stream_bwd(x.cuda(), dy.cuda(), phi.cuda(), bias.cuda(), al.cuda(), scale.cuda(), sk_iters=iters)
    for k in ("dphi", "dbias", "dalpha", "dscale"):
        a, b = got[k].double(), old[k].cpu().double()
        assert ((a - b).norm() / b.norm().clamp_min(1e-30)).item() < 3e-4, (iters, k)
    d = (got["dx"].float() - old["dx"].cpu().float()).abs().max().item()
    assert d <= 2 * 2.0 ** -7 * old["dx"].float().abs().max().item(), (iters, d)


def test_fused_backward_limits_and_determinism():
    import hvs_b200
    from hvs_b200._lib import HvsError
    inp = make_inputs(3000, seed=33, alpha=0.2)
    dy = torch.randn(3000, 4, 512, generator=torch.Generator().manual_seed(6)).to(torch.bfloat16)
    a = run_bwd_saved(*inp[:1], dy, *inp[1:])
    b = run_bwd_saved(*inp[:1], dy, *inp[1:])
    for k in ("dx", "dphi", "dbias", "dalpha", "dscale"):
        va = a[k].view(torch.int16) if a[k].dtype == torch.bfloat16 else a[k]
        vb = b[k].view(torch.int16) if b[k].dtype == torch.bfloat16 else b[k]
        assert torch.equal(va, vb), k
    x = inp[0].cuda()
    P = [p.cuda() for p in inp[1:]]
    saved = hvs_b200.ops.new_saved(x)
    with pytest.raises(HvsError):      # more iterations than the shared-memory history holds
        hvs_b200.ops.mhc_stream_bwd_saved(x, dy.cuda(), saved, *P, sk_iters=25)
    with pytest.raises(HvsError):      # the two-kernel backward has the same limit
        hvs_b200.ops.mhc_stream_bwd(x, dy.cuda(), *P, sk_iters=25)
    # T = 0 is a no-op that still writes zero parameter gradients
    e = hvs_b200.ops.mhc_stream_bwd_saved(x[:0], dy.cuda()[:0], saved[:0], *P)
    assert e["dx"].numel() == 0 and float(e["dphi"].abs().max()) == 0.0 and float(e["dbias"].abs().max()) == 0.0


def test_fused_backward_full_size_properties():
    """2^18 tokens: parameter gradients are linear in dy; dx of zero dy is exactly zero."""
    import hvs_b200
    t = 1 << 18
    g = torch.Generator(device="cuda:0").manual_seed(11)
    x = torch.randn(t, 4, 512, generator=g, device="cuda:0", dtype=torch.bfloat16)
    dy = torch.randn(t, 4, 512, generator=g, device="cuda:0", dtype=torch.bfloat16)
    _, phi, bias, al, scale = make_inputs(1)
    P = [p.cuda() for p in (phi, bias, al, scale)]
    saved = hvs_b200.ops.new_saved(x)
    hvs_b200.ops.mhc_stream_fwd(x, *P, saved=saved)
    a = hvs_b200.ops.mhc_stream_bwd_saved(x, dy, saved, *P)
    c = hvs_b200.ops.mhc_stream_bwd_saved(x, (dy.float() * 2).to(torch.bfloat16), saved, *P)
    for k in ("dphi", "dbias", "dalpha", "dscale"):
        assert torch.allclose(c[k], 2 * a[k], rtol=2e-4, atol=1e-6 * a[k].abs().max().item()), k
    z = hvs_b200.ops.mhc_stream_bwd_saved(x, torch.zeros_like(dy), saved, *P)
    assert float(z["dx"].float().abs().max()) == 0.0 and float(z["dphi"].abs().max()) == 0.0
    # spot-check against the two-kernel backward
    old = hvs_b200.ops.mhc_stream_bwd(x, dy, *P)
    assert ((a["dphi"] - old["dphi"]).norm() / old["dphi"].norm()).item() < 1e-4
    assert (a["dx"].float() - old["dx"].float()).abs().max().item() <= 2.0 ** -6 * old["dx"].float().abs().max().item()


def test_kernel_timing_hooks():
    """hvs_mhc_stream_profile / hvs_mhc_stream_kernel_ms: mean duration per kernel over the launches recorded since
    profiling was enabled, read without waiting inside the profiled region; -1 for kernels that did not run."""
    import ctypes
    import hvs_b200
    lib = hvs_b200._lib.load()
    t = 1 << 14
    x, phi, bias, al, scale = make_inputs(t)
    X = x.cuda()
    P = [p.cuda() for p in (phi, bias, al, scale)]
    dy = torch.randn(t, 4, 512, generator=torch.Generator().manual_seed(3)).to(torch.bfloat16).cuda()
    saved = hvs_b200.ops.new_saved(X)
    assert lib.hvs_mhc_stream_profile(1) == 0
    try:
        for _ in range(3):
            hvs_b200.ops.mhc_stream_fwd(X, *P, saved=saved)
            hvs_b200.ops.mhc_stream_bwd_saved(X, dy, saved, *P)
        buf = (ctypes.c_float * 4)()
        assert lib.hvs_mhc_stream_kernel_ms(buf) == 0
        fwd_ms, bwd_ms, dw_ms, fin_ms = list(buf)
        assert 0.0 < fwd_ms < 5.0 and 0.0 < bwd_ms < 5.0 and 0.0 < fin_ms < 5.0
        assert dw_ms == -1.0                       # the x^T E reduction kernel belongs to the two-kernel backward
        # enabling again resets the ring
        assert lib.hvs_mhc_stream_profile(1) == 0
        assert lib.hvs_mhc_stream_kernel_ms(buf) == 0
        assert list(buf) == [-1.0, -1.0, -1.0, -1.0]
    finally:
        lib.hvs_mhc_stream_profile(0)


def test_more_than_24_iterations_is_inference_only():
    """The forward takes up to 64 Sinkhorn iterations; training (either backward kernel) at most 24.  A training
    call with more must be rejected BEFORE the forward runs, not inside autograd.backward."""
    import hvs_b200
    from hvs_b200._lib import HvsError
    x, phi, bias, al, scale = make_inputs(64, seed=3, alpha=0.3)
    layer = hvs_b200.StreamMHC(sk_iterations=25).cuda()
    with torch.no_grad():
        layer.phi.copy_(phi); layer.bias.copy_(bias); layer.alpha.copy_(al); layer.rms_scale.copy_(scale)
        y = layer(x.cuda())                                       # inference: fine
    ref = mhc_ref.stream_mhc_forward(x, phi, bias, al, scale, sk_iterations=25)
    mag = mhc_ref.mixing_condition_magnitude(x, ref["H_pre"], ref["H_post"], ref["H_res"])
    assert ((y.cpu().float() - ref["y"].float()).abs() <= ULP_BOUND * mhc_ref.bf16_ulp(mag)).all()
    before = hvs_b200._lib.launch_count()
    with pytest.raises(HvsError, match="sk_iterations"):
        layer(x.cuda().requires_grad_(True))
    assert hvs_b200._lib.launch_count() == before                # nothing was launched
    xh = x.pin_memory()
    with pytest.raises(HvsError, match="sk_iterations"):
        hvs_b200.stream_mhc_fwd_bwd_host(xh, xh, layer, torch.empty_like(xh).pin_memory(), torch.empty_like(xh).pin_memory())


def test_bench_configuration_full_size_against_oracle():
    """BASELINE configs[1] as benchmarked: T = 2^20, n = 4, C = 512, bench.py's parameters.  Forward y / coefficients
    and the FUSED backward's dx are compared with the oracle on 4 608 tokens taken from the first, middle and LAST
    tiles (the middle slice straddles the 2^31-byte offset); parameter gradients are compared
    with the oracle on one 2^16-token slice (they are sums over tokens), and the full-size gradients must equal the
    sum of the 16 slices' gradients."""
    import hvs_b200
    t = 1 << 20
    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(t, 4, 512, generator=g, device=dev, dtype=torch.bfloat16)
    dy = torch.randn(t, 4, 512, generator=g, device=dev, dtype=torch.bfloat16)
    gp = torch.Generator().manual_seed(0)
    phi = torch.randn(2048, 24, generator=gp) * 0.02
    bias, al, scale = torch.zeros(24), torch.full((3,), 0.01), torch.ones(2048)
    P = [p.to(dev) for p in (phi, bias, al, scale)]
    saved = hvs_b200.ops.new_saved(x)
    y, _, co = hvs_b200.ops.mhc_stream_fwd(x, *P, want_coeffs=True, saved=saved)
    full = hvs_b200.ops.mhc_stream_bwd_saved(x, dy, saved, *P)
    torch.cuda.synchronize()
    mid = t // 2                                                  # byte offset 2^31 of x / y / dy / dx falls here
    q = 3 * t // 4 + 5
    idx = torch.cat([torch.arange(0, 1536), torch.arange(mid - 512, mid + 512), torch.arange(q - 256, q + 256),
                     torch.arange(t - 1536, t)])
    assert idx.numel() == 4608
    xs, dys = x[idx.to(dev)].cpu(), dy[idx.to(dev)].cpu()
    ys, cos = y[idx.to(dev)].cpu(), co[idx.to(dev)].cpu()
    check_against_oracle(xs, phi, bias, al, scale, ys, torch.einsum("tj,tjc->tc", cos[:, :4], xs.float()).to(torch.bfloat16), cos)
    ref = mhc_ref.stream_mhc_backward(xs, dys, phi, bias, al, scale)
    fwd = mhc_ref.stream_mhc_forward(xs, phi, bias, al, scale)
    m = fwd["H_res"] + fwd["H_post"][:, :, None] * fwd["H_pre"][:, None, :]
    mag = torch.einsum("tij,tic->tjc", m.abs(), dys.float().abs()) + ref["dx"].abs()
    ulps = ((full["dx"][idx.to(dev)].cpu().float() - ref["dx"]).abs() / mhc_ref.bf16_ulp(mag)).max().item()
    assert ulps <= ULP_BOUND, f"full-size fused dx off by {ulps:.2f} bf16 ulp"
    # parameter gradients: additivity over 16 slices, and one slice (the LAST) against the oracle
    sl = t // 16
    acc = None
    last = None
    for k in range(16):
        lo, hi = k * sl, (k + 1) * sl
        part = hvs_b200.ops.mhc_stream_bwd_saved(x[lo:hi], dy[lo:hi], saved[lo:hi], *P)
        acc = {n: part[n].double() for n in PARAM_KEYS} if acc is None else {n: acc[n] + part[n].double() for n in PARAM_KEYS}
        last = part
    for n in PARAM_KEYS:
        rel = ((full[n].double() - acc[n]).norm() / acc[n].norm().clamp_min(1e-30)).item()
        assert rel < 1e-4, f"{n}: full-size gradient != sum of slices ({rel:.2e})"
    lo = t - sl
    ref_sl = mhc_ref.stream_mhc_backward(x[lo:].cpu(), dy[lo:].cpu(), phi, bias, al, scale)
    for n in PARAM_KEYS:
        a, b = last[n].cpu().double(), ref_sl[n].double()
        rel = ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
        assert rel < 3e-4, f"{n} on the last 2^16-token slice: relative error {rel:.2e}"


PARAM_KEYS = ("dphi", "dbias", "dalpha", "dscale")


@pytest.mark.parametrize("alpha,phistd,bstd", [(0.01, 0.02, 0.0), (0.3, 0.03, 0.1), (2.0, 0.2, 0.5)])
def test_adaptive_iterations_stay_within_a_fifth_of_the_coefficient_tolerance(alpha, phistd, bstd):
    """HVS_MHC_ADAPTIVE_ITERS (opt-in): the Sinkhorn loops stop once an iteration changes nothing by more than 2^-20.
    Against the run with all 20 iterations -- at the benchmark's logit scale, trained-like and hot logits (where the test
    never fires and all iterations run): coefficients within 2e-6 relative (north_star's bound against the reference is 1e-5,
    and the oracle comparison below holds it), y within one bf16 ulp (at the condition magnitude) on at most 2 % of the elements, saved statistics
    identical; the fused backward differentiates the iterations actually run: the same bounds against the oracle as the
    full sweep, parameter gradients within 1e-5 of it."""
    import hvs_b200
    dev = "cuda:0"
    t = 6000
    inp = make_inputs(t, seed=77, alpha=alpha, phistd=phistd, bstd=bstd)
    x, phi, bias, al, scale = [v.to(dev) for v in inp]
    dy = torch.randn(t, 4, 512, generator=torch.Generator().manual_seed(3)).to(torch.bfloat16)
    outs = {}
    for adaptive in (False, True):
        saved = hvs_b200.ops.new_saved(x)
        y, u, co = hvs_b200.ops.mhc_stream_fwd(x, phi, bias, al, scale, want_u=True, want_coeffs=True, saved=saved, adaptive=adaptive)
        g = hvs_b200.ops.mhc_stream_bwd_saved(x, dy.to(dev), saved, phi, bias, al, scale, adaptive=adaptive)
        outs[adaptive] = (y, u, co, saved, {k: v.cpu() for k, v in g.items()})
    (yf, uf, cf, sf, gf), (ya, ua, ca, sa, ga) = outs[False], outs[True]
    assert torch.equal(sf, sa) and torch.equal(uf.view(torch.int16), ua.view(torch.int16))      # statistics / H_pre path: no Sinkhorn
    assert ((ca - cf).abs() <= 2e-6 * cf.abs()).all()
    dyy = (ya.float() - yf.float()).abs()
    mag = 3.0 * x.float().abs().sum(1, keepdim=True)                # >= sum_j |M_ij| |x_j| (|M| < 3): where a rounding of y can flip
    assert (dyy <= mhc_ref.bf16_ulp(mag.cpu()).to(dev)).all() and (dyy > 0).float().mean() < 2e-2
    if alpha <= 0.3:
        check_against_oracle(*inp, ya.cpu(), ua.cpu(), ca.cpu())
        check_bwd(inp, dy, ga, f"adaptive alpha={alpha}")
    for k in ("dphi", "dbias", "dalpha", "dscale"):
        rel = ((ga[k].double() - gf[k].double()).norm() / gf[k].double().norm().clamp_min(1e-30)).item()
        assert rel < 1e-5, (k, rel)
    ddx = (ga["dx"].float() - gf["dx"].float()).abs()
    magd = 3.0 * dy.float().abs().sum(1, keepdim=True) + gf["dx"].float().abs()
    assert (ddx <= mhc_ref.bf16_ulp(magd)).all() and (ddx > 0).float().mean() < 2e-2
    if alpha >= 2.0:                                     # hot logits: nothing converges within 20 iterations -> the same numbers
        assert torch.equal(ca, cf) and torch.equal(ya.view(torch.int16), yf.view(torch.int16))
    # the module switch
    layer = hvs_b200.StreamMHC(device=dev, adaptive_sinkhorn=True)
    with torch.no_grad():
        layer.phi.copy_(phi); layer.bias.copy_(bias); layer.alpha.copy_(al); layer.rms_scale.copy_(scale)
        assert torch.equal(layer(x).view(torch.int16), ya.view(torch.int16))
    xg = x.clone().requires_grad_(True)
    layer(xg).backward(dy.to(dev))
    assert torch.equal(xg.grad.cpu().view(torch.int16), ga["dx"].view(torch.int16))
