"""GPU parity tests of the K2 TRAINING kernels: the generalised tcgen05 GEMM (MN-major operands, split-K, the
bias+GELU+dropout "save" epilogue and the GELU'-times-mask data-gradient epilogue), the bias-gradient column sums, the
wide-row LayerNorm backward, the one-node batched coefficient backward and the module's whole training step
(_K2TokenPathFn) against autograd through the oracle (oracle/mhc_ref.py::mhc_module_forward, which restates
/root/reference/src/models/manifold_layers.py:248-267).

Tolerances: the GEMMs multiply exact bf16 operands with fp32 accumulation, so against an fp64 product of the same
operands the error is accumulation order only (1e-5 relative to the row's sum |a||b|) plus one output rounding for bf16
results (2^-8 relative).  The module gradients are compared in norm with the bf16-operand tolerance the library path is
held to (tests/test_gpu_k2.py)."""
import numpy as np
import pytest
import torch

from oracle import mhc_ref

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
BF = torch.bfloat16


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def _close_to_product(got, a, b, out_bf16):
    """got ~ a @ b for exact bf16-valued fp64 a, b."""
    want = a @ b
    mag = a.abs() @ b.abs()
    err = ((got.double().cpu() - want).abs() / mag.clamp_min(1e-30)).max().item()
    assert err < (2.0 ** -8 if out_bf16 else 2e-5), err


@pytest.mark.parametrize("m,n,k", [(128, 64, 64), (300, 256, 72), (1000, 512, 1024), (77, 32, 8), (4099, 1792, 256), (2, 7168, 3584)])
def test_data_gradient_gemm_reads_the_weight_as_it_lies(m, n, k):
    """out[M, N] = a[M, K] @ w[K, N] with w row-major [K, N] (= B MN-major): dX = dY W for nn.Linear's W [out, in]."""
    import hvs_b200
    a = _rand(m, k, seed=1).to(BF)
    w = _rand(k, n, seed=2).to(BF)
    got = hvs_b200.ops.gemm_bf16_ex(a.to(DEV), w.to(DEV), b_mn=True)
    assert got.dtype == BF and got.shape == (m, n)
    _close_to_product(got, a.double(), w.double(), True)
    got32 = hvs_b200.ops.gemm_bf16_ex(a.to(DEV), w.to(DEV), b_mn=True, out_dtype=torch.float32)
    _close_to_product(got32, a.double(), w.double(), False)


@pytest.mark.parametrize("t,fa,fb", [(64, 128, 64), (1000, 32, 32), (5000, 256, 128), (70000, 32, 128), (4099, 1792, 256), (513, 512, 2048),
                                      (7, 64, 64)])
def test_weight_gradient_gemm_contracts_over_tokens(t, fa, fb):
    """out[Fa, Fb] = a^T b over T tokens, both activations as they lie (A and B MN-major); small outputs are split along T
    and the fp32 partials summed in a fixed order: bitwise reproducible."""
    import hvs_b200
    a = _rand(t, fa, seed=3).to(BF)
    b = _rand(t, fb, seed=4).to(BF)
    got = hvs_b200.ops.gemm_wgrad(a.to(DEV), b.to(DEV))
    assert got.dtype == torch.float32 and got.shape == (fa, fb)
    _close_to_product(got, a.double().t(), b.double(), False)
    again = hvs_b200.ops.gemm_wgrad(a.to(DEV), b.to(DEV))
    assert torch.equal(got, again)
    # a strided view (columns of a wider activation) is taken through its row stride
    wide = torch.cat([a, a], 1).to(DEV)
    got_v = hvs_b200.ops.gemm_wgrad(wide[:, fa:], b.to(DEV))
    assert torch.equal(got_v, got)


def test_split_k_choice_and_partials():
    import hvs_b200
    lib = hvs_b200.load_library()
    assert lib.hvs_gemm_choose_split(4096, 2048, 25600) <= 4            # enough tiles already
    s = lib.hvs_gemm_choose_split(256, 128, 409600)
    assert 64 <= s <= 256                                              # two tiles: the token axis feeds the SMs
    assert lib.hvs_gemm_choose_split(32, 32, 64) == 1


@pytest.mark.parametrize("rows,cols", [(1, 32), (511, 64), (512, 256), (513, 256), (10000, 1792), (3000, 7168), (70000, 128)])
def test_bias_gradient_column_sums(rows, cols):
    import hvs_b200
    x = _rand(rows, cols, seed=5).to(BF)
    got = hvs_b200.ops.colsum_bf16(x.to(DEV))
    want = x.double().sum(0)
    assert ((got.cpu().double() - want).abs() <= 1e-5 * x.double().abs().sum(0) + 1e-6).all()
    assert torch.equal(got, hvs_b200.ops.colsum_bf16(x.to(DEV)))


def test_gelu_save_and_dgelu_epilogues_without_dropout():
    """Forward epilogue: z = bf16(a w^T + b), out = bf16(gelu(z)); backward epilogue: dz = da * gelu'(z)."""
    import hvs_b200
    from hvs_b200 import _lib
    m, n, k = 777, 512, 256
    a = _rand(m, k, seed=6).to(BF)
    w = (_rand(n, k, seed=7) * 0.1).to(BF)
    bias = _rand(n, seed=8)
    out, z = hvs_b200.ops.gemm_bf16_ex(a.to(DEV), w.to(DEV), bias=bias.to(DEV), epilogue=_lib.HVS_GEMM_EPI_BIAS_GELU_SAVE)
    z64 = a.double() @ w.double().t() + bias.double()
    assert ((z.cpu().double() - z64).abs() <= 2.0 ** -8 * z64.abs() + 1e-6).all()
    g64 = torch.nn.functional.gelu(z.cpu().double())
    assert ((out.cpu().double() - g64).abs() <= 2.0 ** -8 * g64.abs() + 2e-6).all()
    # backward: da arrives as a GEMM result (da = dz2 @ W), here a plain product
    dz2 = _rand(m, k, seed=9).to(BF)
    w2 = (_rand(k, n, seed=10) * 0.1).to(BF)
    got = hvs_b200.ops.gemm_bf16_ex(dz2.to(DEV), w2.to(DEV), b_mn=True, epilogue=_lib.HVS_GEMM_EPI_DGELU, aux=z)
    zz = z.cpu().double().requires_grad_(True)
    torch.nn.functional.gelu(zz).backward(dz2.double() @ w2.double())
    mag = (dz2.double().abs() @ w2.double().abs()) * 1.2
    assert ((got.cpu().double() - zz.grad).abs() <= 2.0 ** -8 * mag + 1e-6).all()
    # the same launch can also produce the bias gradient: column sums of the bf16 output, summed per 32-row group in the
    # epilogue (transpose-reduce by shuffles) and finished by hvs_colsum_f32 -- no second pass over d z
    got2, db = hvs_b200.ops.gemm_bf16_ex(dz2.to(DEV), w2.to(DEV), b_mn=True, epilogue=_lib.HVS_GEMM_EPI_DGELU, aux=z, want_colsum=True)
    assert torch.equal(got2, got)
    want_db = got.double().sum(0).cpu()
    assert ((db.cpu().double() - want_db).abs() <= 1e-5 * got.double().abs().sum(0).cpu() + 1e-6).all()
    _, db_again = hvs_b200.ops.gemm_bf16_ex(dz2.to(DEV), w2.to(DEV), b_mn=True, epilogue=_lib.HVS_GEMM_EPI_DGELU, aux=z, want_colsum=True)
    assert torch.equal(db, db_again)


def test_dropout_mask_is_reproducible_unbiased_and_shared_by_forward_and_backward():
    import hvs_b200
    from hvs_b200 import _lib
    m, n, k, p = 4096, 256, 64, 0.1
    a = (_rand(m, k, seed=11).abs() + 0.5).to(BF)
    w = (_rand(n, k, seed=12).abs() * 0.1 + 0.05).to(BF)                 # z > 0 everywhere: gelu(z) != 0, so zeros are drops
    bias = torch.zeros(n)
    run = lambda seed: hvs_b200.ops.gemm_bf16_ex(a.to(DEV), w.to(DEV), bias=bias.to(DEV), epilogue=_lib.HVS_GEMM_EPI_BIAS_GELU_SAVE,
                                                 dropout_p=p, dropout_seed=seed)
    out, z = run(1234)
    out_b, _ = run(1234)
    out_c, _ = run(1235)
    assert torch.equal(out, out_b) and not torch.equal(out, out_c)
    keep = (out != 0).float()
    rate = 1.0 - keep.mean().item()
    assert abs(rate - p) < 3e-3, rate                                   # 1 M decisions: sigma = 3e-4
    assert (keep.mean(0) - (1 - p)).abs().max() < 0.03 and (keep.mean(1) - (1 - p)).abs().max() < 0.1   # no dead rows / columns
    nodrop, _ = hvs_b200.ops.gemm_bf16_ex(a.to(DEV), w.to(DEV), bias=bias.to(DEV), epilogue=_lib.HVS_GEMM_EPI_BIAS_GELU_SAVE)
    scaled = nodrop.float() * (65536.0 / (65536.0 - round(p * 65536)))
    kept = out != 0
    assert ((out.float() - scaled)[kept].abs() <= 2.0 ** -7 * scaled[kept].abs()).all()
    # the backward epilogue regenerates the same mask from (seed, row, column)
    da = torch.ones(m, 64).to(BF)
    w2 = torch.ones(64, n).to(BF)
    dz = hvs_b200.ops.gemm_bf16_ex(da.to(DEV), w2.to(DEV), b_mn=True, epilogue=_lib.HVS_GEMM_EPI_DGELU, aux=z, dropout_p=p, dropout_seed=1234)
    assert torch.equal(dz != 0, kept)


@pytest.mark.parametrize("rows,dim", [(130, 1024), (16, 1792), (6400, 1024), (1, 1024), (65, 2048)])
def test_layernorm_backward_wide_rows(rows, dim):
    import hvs_b200
    x = _rand(rows, dim, seed=13) * 1.5 + 0.3
    w = 1.0 + 0.1 * _rand(dim, seed=14)
    dy = _rand(rows, dim, seed=15)
    for xdt, gdt in ((torch.float32, torch.float32), (BF, BF), (BF, torch.float32)):
        xx, gg = x.to(xdt), dy.to(gdt)
        dx, dw, db = hvs_b200.ops.layernorm_bwd(xx.to(DEV), w.to(DEV), gg.to(DEV), 1e-5)
        xl = xx.double().requires_grad_(True)
        wl = w.double().requires_grad_(True)
        bl = torch.zeros(dim, dtype=torch.float64, requires_grad=True)
        torch.nn.functional.layer_norm(xl, (dim,), wl, bl, 1e-5).backward(gg.double())
        tol = 2e-5 if xdt == torch.float32 else 2.0 ** -7
        assert dx.dtype == xdt
        assert ((dx.cpu().double() - xl.grad).abs().max() / xl.grad.abs().max()).item() < tol
        assert ((dw.cpu().double() - wl.grad).norm() / wl.grad.norm()).item() < 1e-5
        assert ((db.cpu().double() - bl.grad).norm() / bl.grad.norm().clamp_min(1e-30)).item() < 1e-5


def _module_pair(d, n, hidden=None, seed=0, raw_std=1.0):
    import hvs_b200
    torch.manual_seed(seed)
    mod = hvs_b200.ManifoldHyperConnection(d, expansion_rate=n, hidden_dim=hidden)
    with torch.no_grad():
        for p in (mod.norm_pre.weight, mod.norm_post.weight):
            p.add_(0.1 * torch.randn_like(p))
        for p in (mod.norm_pre.bias, mod.norm_post.bias, mod.mlp[0].bias, mod.mlp[3].bias):
            p.add_(0.05 * torch.randn_like(p))
        for p in (mod.H_pre_raw, mod.H_post_raw, mod.H_res_raw):
            p.normal_(0, raw_std)
    params = {k: v.detach().clone() for k, v in mod.state_dict().items()}
    return mod.to(DEV), params


GRAD_KEYS = ("H_pre_raw", "H_post_raw", "H_res_raw", "mlp.0.weight", "mlp.0.bias", "mlp.3.weight", "mlp.3.bias",
             "norm_pre.weight", "norm_pre.bias", "norm_post.weight", "norm_post.bias")


def _no_dropout(mod):
    mod.dropout.p = 0.0
    for m in mod.mlp:
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0


@pytest.mark.parametrize("d,n,t", [(64, 4, 200), (32, 4, 4099), (256, 2, 401), (512, 4, 300), (1024, 2, 130), (128, 4, 1000)])
def test_module_training_step_on_the_kernels_vs_oracle_autograd(d, n, t):
    """forward + backward of the module in train mode (dropout off) on _K2TokenPathFn: every gradient against autograd
    through the fp32 oracle, in norm, at the tolerance the library path (cuBLAS under autocast) is held to; and against
    that library path itself, which follows the same bf16 convention (tighter)."""
    import hvs_b200
    mod, p = _module_pair(d, n, seed=3 + d)
    mod.train()
    _no_dropout(mod)
    x = _rand(t, d, seed=8) * 1.5 + 0.2
    dy = _rand(t, d, seed=9)
    xg = x.to(DEV).requires_grad_(True)
    before = hvs_b200._lib.launch_count()
    y = mod(xg)
    y.backward(dy.to(DEV))
    launches = hvs_b200._lib.launch_count() - before
    assert y.dtype == torch.float32
    # everything on this library's kernels: coefficients fwd + bwd, 2 LN fwd, 4 GEMMs, 2 LN bwd (2-3 launches each),
    # 5 data-gradient GEMMs, 5 weight-gradient GEMMs (+ reductions), 2 column sums (2 launches each)
    assert launches >= 2 + 2 + 4 + 4 + 5 + 5 + 4, launches
    got = {k: v.grad.detach().clone() for k, v in mod.named_parameters()}
    gx = xg.grad.detach().clone()
    leaf = {k: (v.clone().requires_grad_(True) if k in GRAD_KEYS else v) for k, v in p.items()}
    xl = x.clone().requires_grad_(True)
    y_ref = mhc_ref.mhc_module_forward(xl, leaf)
    y_ref.backward(dy)

    def rel(a, b):
        return ((a.cpu().double() - b.cpu().double()).norm() / b.cpu().double().norm()).item()
    errs = {k: rel(got[k], leaf[k].grad) for k in GRAD_KEYS}
    errs["x"] = rel(gx, xl.grad)
    errs["y"] = rel(y.detach(), y_ref.detach())
    # the same step on the library path (torch ops under bf16 autocast)
    for q in mod.parameters():
        q.grad = None
    mod.use_training_kernels = False
    xg2 = x.to(DEV).requires_grad_(True)
    mod(xg2).backward(dy.to(DEV))
    mod.use_training_kernels = True
    lib = {k: rel(got[k], v.grad) for k, v in mod.named_parameters()}
    lib["x"] = rel(gx, xg2.grad)
    print(f"[k2 train kernels d={d} n={n} t={t}] vs oracle: " + " ".join(f"{k}:{v:.1e}" for k, v in errs.items()))
    print(f"[k2 train kernels d={d} n={n} t={t}] vs library path: " + " ".join(f"{k}:{v:.1e}" for k, v in lib.items()))
    # measured on B200: 0.5-4 % against the fp32 oracle for the matrices / MLP / x, up to 11 % for norm_post.weight (the
    # LayerNorm input is a bf16 tensor in both bf16 pipelines); 0.3-6 % between the two bf16 pipelines
    assert max(errs.values()) < 0.15, errs
    assert max(v for k, v in errs.items() if "norm_post" not in k and k != "y") < 5e-2, errs
    assert max(lib.values()) < 8e-2, lib


def test_module_training_step_is_deterministic_and_dropout_consistent():
    """Same seed -> bitwise identical step; with dropout on, the backward uses the forward's mask: d loss / d b1 of a loss
    that is linear in the output equals a finite difference of the SAME masked forward."""
    import hvs_b200
    from hvs_b200 import mhc as mhc_mod
    mod, _ = _module_pair(64, 4, seed=21)
    mod.train()
    mod.dropout.p = 0.0                                                 # output dropout is torch's (own RNG stream); the MLP's are the kernel's
    x = (_rand(512, 64, seed=22) * 1.5).to(DEV)
    dy = _rand(512, 64, seed=23).to(DEV)

    def step():
        torch.manual_seed(5)
        mhc_mod._DROPOUT_CALLS[0] = 0
        for q in mod.parameters():
            q.grad = None
        xg = x.clone().requires_grad_(True)
        y = mod(xg)
        y.backward(dy)
        return y.detach().clone(), xg.grad.clone(), {k: v.grad.clone() for k, v in mod.named_parameters()}

    y1, gx1, g1 = step()
    y2, gx2, g2 = step()
    assert torch.equal(y1, y2) and torch.equal(gx1, gx2) and all(torch.equal(g1[k], g2[k]) for k in g1)
    mod.eval()
    with torch.no_grad():
        y_eval = mod(x)
    assert not torch.equal(y_eval, y1)                                  # the masks did something
    assert all(torch.isfinite(v).all() for v in g1.values())


def test_all_layers_coefficient_backward_is_one_launch_and_matches_per_layer_path():
    import hvs_b200
    from hvs_b200.mhc import batched_training_coefficients, clear_training_coefficients
    torch.manual_seed(0)
    dims = (32, 64, 256, 512)
    net = torch.nn.ModuleList([hvs_b200.ManifoldHyperConnection(d, expansion_rate=2) for d in dims]).to(DEV).train()
    for m in net:
        _no_dropout(m)
        with torch.no_grad():
            for q in (m.H_pre_raw, m.H_post_raw, m.H_res_raw):
                q.normal_(0, 1.0)
    xs = [_rand(100, d, seed=d).to(DEV) for d in dims]
    dys = [_rand(100, d, seed=d + 1).to(DEV) for d in dims]

    def run(batched):
        for q in net.parameters():
            q.grad = None
        if batched:
            assert batched_training_coefficients(net) == len(dims)
        try:
            loss = sum((m(x) * dy).sum() for m, x, dy in zip(net, xs, dys))
        finally:
            clear_training_coefficients(net)
        loss.backward()
        return {k: v.grad.clone() for k, v in net.named_parameters()}

    per_layer = run(False)
    before = hvs_b200._lib.launch_count()
    batched = run(True)
    n_batched = hvs_b200._lib.launch_count() - before
    before = hvs_b200._lib.launch_count()
    run(False)
    assert (hvs_b200._lib.launch_count() - before) - n_batched == len(dims) - 1    # one coefficient backward instead of one per layer
    for k in per_layer:                                                 # same arithmetic; the slab cuts (summation order) differ
        assert ((per_layer[k] - batched[k]).norm() / per_layer[k].norm().clamp_min(1e-30)).item() < 1e-5, k


@pytest.mark.parametrize("rows,dim,dts", [(1000, 256, (torch.float32, torch.float32)), (4099, 32, (BF, BF)), (7, 1792, (torch.float32, BF)),
                                          (200000, 64, (BF, torch.float32))])
def test_signal_ratio_monitor_kernel(rows, dim, dts):
    """mean_rows ||out|| / (mean_rows ||x|| + 1e-8) (manifold_layers.py:295-303) written into one slot of the history buffer."""
    import hvs_b200
    out = (_rand(rows, dim, seed=31) * 1.7).to(dts[0])
    x = (_rand(rows, dim, seed=32) * 0.6 + 0.1).to(dts[1])
    hist = torch.full((5,), -1.0, device=DEV)
    hvs_b200.ops.signal_ratio(out.to(DEV), x.to(DEV), hist[2:3])
    want = out.double().norm(dim=-1).mean() / (x.double().norm(dim=-1).mean() + 1e-8)
    got = hist.cpu()
    assert abs(got[2].item() - want.item()) < 2e-5 * want.item()
    assert (got[[0, 1, 3, 4]] == -1.0).all()
    hist2 = torch.zeros(1, device=DEV)
    hvs_b200.ops.signal_ratio(out.to(DEV), x.to(DEV), hist2)
    assert hist2.item() == got[2].item()                                # fixed-order reduction


@pytest.mark.parametrize("m,n,k", [(1, 32, 64), (129, 64, 32), (4099, 256, 128), (70000, 128, 32)])
def test_bias_gradient_from_the_data_gradient_epilogue(m, n, k):
    import hvs_b200
    from hvs_b200 import _lib
    a = _rand(m, k, seed=41).to(BF)
    w = (_rand(k, n, seed=42) * 0.3).to(BF)
    z = _rand(m, n, seed=43).to(BF)
    out, db = hvs_b200.ops.gemm_bf16_ex(a.to(DEV), w.to(DEV), b_mn=True, epilogue=_lib.HVS_GEMM_EPI_DGELU, aux=z.to(DEV), dropout_p=0.1,
                                        dropout_seed=7, want_colsum=True)
    want = out.double().sum(0)
    assert ((db.double() - want).abs() <= 1e-5 * out.double().abs().sum(0) + 1e-6).all()
    assert torch.equal(hvs_b200.ops.colsum_bf16(out), hvs_b200.ops.colsum_bf16(out)) and (hvs_b200.ops.colsum_bf16(out).double() - want).abs().max() <= 1e-4 * out.double().abs().sum(0).max()
