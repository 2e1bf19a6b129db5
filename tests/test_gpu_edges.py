"""Edges of the path: the fused dual-threshold gradient clipping against torch's
clip_grad_norm_ applied the way the reference trainer applies it (mhc_trainer.py:342-383), and the frame preprocessing
kernel against cv2 + torch (preprocessing.py:252-273)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _named_grads(seed, scale):
    g = torch.Generator().manual_seed(seed)
    shapes = {"backbone.stem.0.mhc.H_pre_raw": (32, 128), "backbone.stem.0.mhc.mlp.0.weight": (256, 128), "backbone.stem.0.conv.weight": (32, 3, 3, 3),
              "final_fusion.H_res_raw": (300, 300), "vit.blocks.0.mlp.0.bias": (1024,), "head.pred_conv.weight": (255, 256, 1, 1),
              "big.weight": (70000, 3), "nograd.weight": (4, 4)}
    params = []
    for n, s in shapes.items():
        p = torch.nn.Parameter(torch.zeros(s))
        if not n.startswith("nograd"):
            p.grad = torch.randn(s, generator=g) * scale
        params.append((n, p))
    return params


@pytest.mark.parametrize("scale", [1e-4, 0.01, 1.0])
def test_grad_clip_dual_matches_reference_rule(scale):
    import hvs_b200
    cpu = _named_grads(3, scale)
    gpu = [(n, torch.nn.Parameter(p.detach().to(DEV))) for n, p in cpu]
    for (n, p), (_, q) in zip(cpu, gpu):
        if p.grad is not None:
            q.grad = p.grad.to(DEV)
    mhc = [p for n, p in cpu if p.grad is not None and hvs_b200.ops.is_mhc_parameter(n)]
    oth = [p for n, p in cpu if p.grad is not None and not hvs_b200.ops.is_mhc_parameter(n)]
    assert len(mhc) == 3 and len(oth) == 4
    n_mhc = torch.nn.utils.clip_grad_norm_(mhc, max_norm=0.5, norm_type=2)
    n_oth = torch.nn.utils.clip_grad_norm_(oth, max_norm=1.0, norm_type=2)
    before = hvs_b200._lib.launch_count()
    res = hvs_b200.ops.clip_grad_dual(gpu, max_grad_norm=1.0, mhc_max_norm=0.5)
    assert hvs_b200._lib.launch_count() - before == 3
    r = res.cpu()
    assert abs(r[0] - n_mhc) <= 1e-5 * n_mhc and abs(r[1] - n_oth) <= 1e-5 * n_oth
    for (n, p), (_, q) in zip(cpu, gpu):
        if p.grad is not None:
            assert torch.allclose(q.grad.cpu(), p.grad, rtol=2e-6, atol=1e-12), n
    # deterministic
    gpu2 = [(n, torch.nn.Parameter(p.detach().to(DEV))) for n, p in _named_grads(3, scale)]
    for (n, p), (_, q) in zip(_named_grads(3, scale), gpu2):
        if p.grad is not None:
            q.grad = p.grad.to(DEV)
    assert torch.equal(hvs_b200.ops.clip_grad_dual(gpu2, 1.0, 0.5), res)


@pytest.mark.parametrize("shape,out", [((480, 640, 3), (640, 640)), ((1080, 1920, 3), (640, 640)), ((100, 77, 3), (64, 96)), ((640, 640, 3), (640, 640)),
                                       ((50, 60), (32, 32))])
def test_preprocess_frame_matches_cv2_pipeline(shape, out):
    import cv2
    import hvs_b200
    rng = np.random.default_rng(sum(shape))
    frame = rng.integers(0, 256, size=shape, dtype=np.uint8)
    h, w = out
    # the reference pipeline: colour conversion (preprocessing.py:199-213), cv2.resize INTER_LINEAR (:255-259), /255, normalise (:262-271)
    rgb = cv2.cvtColor(frame, cv2.COLOR_BGR2RGB) if frame.ndim == 3 else cv2.cvtColor(frame, cv2.COLOR_GRAY2RGB)
    resized = cv2.resize(rgb, (w, h), interpolation=cv2.INTER_LINEAR)
    want = torch.from_numpy(resized).permute(2, 0, 1).float() / 255.0
    mean, std = torch.tensor(hvs_b200.ops.IMAGENET_MEAN).view(3, 1, 1), torch.tensor(hvs_b200.ops.IMAGENET_STD).view(3, 1, 1)
    want = (want - mean) / std
    got = hvs_b200.ops.preprocess_frame(torch.from_numpy(frame).to(DEV), h, w).cpu()
    # cv2 interpolates uint8 in fixed point and rounds to uint8; the kernel interpolates in fp32: at most one grey level apart
    assert ((got - want).abs() * std).max() <= 1.01 / 255.0
    if shape[:2] == out:
        assert ((got - want).abs() * std).max() < 1e-6             # no resampling: identical
    g16 = hvs_b200.ops.preprocess_frame(torch.from_numpy(frame).to(DEV), h, w, out_dtype=torch.bfloat16)
    assert torch.equal(g16.cpu(), got.to(torch.bfloat16))


@pytest.mark.parametrize("b,c,h,w,with_res", [(2, 64, 7, 5, True), (3, 32, 16, 16, False), (1, 1024, 3, 3, True), (4, 128, 40, 40, True)])
def test_gate_residual_matches_torch(b, c, h, w, with_res):
    """y * gate (+ x) of ConvMHCLayer (vision_backbone.py:125-133) in one pass: fp32 arithmetic, one rounding -- within one
    bf16 ulp of torch's two rounded passes."""
    import hvs_b200
    g = torch.Generator().manual_seed(b * c + h)
    cl = torch.channels_last
    y = torch.randn(b, c, h, w, generator=g).to(torch.bfloat16).to(DEV).contiguous(memory_format=cl)
    x = torch.randn(b, c, h, w, generator=g).to(torch.bfloat16).to(DEV).contiguous(memory_format=cl)
    gate = torch.sigmoid(torch.randn(b, c, 1, 1, generator=g)).to(torch.bfloat16).to(DEV)
    got = hvs_b200.ops.gate_residual(y, gate, x if with_res else None)
    exact = y.double() * gate.double() + (x.double() if with_res else 0.0)
    assert got.shape == y.shape and got.is_contiguous(memory_format=cl) and got.dtype == torch.bfloat16
    ulp = 2.0 ** (torch.floor(torch.log2(exact.abs().clamp_min(1e-30))) - 7)
    assert ((got.double() - exact).abs() <= 0.51 * ulp + 1e-30).all()           # fp32 fma, one rounding to bf16
    want = y * gate + x if with_res else y * gate                              # torch: the product is rounded before the add
    mag = (y.double() * gate.double()).abs() + (x.double().abs() if with_res else 0.0)
    assert ((got.double() - want.double()).abs() <= 2.0 ** -7 * mag + 1e-30).all()


@pytest.mark.parametrize("act", ["silu", "relu", "none", "leaky_relu_0.1"])
def test_bias_act_matches_torch(act):
    import hvs_b200
    g = torch.Generator().manual_seed(3)
    y = torch.randn(3, 64, 9, 11, generator=g).to(torch.bfloat16).to(DEV).contiguous(memory_format=torch.channels_last)
    bias = torch.randn(64, generator=g).to(DEV)
    z = y.double() + bias.double().view(1, -1, 1, 1)
    want = {"silu": torch.nn.functional.silu, "relu": torch.relu, "none": lambda t: t,
            "leaky_relu_0.1": lambda t: torch.nn.functional.leaky_relu(t, 0.1)}[act](z)
    got = hvs_b200.ops.bias_act(y.clone(memory_format=torch.channels_last), bias, act)
    assert got.is_contiguous(memory_format=torch.channels_last)
    assert ((got.double() - want).abs() <= 2.0 ** -8 * want.abs() + 1e-6).all()


@pytest.mark.parametrize("b,c,hw,act", [(1, 32, (320, 320), "silu"), (3, 64, (40, 28), "silu"), (2, 256, (20, 20), "relu"),
                                        (64, 512, (20, 20), "silu"), (1, 96, (7, 5), "silu"), (2, 2048, (3, 3), "none"), (5, 160, (9, 11), "relu")])
def test_se_gate_matches_the_torch_modules_under_autocast(b, c, hw, act):
    """hvs_se_gate_bf16 against ConvMHCLayer.channel_attention as torch runs it under bf16 autocast
    (AdaptiveAvgPool2d(1) -> 1x1 conv -> act -> 1x1 conv -> sigmoid, vision_backbone.py:77-83): same rounding points, so the
    gates agree to a bf16 ulp or two; repeated calls (workspace reuse) are bitwise identical."""
    import hvs_b200
    from hvs_b200 import ops
    g = torch.Generator().manual_seed(c + b)
    dev = "cuda"
    y = (torch.randn(b, c, *hw, generator=g) * 1.5 + 0.3).to(torch.bfloat16).to(dev).contiguous(memory_format=torch.channels_last)
    actm = {"silu": torch.nn.SiLU(), "relu": torch.nn.ReLU(), "none": torch.nn.Identity()}[act]
    ca = torch.nn.Sequential(torch.nn.AdaptiveAvgPool2d(1), torch.nn.Conv2d(c, c // 4, 1), actm, torch.nn.Conv2d(c // 4, c, 1),
                             torch.nn.Sigmoid()).to(dev)
    with torch.no_grad():
        for m in (ca[1], ca[3]):
            m.weight.mul_(3.0); m.bias.normal_(0, 0.5, generator=None)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            want = ca(y)
        got, ws = ops.se_gate(y, ca[1].weight, ca[1].bias, ca[3].weight, ca[3].bias, act)
        again, ws2 = ops.se_gate(y, ca[1].weight, ca[1].bias, ca[3].weight, ca[3].bias, act, ws)
        # fp64 evaluation of the same bf16 operands, for scale
        p = y.double().mean((2, 3))
        w1, b1 = ca[1].weight.to(torch.bfloat16).double().flatten(1), ca[1].bias.to(torch.bfloat16).double()
        w2, b2 = ca[3].weight.to(torch.bfloat16).double().flatten(1), ca[3].bias.to(torch.bfloat16).double()
        hdn = p @ w1.t() + b1
        hdn = {"silu": torch.nn.functional.silu, "relu": torch.relu, "none": lambda t: t}[act](hdn)
        exact = torch.sigmoid(hdn @ w2.t() + b2)
    assert ws2 is ws and torch.equal(got, again)
    assert got.shape == (b, c, 1, 1) and got.dtype == torch.bfloat16 and want.dtype == torch.bfloat16
    e_got = (got.double().flatten(1) - exact).abs().max().item()
    e_torch = (want.double().flatten(1) - exact).abs().max().item()
    assert e_got <= max(2.0 * e_torch, 8e-3), (e_got, e_torch)            # no further from fp64 than the torch path (bf16 ulp at 1: 7.8e-3)
    assert (got.float() - want.float()).abs().max().item() <= 1.6e-2      # two bf16 ulp at a gate near 1
