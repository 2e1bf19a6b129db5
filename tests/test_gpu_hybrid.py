"""Whole-model checks on the GPU: hvs_b200's host model (hybrid_vision.py) with the CUDA modules dropped in, against the
fixture written by the REFERENCE's HybridVisionSystem (oracle/make_golden.py::gen_hybrid, name-seeded parameters).

* fp32 mode (every mHC with use_mixed_precision=False, TF32 off): the model must reproduce the reference's CPU outputs
  -- pins composition + coefficient kernels + decode inside the real model;
* default mode (bf16 tcgen05 token path, the reference's CUDA-autocast convention): agreement at bf16-operand level
  (the fixture's coefficients are trained-like, i.e. well conditioned, see tests/test_gpu_k2.py);
* channels_last (layout fold, row a10) changes nothing; detect() end to end; CUDA-graph replay is bitwise stable."""
import numpy as np
import pytest
import torch

from oracle import detect_ref, reference_repaired

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def model():
    import hvs_b200
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    m = hvs_b200.HybridVisionSystem({"num_classes": 80, "image_size": 640}).eval()
    reference_repaired.fill_by_name(m, 0)                   # pure torch helper: same weights as the fixture's model
    return m.to(DEV)


def _set_mixed(model, mixed):
    import hvs_b200
    for m in model.modules():
        if isinstance(m, hvs_b200.ManifoldHyperConnection):
            m.use_mixed_precision = mixed


def test_fp32_mode_reproduces_reference_outputs(model, golden):
    import hvs_b200
    g = golden("hybrid_vision")
    _set_mixed(model, False)
    try:
        before = hvs_b200._lib.launch_count()
        with torch.no_grad():
            out = model(torch.from_numpy(g["x"]).to(DEV))
        assert hvs_b200._lib.launch_count() - before >= 1 + 3       # one coefficient launch for all 76 layers + 3 decodes
    finally:
        _set_mixed(model, True)
    for s in range(3):
        got, want = out["predictions"][f"scale_{s}"].cpu(), torch.from_numpy(g[f"pred{s}"])
        assert torch.allclose(got, want, rtol=5e-3, atol=5e-3), (s, (got - want).abs().max())
    assert torch.allclose(out["final_features"].cpu(), torch.from_numpy(g["final_features"]), rtol=5e-3, atol=5e-3)
    assert torch.allclose(out["vit_features"].mean((2, 3)).cpu(), torch.from_numpy(g["vit_features_mean"]), rtol=5e-3, atol=5e-3)
    for k in ("fused_small", "fused_medium", "fused_large"):
        assert torch.allclose(out["fused_features"][k].mean((2, 3)).cpu(), torch.from_numpy(g[f"{k}_chanmean"]), rtol=5e-3, atol=5e-3)
    # decode inside the model: boxes of the GPU kernel == oracle decode of the same raw predictions
    for s in range(3):
        d = detect_ref.yolo_decode(out["predictions"][f"scale_{s}"].cpu(), detect_ref.anchors_wh(s))
        assert torch.allclose(out["decoded"][f"scale_{s}"]["boxes"].cpu(), d["boxes"], rtol=1e-5, atol=1e-6)
        assert torch.equal(out["decoded"][f"scale_{s}"]["class_indices"].cpu(), d["class_indices"])


def test_bf16_kernel_path_tracks_reference_and_uses_the_k2_kernels(model, golden):
    import hvs_b200
    g = golden("hybrid_vision")
    x = torch.from_numpy(g["x"]).to(DEV)
    with torch.no_grad():
        model(x)
        before = hvs_b200._lib.launch_count()
        out = model(x)
    launches = hvs_b200._lib.launch_count() - before
    # 76 layers x (LN + 4 GEMMs) (+1 LN for D > 512), 14 RMSNorms, 3 decodes; coefficients cached: no refresh launch
    assert launches == 76 * 5 + 2 + 14 + 3, launches
    for s in range(3):
        got, want = out["predictions"][f"scale_{s}"].cpu(), torch.from_numpy(g[f"pred{s}"])
        rel = ((got - want).norm() / want.norm()).item()
        print(f"[hybrid bf16] scale {s}: relative error of raw predictions vs the reference's fp32 CPU forward {rel:.3e}")
        assert rel < 0.15, (s, rel)
    ff, want = out["final_features"].cpu(), torch.from_numpy(g["final_features"])
    assert ((ff - want).norm() / want.norm()) < 0.15


def test_channels_last_is_a_free_layout_fold(model, golden):
    """Row a10: in channels_last the NCHW -> [B*H*W, C] hop is a view (no copy) and the model computes the same thing.
    Compared in fp32 mode (in the bf16 mode 76 layers of re-rounding amplify cuDNN's algorithm choice to ~8 %)."""
    g = golden("hybrid_vision")
    x = torch.from_numpy(g["x"]).to(DEV)
    xc = x.contiguous(memory_format=torch.channels_last)
    t = xc.permute(0, 2, 3, 1)
    assert t.is_contiguous() and t.reshape(-1, 3).data_ptr() == xc.data_ptr()       # the token view aliases the tensor
    _set_mixed(model, False)
    try:
        _channels_last_body(model, x)
    finally:
        _set_mixed(model, True)


def _channels_last_body(model, x):
    with torch.no_grad():
        a = model(x)["predictions"]
        from hvs_b200.hybrid_vision import to_channels_last
        to_channels_last(model)
        try:
            b = model(x.contiguous(memory_format=torch.channels_last))["predictions"]
        finally:
            for m in model.modules():
                if isinstance(m, torch.nn.Conv2d):
                    m.weight.data = m.weight.data.contiguous()
    for s in range(3):
        rel = ((a[f"scale_{s}"] - b[f"scale_{s}"]).norm() / a[f"scale_{s}"].norm()).item()
        assert rel < 2e-3, rel                                      # cuDNN picks other conv algorithms; same math


def test_detect_and_graph_replay_are_deterministic(model):
    x = torch.randn(1, 3, 256, 256, device=DEV, generator=torch.Generator(device=DEV).manual_seed(1))
    with torch.no_grad():
        dets = model.detect(x, confidence_threshold=0.25, iou_threshold=0.45, max_detections=100)
        assert len(dets) == 1 and dets[0]["boxes"].shape[1] == 4 and dets[0]["boxes"].shape[0] <= 100
        assert dets[0]["labels"].dtype == torch.int64
        # whole forward under a CUDA graph (streaming config 5): replay == eager, and replays are bitwise identical
        model.detection_head.want_scores = False
        static_x = x.clone()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                model(static_x)
        torch.cuda.current_stream().wait_stream(s)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = model(static_x)
        eager = {k: v.clone() for k, v in model(x)["predictions"].items()}
        graph.replay()
        first = {k: v.clone() for k, v in out["predictions"].items()}
        graph.replay()
        for k in first:
            assert torch.equal(first[k], out["predictions"][k])
            assert torch.equal(first[k], eager[k])
        model.detection_head.want_scores = True


def test_stability_metrics_match_reference_values(golden):
    """Row a5: get_stability_metrics / the monitoring the reference does inside its training forward
    (manifold_layers.py:282-341) against the values the reference module produced."""
    import hvs_b200
    g = golden("stability")
    for tag, (d, n) in {"d48n2": (48, 2), "d64n4": (64, 4)}.items():
        mod = hvs_b200.ManifoldHyperConnection(d, expansion_rate=n, dropout_rate=0.0, use_mixed_precision=False)
        keys = {k.split("/p/")[1] for k in g.files if k.startswith(tag + "/p/")}
        sd = {k: torch.from_numpy(g[f"{tag}/p/{k}"]) for k in keys}
        for k in ("signal_ratio_history", "eigenvalues", "gradient_norms", "sinkhorn.convergence_history"):
            sd[k] = torch.zeros_like(sd[k])                         # the module has not run yet
        mod.load_state_dict(sd)
        mod = mod.to(DEV).train()
        with torch.no_grad():
            for i in range(3):
                mod(torch.from_numpy(g[f"{tag}/x{i}"]).to(DEV))
        m = mod.get_stability_metrics()
        for k in ("max_eigenvalue", "min_eigenvalue", "eigenvalue_range", "signal_ratio_mean", "signal_ratio_std",
                  "signal_ratio_min", "signal_ratio_max", "signal_ratio", "row_sum_error", "col_sum_error"):
            want = float(g[f"{tag}/m/{k}"])
            assert abs(m[k] - want) <= 2e-4 * max(abs(want), 1e-2), (tag, k, m[k], want)
        assert abs(m["sk_convergence"]["final_convergence"] - float(g[f"{tag}/m/final_convergence"])) < 2e-6
        assert torch.allclose(mod.eigenvalues.cpu(), torch.from_numpy(g[f"{tag}/p/eigenvalues"]), atol=2e-5)


def test_training_step_on_the_kernels_is_as_close_to_fp32_as_the_library_path():
    """Gradients of one whole-model training step (YOLOLoss, dense synthetic targets, BatchNorm in train mode, dropout off)
    in three executions: (a) bf16 autocast with the 76 mHC layers on this library's training kernels (_K2TokenPathFn + one
    coefficient node for all layers), (b) bf16 autocast with their token path on torch ops / library GEMMs -- the
    reference's own CUDA execution --, (c) everything fp32.
    Measured (tools/dbg_train_parity.py): in TRAIN mode this model is chaotic under bf16 -- (b) differs from (c) by ~100 % in
    the raw predictions and by 1.38 in the gradient norm (uncorrelated vectors give sqrt 2), only the last layers of the head
    agree (pred_conv 0.15) -- so "parity" can only mean: (a) is no further from (c) than (b) is, the losses agree to a few
    per cent, the head's last layers agree, and (a) is bitwise reproducible."""
    import hvs_b200
    from hvs_b200 import harness
    torch.manual_seed(0)
    m = hvs_b200.HybridVisionSystem({"num_classes": 80, "image_size": 320})
    reference_repaired.fill_by_name(m, 0)
    m = m.to(DEV).train()
    hvs_b200.hybrid_vision.to_channels_last(m)
    for mod in m.modules():
        if isinstance(mod, (torch.nn.Dropout, torch.nn.Dropout2d)):
            mod.p = 0.0
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 3, 320, 320, generator=g).to(DEV).contiguous(memory_format=torch.channels_last)
    targets = harness.synthetic_targets(2, 320, 0, DEV)
    bn_state = {k: v.clone() for k, v in m.state_dict().items() if "running_" in k or "num_batches" in k}

    def step(kernels, fp32=False):
        for mod in m.modules():
            if isinstance(mod, hvs_b200.ManifoldHyperConnection):
                mod.use_training_kernels = kernels
                mod.use_mixed_precision = not fp32
                mod.dtype = torch.float32 if fp32 else torch.bfloat16
                mod.signal_ratio_idx = 0
        m.load_state_dict(bn_state, strict=False)
        for p in m.parameters():
            p.grad = None
        before = hvs_b200._lib.launch_count()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=not fp32):
            out = m(x, targets=targets, compute_loss=True)
            loss = out["loss"]["total_loss"] + 0.0 * out["final_features"].float().sum()
        loss.backward()
        launches = hvs_b200._lib.launch_count() - before
        return float(loss.detach()), {k: p.grad.detach().double().clone() for k, p in m.named_parameters() if p.grad is not None}, launches

    loss_k, grad_k, n_k = step(True)
    loss_k2, grad_k2, _ = step(True)
    loss_l, grad_l, n_l = step(False)
    loss_f, grad_f, _ = step(False, fp32=True)
    assert loss_k == loss_k2 and all(torch.equal(grad_k[k], grad_k2[k]) for k in grad_k)
    assert n_k > n_l + 76 * 15                                    # the token path's ~30 launches per layer are this library's now
    assert set(grad_k) == set(grad_l) == set(grad_f)
    assert all(torch.isfinite(v).all() for v in grad_k.values())
    assert abs(loss_k - loss_f) <= 5e-2 * abs(loss_f) and abs(loss_l - loss_f) <= 5e-2 * abs(loss_f), (loss_k, loss_l, loss_f)

    def dist(a, b, keys):
        num = sum(float((a[k] - b[k]).pow(2).sum()) for k in keys)
        return (num / sum(float(b[k].pow(2).sum()) for k in keys)) ** 0.5
    every = list(grad_f)
    tail = [k for k in every if "pred_conv" in k or "mhc_enhance.norm_post" in k]
    d_k, d_l = dist(grad_k, grad_f, every), dist(grad_l, grad_f, every)
    t_k, t_l = dist(grad_k, grad_f, tail), dist(grad_l, grad_f, tail)
    print(f"[hybrid training step] loss kernels {loss_k:.3f} library {loss_l:.3f} fp32 {loss_f:.3f}; all gradients vs fp32: kernels {d_k:.3f} "
          f"library {d_l:.3f}; head tail vs fp32: kernels {t_k:.3f} library {t_l:.3f}")
    assert d_k <= 1.1 * d_l + 0.05, (d_k, d_l)
    assert t_k <= 1.1 * t_l + 0.05 and t_k < 0.4, (t_k, t_l)


def test_batchnorm_folding_with_fused_bias_activation_keeps_the_outputs(model):
    """harness.fold_batchnorm_for_inference (eval-mode BatchNorm folded into the convolutions; in ConvMHCLayer the folded
    bias + SiLU run as hvs_bias_act_bf16, the gate multiply + residual as hvs_gate_residual_bf16): same function, bf16
    rounding differences only."""
    import copy
    import hvs_b200
    from hvs_b200 import harness
    m0 = copy.deepcopy(model).eval()
    hvs_b200.hybrid_vision.to_channels_last(m0)
    m1 = copy.deepcopy(m0)
    assert harness.fold_batchnorm_for_inference(m1) >= 40
    hvs_b200.hybrid_vision.to_channels_last(m1)
    # conv -> BatchNorm -> activation stacks of the FPN and the prediction heads: bias + activation as one pass
    assert sum(isinstance(mod, harness.FoldedBiasAct) for mod in m1.modules()) >= 12
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 3, 320, 320, generator=g).to(torch.bfloat16).to(DEV).contiguous(memory_format=torch.channels_last)
    before = hvs_b200._lib.launch_count()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        a = m0(x)
        mid = hvs_b200._lib.launch_count()
        b = m1(x)
    assert (hvs_b200._lib.launch_count() - mid) > (mid - before)          # the fused glue kernels are this library's
    for s in range(3):
        pa, pb = a["predictions"][f"scale_{s}"].float(), b["predictions"][f"scale_{s}"].float()
        rel = ((pa - pb).norm() / pa.norm()).item()
        assert rel < 0.15, (s, rel)                                       # two bf16 pipelines through 76 layers (cf. 9-11 % vs fp32)


def test_weights_cast_once_for_bf16_inference_give_the_same_outputs(model):
    """harness.cast_weights_for_bf16_inference: the per-call bf16 casts of autocast done once.  Same bf16 operands into the
    same kernels: identical predictions, ~230 fewer launches per forward."""
    import copy
    from torch.profiler import profile, ProfilerActivity
    import hvs_b200
    from hvs_b200 import harness
    m0 = copy.deepcopy(model).eval()
    harness.fold_batchnorm_for_inference(m0)
    hvs_b200.hybrid_vision.to_channels_last(m0)
    m1 = copy.deepcopy(m0)
    assert harness.cast_weights_for_bf16_inference(m1) > 100
    assert all(p.dtype == torch.float32 for mod in m1.modules() if isinstance(mod, hvs_b200.ManifoldHyperConnection) for p in mod.parameters())
    g = torch.Generator().manual_seed(12)
    x = torch.randn(2, 3, 320, 320, generator=g).to(torch.bfloat16).to(DEV).contiguous(memory_format=torch.channels_last)

    def run(m):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            m(x)
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):     # a fresh autocast region: no cached casts
                out = m(x)
            torch.cuda.synchronize()
        return out, sum(e.count for e in prof.key_averages())

    a, launches_a = run(m0)
    b, launches_b = run(m1)
    for s in range(3):
        assert torch.equal(a["predictions"][f"scale_{s}"], b["predictions"][f"scale_{s}"]), s
    assert launches_a - launches_b > 100, (launches_a, launches_b)


def test_fused_attention_path_tracks_the_five_op_form():
    """MultiHeadManifoldAttention under bf16 autocast inference uses F.scaled_dot_product_attention; the five-op form
    (matmul, scale, softmax, matmul: manifold_layers.py:410-424) stays for fp32, training and need_weights.  Same function:
    both are equally far from an fp64 evaluation of the same q, k, v -- also with a key padding mask.  (Compared BEFORE the
    output projection: at random initialisation every token's attention output is close to the mean of v, and the
    projection's LayerNorm amplifies the bf16 rounding of either form to O(1) -- 86 % from fp64 for both.)"""
    import hvs_b200
    from hvs_b200.hybrid_vision import MultiHeadManifoldAttention
    torch.manual_seed(3)
    att = MultiHeadManifoldAttention(256, 8).to(DEV).eval()
    att.out_proj = torch.nn.Identity()
    x = torch.randn(3, 401, 256, device=DEV)
    pad = torch.zeros(3, 401, dtype=torch.bool, device=DEV)
    pad[1, 300:] = True
    rel = lambda u, v: ((u.double() - v.double()).norm() / v.double().norm()).item()
    with torch.no_grad():
        split = lambda t: t.reshape(3, -1, 8, 32).transpose(1, 2).double()
        q, k, v = split(att.q_proj(x)), split(att.k_proj(x)), split(att.v_proj(x))
        for mask in (None, pad):
            sc = (q @ k.transpose(-2, -1)) * att.scaling
            if mask is not None:
                sc = sc.masked_fill(mask[:, None, None, :], float("-inf"))
            exact = (torch.softmax(sc, -1) @ v).transpose(1, 2).reshape(3, 401, 256)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                att.use_sdpa = True
                a, wa = att(x, x, x, key_padding_mask=mask)
                att.use_sdpa = False
                b, _ = att(x, x, x, key_padding_mask=mask)
            assert wa is None and a.shape == b.shape
            assert rel(a, exact) < 1e-2 and rel(b, exact) < 1e-2, (rel(a, exact), rel(b, exact))
            assert rel(a, exact) < 1.5 * rel(b, exact) + 1e-3                 # no further from fp64 than the five-op form
        with torch.autocast("cuda", dtype=torch.bfloat16):
            att.use_sdpa = True
            _, w = att(x, x, x, need_weights=True)                   # the weights are only available from the five-op form
        assert w is not None and w.shape == (3, 8, 401, 401)
