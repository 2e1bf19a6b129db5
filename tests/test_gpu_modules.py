"""The nn.Module drop-ins: reference signatures, state_dict keys, parity with the golden vectors."""
import numpy as np
import pytest
import torch

from oracle import detect_ref, mhc_ref

pytestmark = pytest.mark.gpu


def test_stream_mhc_module_autograd_matches_oracle():
    import hvs_b200
    torch.manual_seed(0)
    layer = hvs_b200.StreamMHC(alpha=0.2, device="cuda")
    with torch.no_grad():
        layer.bias.normal_(0, 0.1)
    x = torch.randn(3, 50, 4, 512, device="cuda").to(torch.bfloat16).requires_grad_(True)
    y = layer(x)
    assert y.shape == x.shape and y.dtype == torch.bfloat16
    dy = torch.randn_like(y)
    y.backward(dy)
    xc = x.detach().cpu().reshape(-1, 4, 512)
    ref = mhc_ref.stream_mhc_backward(xc, dy.cpu().reshape(-1, 4, 512), layer.phi.detach().cpu(), layer.bias.detach().cpu(),
                                      layer.alpha.detach().cpu(), layer.rms_scale.detach().cpu())
    assert ((layer.phi.grad.cpu() - ref["dphi"]).norm() / ref["dphi"].norm()) < 2e-3
    assert ((layer.alpha.grad.cpu() - ref["dalpha"]).norm() / ref["dalpha"].norm()) < 2e-3
    assert ((x.grad.cpu().float().reshape(-1, 4, 512) - ref["dx"]).abs().max() / ref["dx"].abs().max()) < 2e-2
    hp, hq, hr = layer.coefficients(x.detach().reshape(-1, 4, 512))
    assert (hr.sum(-1) - 1).abs().max() < 1e-4 and (hr.sum(-2) - 1).abs().max() < 1e-4
    assert (hp >= 0).all() and (hp <= 1).all() and (hq >= 0).all() and (hq <= 2).all()


def test_stream_mhc_wrapped_fn_forward():
    import hvs_b200
    torch.manual_seed(1)
    lin = torch.nn.Linear(512, 512, device="cuda", dtype=torch.bfloat16)
    layer = hvs_b200.StreamMHC(fn=lin, device="cuda")
    x = torch.randn(200, 4, 512, device="cuda").to(torch.bfloat16)
    with torch.no_grad():
        y = layer(x)
    ref = mhc_ref.stream_mhc_forward(x.cpu(), layer.phi.detach().cpu(), layer.bias.detach().cpu(), layer.alpha.detach().cpu(),
                                     layer.rms_scale.detach().cpu(), fn=lambda u: lin(u.cuda()).cpu())
    mag = mhc_ref.mixing_condition_magnitude(x.cpu(), ref["H_pre"], ref["H_post"], ref["H_res"]) + ref["y"].float().abs()
    assert ((y.cpu().float() - ref["y"].float()).abs() <= 3 * mhc_ref.bf16_ulp(mag)).all()


def test_host_buffer_entry_matches_device_path():
    import hvs_b200
    torch.manual_seed(2)
    layer = hvs_b200.StreamMHC(device="cuda")
    t = 5000
    xh = torch.randn(t, 4, 512).to(torch.bfloat16).pin_memory()
    dyh = torch.randn(t, 4, 512).to(torch.bfloat16).pin_memory()
    yh = torch.empty_like(xh).pin_memory()
    dxh = torch.empty_like(xh).pin_memory()
    g = hvs_b200.stream_mhc_fwd_bwd_host(xh, dyh, layer, yh, dxh, chunk_tokens=2048)
    p = (layer.phi.detach(), layer.bias.detach(), layer.alpha.detach(), layer.rms_scale.detach())
    saved = hvs_b200.ops.new_saved(xh.cuda())
    y, _, _ = hvs_b200.ops.mhc_stream_fwd(xh.cuda(), *p, saved=saved)
    gd = hvs_b200.ops.mhc_stream_bwd_saved(xh.cuda(), dyh.cuda(), saved, *p)     # the training path the pipeline uses
    assert torch.equal(yh.view(torch.int16), y.cpu().view(torch.int16))
    assert torch.equal(dxh.view(torch.int16), gd["dx"].cpu().view(torch.int16))
    assert torch.allclose(g["dphi"], gd["dphi"].cpu(), rtol=1e-3, atol=1e-5 * gd["dphi"].abs().max().item())


def test_sinkhorn_module_reference_properties():
    """reference src/tests/test_models.py:33-56, :85-100 on the CUDA module."""
    import hvs_b200
    sk = hvs_b200.SinkhornKnoppProjection(num_iterations=20).cuda()
    m = torch.randn(4, 8, 8, device="cuda")
    p = sk(m)
    assert torch.all(p >= 0)
    assert torch.allclose(p.sum(2), torch.ones(4, 8, device="cuda"), rtol=1e-4)
    assert torch.allclose(p.sum(1), torch.ones(4, 8, device="cuda"), rtol=1e-4)
    assert torch.equal(sk(m), p)
    assert set(sk.get_convergence_metrics()) == {"mean_convergence", "max_convergence", "final_convergence"}
    # gradient flows (test_models.py:58-83)
    w = torch.nn.Parameter(torch.randn(3, 3, device="cuda"))
    loss = torch.nn.functional.mse_loss(hvs_b200.SinkhornKnoppProjection(5).cuda()(w), torch.eye(3, device="cuda"))
    loss.backward()
    assert w.grad is not None and w.grad.abs().sum() > 0


def test_manifold_hyper_connection_dropin(golden):
    import hvs_b200
    g = golden("mhc_module")
    for tag, (d, n) in {"d64n4": (64, 4), "d32n2": (32, 2)}.items():
        mod = hvs_b200.ManifoldHyperConnection(d, expansion_rate=n, use_mixed_precision=False).cuda().eval()
        ref_keys = {k.split("/p/")[1] for k in g.files if k.startswith(tag + "/p/")}
        assert set(mod.state_dict().keys()) == ref_keys              # reference checkpoints load
        mod.load_state_dict({k: torch.from_numpy(g[f"{tag}/p/{k}"]) for k in ref_keys})
        x = torch.from_numpy(g[f"{tag}/x"]).cuda()
        with torch.no_grad():
            y = mod(x)
            hp, hq, hr = mod.constrained_matrices()
        assert y.shape == x.shape
        assert torch.allclose(y.cpu(), torch.from_numpy(g[f"{tag}/y"]), rtol=2e-4, atol=2e-4)   # fp32 GEMMs (TF32 off)
        assert ((hr.cpu() - torch.from_numpy(g[f"{tag}/H_res"])).abs() / torch.from_numpy(g[f"{tag}/H_res"])).max() < 1e-5
        # test_models.py:145-159 properties
        assert (hp >= 0).all() and (hp <= 1).all() and (hq >= 0).all() and (hq <= 2).all()
        assert torch.allclose(hr.sum(0), torch.ones(d, device="cuda"), rtol=1e-3)
        m = mod.get_stability_metrics()
        for key in ("max_eigenvalue", "min_eigenvalue", "eigenvalue_range", "sk_convergence"):
            assert key in m
        assert m["max_eigenvalue"] <= 1.0 + 1e-4


def test_manifold_hyper_connection_bf16_and_grad():
    """test_models.py:163-204: finite gradients with 0 < |grad x| < 100; bf16 autocast path runs."""
    import hvs_b200
    torch.manual_seed(3)
    mod = hvs_b200.ManifoldHyperConnection(64, expansion_rate=4, dropout_rate=0.0).cuda().train()
    x = torch.randn(16, 64, device="cuda", requires_grad=True)
    out = mod(x)
    (out * torch.randn_like(out)).sum().backward()      # (a plain sum of a LayerNorm output has zero gradient)
    gn = x.grad.norm().item()
    assert np.isfinite(gn) and 0 < gn < 100
    assert mod.H_res_raw.grad is not None and torch.isfinite(mod.H_res_raw.grad).all()
    mod.eval()
    with torch.no_grad():
        y = mod(torch.randn(2, 5, 7, 64, device="cuda"))
    assert y.shape == (2, 5, 7, 64) and torch.isfinite(y).all()
    # cache invalidation when a parameter changes
    h1 = mod.constrained_matrices()[2].clone()
    with torch.no_grad():
        mod.H_res_raw.add_(0.5 * torch.randn_like(mod.H_res_raw))
    assert not torch.equal(h1, mod.constrained_matrices()[2])


def test_detection_head_postprocess_and_nms(golden):
    import hvs_b200
    g = golden("nms")
    head = hvs_b200.YOLODetectionHead([32, 64, 128], num_classes=80, use_mhc=False).cuda().eval()
    keys = set(head.state_dict().keys())
    assert "anchor_generator.anchors" in keys and "pred_heads.0.pred_conv.weight" in keys and "pred_heads.2.conv_layers.4.weight" in keys
    # known answer + agnostic golden through the module methods
    keep = head.non_max_suppression(torch.from_numpy(g["ag300/boxes"]).cuda(), torch.from_numpy(g["ag300/scores"]).cuda(),
                                    iou_threshold=float(g["ag300/thr"]), max_detections=int(g["ag300/cap"]))
    assert keep.cpu().tolist() == g["ag300/keep"].tolist()
    f = hvs_b200.NMSFilter(hvs_b200.PostprocessingConfig(nms_iou_threshold=0.5))
    keep = f.apply(torch.from_numpy(g["ka/boxes"]).cuda(), torch.from_numpy(g["ka/scores"]).cuda(), torch.from_numpy(g["ka/classes"]).cuda())
    assert keep.cpu().tolist() == [0, 2]
    # full path: features -> head -> decode (strided view) -> post_process, against the oracle on the same predictions
    torch.manual_seed(4)
    feats = {"scale_small": torch.randn(2, 32, 16, 16, device="cuda"), "scale_medium": torch.randn(2, 64, 8, 8, device="cuda"),
             "scale_large": torch.randn(2, 128, 4, 4, device="cuda")}
    with torch.no_grad():
        for h in head.pred_heads:                      # spread the scores so some candidates pass the threshold
            h.pred_conv.weight.normal_(0, 0.5)
        out = head(feats)
    dets = head.post_process(out["decoded"], confidence_threshold=0.3, iou_threshold=0.45, max_detections=20)
    decoded_cpu = []
    for s in range(3):
        pred = out["predictions"][f"scale_{s}"]
        assert not pred.is_contiguous()                # the decode kernel read the permuted view in place
        d = detect_ref.yolo_decode(pred.cpu(), detect_ref.anchors_wh(s))
        assert torch.allclose(out["decoded"][f"scale_{s}"]["boxes"].cpu(), d["boxes"], rtol=1e-5, atol=1e-6)
        # NMS is compared on the SAME decoded tensors: take the GPU decode as the common input
        decoded_cpu.append({k: out["decoded"][f"scale_{s}"][k].cpu() for k in ("boxes", "class_scores", "class_indices")})
    want = detect_ref.post_process(decoded_cpu, 0.3, 0.45, 20)
    for b in range(2):
        assert np.array_equal(dets[b]["boxes"].cpu().numpy(), want[b]["boxes"])
        assert np.array_equal(dets[b]["labels"].cpu().numpy(), want[b]["labels"])
