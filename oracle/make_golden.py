"""Pin the oracle against the reference and write tests/golden/*.npz.

TEST INFRASTRUCTURE.  Run in the build container only (needs /root/reference):

    python -m oracle.make_golden

For every fixture the outputs are produced by the REFERENCE's own modules
(oracle/reference_repaired.py) and the restatement (oracle/mhc_ref.py,
oracle/detect_ref.py) is asserted against them before anything is written.
The report is written to tests/golden/PINNING.txt.
"""
from __future__ import annotations

import io
import os
import sys

import numpy as np
import torch

from . import detect_ref, mhc_ref, reference_repaired

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
_report = io.StringIO()


def say(msg: str):
    print(msg)
    _report.write(msg + "\n")


def bf16_bits(t: torch.Tensor) -> np.ndarray:
    return t.to(torch.bfloat16).view(torch.int16).numpy().astype(np.uint16)


def maxrel(a: torch.Tensor, b: torch.Tensor) -> float:
    return float(((a - b).abs() / b.abs().clamp_min(1e-30)).max())


def gen_sinkhorn(ref):
    g = torch.Generator().manual_seed(100)
    cases = {
        "a": torch.randn(4, 8, 8, generator=g),                 # test_models.py:38-41
        "b": torch.randn(64, 4, 4, generator=g) * 0.3,          # per-token 4x4 blocks
        "c": torch.randn(16, 16, generator=g) * 0.05,           # 2-D (R1)
        "d": torch.randn(2, 5, 5, generator=g),                 # test_models.py:89
    }
    out = {}
    for k, m in cases.items():
        iters = 10 if k == "d" else 20
        sk = ref.SinkhornKnoppProjection(num_iterations=iters)
        with torch.no_grad():
            want = sk(m)
            hist = sk.convergence_history.clone()
        got, ghist = mhc_ref.sinkhorn_knopp(m, iters, return_history=True)
        assert torch.equal(got, want), f"sinkhorn case {k}: restatement != reference"
        assert torch.allclose(ghist, hist, atol=1e-7), f"sinkhorn history {k}"
        rs = (want.sum(-1) - 1).abs().max().item()
        cs = (want.sum(-2) - 1).abs().max().item()
        say(f"sinkhorn[{k}] shape={tuple(m.shape)} iters={iters}: restatement bit-equal to reference; "
            f"row err {rs:.2e} col err {cs:.2e}")
        out[f"{k}_in"] = m.numpy()
        out[f"{k}_out"] = want.numpy()
        out[f"{k}_hist"] = hist.numpy()
        out[f"{k}_iters"] = np.int64(iters)
    np.savez_compressed(os.path.join(GOLDEN, "sinkhorn.npz"), **out)


def gen_rmsnorm(ref):
    g = torch.Generator().manual_seed(101)
    x = torch.randn(7, 2048, generator=g)
    mod = ref.RMSNorm(2048)
    with torch.no_grad():
        mod.scale.copy_(1.0 + 0.1 * torch.randn(2048, generator=g))
        want = mod(x)
    got = mhc_ref.rms_norm(x, mod.scale.detach())
    assert torch.equal(got, want)
    say("rmsnorm dim=2048: restatement bit-equal to reference")
    np.savez_compressed(os.path.join(GOLDEN, "rmsnorm.npz"), x=x.numpy(),
                        scale=mod.scale.detach().numpy(), out=want.numpy())


def gen_module(ref):
    """K2: reference-literal ManifoldHyperConnection, eval mode (CPU: fp32)."""
    out = {}
    for tag, (d, n, shape) in {"d64n4": (64, 4, (37, 64)), "d32n2": (32, 2, (2, 5, 3, 32))}.items():
        torch.manual_seed(102)
        mod = ref.ManifoldHyperConnection(d, expansion_rate=n).eval()
        x = torch.randn(*shape)
        with torch.no_grad():
            want = mod(x)
            hp, hq, hr = mod.constrained_matrices()
        params = {k: v.detach() for k, v in mod.state_dict().items()}
        got = mhc_ref.mhc_module_forward(x, params)
        err = (got - want).abs().max().item()
        assert err < 1e-5, err
        g_hp, g_hq, g_hr = mhc_ref.constrained_matrices(params["H_pre_raw"], params["H_post_raw"], params["H_res_raw"])
        assert torch.equal(g_hr, hr) and torch.equal(g_hp, hp) and torch.equal(g_hq, hq)
        say(f"mhc_module[{tag}] x{tuple(shape)}: restatement max|diff| vs reference {err:.2e}; "
            f"constrained_matrices bit-equal; H_res row err {(hr.sum(1)-1).abs().max():.2e}")
        for k, v in params.items():
            out[f"{tag}/p/{k}"] = v.numpy()
        out[f"{tag}/x"] = x.numpy()
        out[f"{tag}/y"] = want.numpy()
        out[f"{tag}/H_pre"] = hp.numpy(); out[f"{tag}/H_post"] = hq.numpy(); out[f"{tag}/H_res"] = hr.numpy()
    np.savez_compressed(os.path.join(GOLDEN, "mhc_module.npz"), **out)


def gen_stream(ref):
    """K1: the stream layer composed ONLY of reference primitives (RMSNorm
    module, torch.sigmoid gates as in :213/:216, batched SinkhornKnoppProjection)."""
    out = {}
    for tag, (t, n, c, alpha, phistd, bstd) in {
        "n4c512": (96, 4, 512, 0.01, 0.02, 0.0),
        "n4c512_hot": (48, 4, 512, 0.3, 0.05, 0.2),     # larger logits: stresses Sinkhorn
    }.items():
        g = torch.Generator().manual_seed(103)
        x = torch.randn(t, n, c, generator=g).to(torch.bfloat16)
        k = n * n + 2 * n
        phi = torch.randn(n * c, k, generator=g) * phistd
        bias = torch.randn(k, generator=g) * bstd
        al = torch.full((3,), alpha)
        scale = 1.0 + 0.05 * torch.randn(n * c, generator=g)
        # --- reference primitives ---
        norm = ref.RMSNorm(n * c)
        sk = ref.SinkhornKnoppProjection(20)
        with torch.no_grad():
            norm.scale.copy_(scale)
            xf = x.float().reshape(t, n * c)
            # operand convention: projection weights bf16 (autocast, :248); the
            # RMSNorm gain is folded into them
            w = (scale[:, None] * phi).to(torch.bfloat16).float()
            xn = norm(xf) / scale            # normalised row without the gain
            raw = xn @ w
            a = torch.cat([al[0].expand(n), al[1].expand(n), al[2].expand(n * n)])
            logits = raw * a + bias
            h_pre = torch.sigmoid(logits[:, :n])
            h_post = 2 * torch.sigmoid(logits[:, n:2 * n])
            h_res = sk(logits[:, 2 * n:].reshape(t, n, n))
            xs = x.float()
            u = torch.einsum("tj,tjc->tc", h_pre, xs)
            y = torch.einsum("tij,tjc->tic", h_res, xs) + h_post[:, :, None] * u[:, None, :]
        got = mhc_ref.stream_mhc_forward(x, phi, bias, al, scale)
        e_res = maxrel(got["H_res"], h_res)
        e_pre = maxrel(got["H_pre"], h_pre)
        e_post = maxrel(got["H_post"], h_post)
        assert max(e_res, e_pre, e_post) < 5e-6, (e_res, e_pre, e_post)
        ey = (got["y"].float() - y).abs().max().item()
        say(f"stream_mhc[{tag}] T={t}: restatement vs reference-primitive composition: "
            f"rel err H_pre {e_pre:.1e} H_post {e_post:.1e} H_res {e_res:.1e}; "
            f"row err {(h_res.sum(-1)-1).abs().max():.1e} col err {(h_res.sum(-2)-1).abs().max():.1e}; "
            f"max|y_bf16 - y_fp32| {ey:.2e}")
        bw = mhc_ref.stream_mhc_backward(x, torch.randn(t, n, c, generator=g).to(torch.bfloat16),
                                         phi, bias, al, scale)
        out[f"{tag}/x_bits"] = bf16_bits(x)
        out[f"{tag}/phi"] = phi.numpy(); out[f"{tag}/bias"] = bias.numpy()
        out[f"{tag}/alpha"] = al.numpy(); out[f"{tag}/scale"] = scale.numpy()
        out[f"{tag}/H_pre"] = h_pre.numpy(); out[f"{tag}/H_post"] = h_post.numpy()
        out[f"{tag}/H_res"] = h_res.numpy(); out[f"{tag}/y"] = y.numpy()
        out[f"{tag}/u"] = u.numpy()
        del bw
    np.savez_compressed(os.path.join(GOLDEN, "stream_mhc.npz"), **out)


def gen_decode(ref):
    g = torch.Generator().manual_seed(104)
    b, a, h, w, nc = 2, 3, 8, 8, 80
    pred = torch.randn(b, a, h, w, 5 + nc, generator=g) * 1.5
    head = ref.YOLODetectionHead([32, 64, 128], num_classes=nc, use_mhc=False)
    out = {}
    for s in range(3):
        anchors = head.anchor_generator(s)                 # [A,1,1,4] (R2/R5)
        with torch.no_grad():
            ship = head.decoder(pred, anchors, (h, w))     # shipped decoder, D4 unpatched
        mine = detect_ref.yolo_decode(pred, detect_ref.anchors_wh(s))
        assert torch.equal(mine["scores"], ship["scores"])
        assert torch.equal(mine["class_scores"], ship["class_scores"])
        assert torch.equal(mine["class_indices"], ship["class_indices"])
        assert torch.equal(mine["objectness"], ship["objectness"])
        sb = ship["boxes"]                                   # [B,A,H,W,W,4]
        assert sb.dim() == 6
        diag = sb.diagonal(dim1=3, dim2=4).permute(0, 1, 2, 4, 3)   # [B,A,H,W,4] at w1==w2
        assert torch.equal(mine["boxes"][..., 0], diag[..., 0])     # x1
        assert torch.equal(mine["boxes"][..., 2], diag[..., 2])     # x2
        # box height: y2-y1 of the shipped output equals anchor_h*exp(th) on any slice
        assert torch.allclose(mine["boxes"][..., 3] - mine["boxes"][..., 1],
                              diag[..., 3] - diag[..., 1], atol=1e-6)
        say(f"decode[scale {s}]: scores/class max/argmax/objectness/x1/x2 bit-equal to the shipped decoder; "
            f"y1,y2 follow the documented formula (yolo_head.py:257-258), shipped y is D4-garbled")
        out[f"s{s}/boxes"] = mine["boxes"].numpy()
        out[f"s{s}/class_scores"] = mine["class_scores"].numpy()
        out[f"s{s}/class_indices"] = mine["class_indices"].numpy()
        out[f"s{s}/objectness"] = mine["objectness"].numpy()
        out[f"s{s}/anchor_wh"] = detect_ref.anchors_wh(s).numpy()
    out["pred"] = pred.numpy()
    np.savez_compressed(os.path.join(GOLDEN, "decode.npz"), **out)
    return head


def random_boxes(g, n, spread=1.0):
    cx = torch.rand(n, generator=g) * spread
    cy = torch.rand(n, generator=g) * spread
    w = 0.05 + 0.3 * torch.rand(n, generator=g)
    h = 0.05 + 0.3 * torch.rand(n, generator=g)
    return torch.stack([cx, cy, w, h], -1)


def gen_nms(ref, head):
    out = {}
    # (1) the reference's own known-answer vector, test_inference.py:361-379;
    # consistent only with NMSFilter.apply's cxcywh reading at thr 0.5 -> keep [0, 2]
    kb = torch.tensor([[0.1, 0.1, 0.3, 0.3], [0.15, 0.15, 0.35, 0.35], [0.6, 0.6, 0.8, 0.8]])
    ks = torch.tensor([0.9, 0.8, 0.7])
    kc = torch.zeros(3, dtype=torch.long)
    cfg = ref.PostprocessingConfig(nms_iou_threshold=0.5)
    keep = ref.NMSFilter(cfg).apply(kb, ks, kc)
    assert keep.tolist() == [0, 2]
    assert detect_ref.nms_class_aware(kb.numpy(), ks.numpy(), kc.numpy(), 0.5).tolist() == [0, 2]
    say("nms known-answer (test_inference.py:361-379): reference and restatement keep [0, 2]")
    out["ka/boxes"] = kb.numpy(); out["ka/scores"] = ks.numpy(); out["ka/classes"] = kc.numpy()
    out["ka/keep"] = keep.numpy()

    g = torch.Generator().manual_seed(105)
    # (2) class-agnostic, yolo_head.py:678-731
    for tag, (n, thr, cap) in {"ag300": (300, 0.5, 100), "ag1000": (1000, 0.45, 100),
                               "ag64cap5": (64, 0.3, 5), "ag1": (1, 0.5, 100)}.items():
        bc = random_boxes(g, n)
        boxes = torch.from_numpy(detect_ref.center_to_corner(bc.numpy()))
        scores = torch.rand(n, generator=g)
        assert len(torch.unique(scores)) == n
        want = head.non_max_suppression(boxes, scores, iou_threshold=thr, max_detections=cap)
        got = detect_ref.nms_agnostic(boxes.numpy(), scores.numpy(), thr, cap)
        assert got.tolist() == want.tolist(), tag
        say(f"nms_agnostic[{tag}] N={n} thr={thr} cap={cap}: restatement keep list identical ({len(got)} kept)")
        out[f"{tag}/boxes"] = boxes.numpy(); out[f"{tag}/scores"] = scores.numpy()
        out[f"{tag}/keep"] = want.numpy(); out[f"{tag}/thr"] = np.float32(thr); out[f"{tag}/cap"] = np.int64(cap)
    # (3) class-aware, postprocessing.py:505-607 (R9)
    for tag, (n, ncls, thr, cap) in {"ca400": (400, 5, 0.45, 100), "ca600": (600, 80, 0.45, 100),
                                     "ca200cap1000": (200, 3, 0.3, 1000)}.items():
        bc = random_boxes(g, n)
        scores = torch.rand(n, generator=g)
        cls = torch.randint(0, ncls, (n,), generator=g)
        cfg = ref.PostprocessingConfig(nms_iou_threshold=thr, nms_max_detections=cap)
        want = ref.NMSFilter(cfg).apply(bc, scores, cls)
        got = detect_ref.nms_class_aware(bc.numpy(), scores.numpy(), cls.numpy(), thr, cap)
        assert got.tolist() == want.tolist(), tag
        say(f"nms_class_aware[{tag}] N={n} classes={ncls} thr={thr} cap={cap}: restatement keep list identical ({len(got)} kept)")
        out[f"{tag}/boxes"] = bc.numpy(); out[f"{tag}/scores"] = scores.numpy(); out[f"{tag}/classes"] = cls.numpy()
        out[f"{tag}/keep"] = want.numpy(); out[f"{tag}/thr"] = np.float32(thr); out[f"{tag}/cap"] = np.int64(cap)
    # (4) two-stage post_process, yolo_head.py:571-676, on decoded random predictions
    b, a, nc = 2, 3, 80
    decoded_ref, decoded_mine, preds = {}, [], []
    for s, hw in enumerate((16, 8, 4)):
        pred = torch.randn(b, a, hw, hw, 5 + nc, generator=g) * 2.0
        pred[..., 2:4] *= 0.25
        d = detect_ref.yolo_decode(pred, detect_ref.anchors_wh(s))
        decoded_mine.append(d)
        decoded_ref[f"scale_{s}"] = d
        preds.append(pred)
    conf, thr, cap = 0.5, 0.45, 20
    want = head.post_process(decoded_ref, confidence_threshold=conf, iou_threshold=thr, max_detections=cap)
    got = detect_ref.post_process(decoded_mine, conf, thr, cap)
    for bi in range(b):
        assert np.array_equal(got[bi]["boxes"], want[bi]["boxes"].numpy())
        assert np.array_equal(got[bi]["scores"], want[bi]["scores"].numpy())
        assert np.array_equal(got[bi]["labels"], want[bi]["labels"].numpy())
        say(f"post_process image {bi}: restatement detections bit-equal to reference ({len(got[bi]['scores'])} kept, "
            f"per-scale {[len(k) for k in got[bi]['scale_keep']]})")
        out[f"pp/{bi}/boxes"] = want[bi]["boxes"].numpy()
        out[f"pp/{bi}/scores"] = want[bi]["scores"].numpy()
        out[f"pp/{bi}/labels"] = want[bi]["labels"].numpy()
    for s, p in enumerate(preds):
        out[f"pp/pred{s}"] = p.numpy()
    out["pp/conf"] = np.float32(conf); out["pp/thr"] = np.float32(thr); out["pp/cap"] = np.int64(cap)
    np.savez_compressed(os.path.join(GOLDEN, "nms.npz"), **out)


def gen_stability(ref):
    """_monitor_stability / get_stability_metrics (manifold_layers.py:282-341): the reference module in TRAIN mode
    (dropout 0) records eigenvalues of sym(H_res), the signal ratio and the row / column sum errors."""
    out = {}
    for tag, (d, n, t) in {"d48n2": (48, 2, 64), "d64n4": (64, 4, 33)}.items():
        torch.manual_seed(106)
        mod = ref.ManifoldHyperConnection(d, expansion_rate=n, dropout_rate=0.0).train()
        with torch.no_grad():
            mod.H_res_raw.normal_(0, 0.7)
        xs = [torch.randn(t, d) * (1 + i) for i in range(3)]
        with torch.no_grad():
            for x in xs:
                mod(x)
        m = mod.get_stability_metrics()
        for k, v in mod.state_dict().items():
            out[f"{tag}/p/{k}"] = v.detach().numpy()
        for i, x in enumerate(xs):
            out[f"{tag}/x{i}"] = x.numpy()
        for k in ("max_eigenvalue", "min_eigenvalue", "eigenvalue_range", "signal_ratio_mean", "signal_ratio_std",
                  "signal_ratio_min", "signal_ratio_max", "signal_ratio", "row_sum_error", "col_sum_error"):
            out[f"{tag}/m/{k}"] = np.float64(m[k])
        out[f"{tag}/m/final_convergence"] = np.float64(m["sk_convergence"]["final_convergence"])
        say(f"stability[{tag}]: reference metrics recorded (max eig {m['max_eigenvalue']:.6f}, signal ratio {m['signal_ratio']:.4f})")
    np.savez_compressed(os.path.join(GOLDEN, "stability.npz"), **out)


def gen_loss(ref):
    """YOLOLoss (yolo_head.py:374-465) on dense targets, including a scale without objects."""
    g = torch.Generator().manual_seed(107)
    loss = ref.yolo_head.YOLOLoss(num_classes=80)
    preds, tgts = {}, []
    for s, hw in enumerate((8, 4, 2)):
        preds[f"scale_{s}"] = torch.randn(2, 3, hw, hw, 85, generator=g)
        t = torch.zeros(2, 3, hw, hw, 85)
        if s != 1:                                          # scale 1 has no object: contributes nothing (:411-413)
            for _ in range(5):
                b, a, y, x = (int(torch.randint(0, m, (1,), generator=g)) for m in (2, 3, hw, hw))
                t[b, a, y, x, :4] = torch.rand(4, generator=g)
                t[b, a, y, x, 4] = 1.0
                t[b, a, y, x, 5 + int(torch.randint(0, 80, (1,), generator=g))] = 1.0
        tgts.append(t)
    want = loss(preds, tgts)
    out = {f"pred{s}": preds[f"scale_{s}"].numpy() for s in range(3)}
    out.update({f"tgt{s}": tgts[s].numpy() for s in range(3)})
    for k, v in want.items():
        out[k] = np.float64(float(v))
    say(f"yolo_loss: reference total {float(want['total_loss']):.6f} (coord {want['coord_loss']:.4f} obj {want['obj_loss']:.4f} "
        f"noobj {want['noobj_loss']:.4f} cls {want['cls_loss']:.4f})")
    np.savez_compressed(os.path.join(GOLDEN, "yolo_loss.npz"), **out)


def gen_hybrid():
    """Whole model: the reference's HybridVisionSystem (repairs R1-R8, reference_repaired.load_full_model) with
    name-seeded parameters on a 2 x 3 x 128 x 128 input.  The fixture holds the state_dict key / shape table and the
    outputs; the GPU test rebuilds the same parameters by name in hvs_b200's host model."""
    import json
    model = reference_repaired.load_full_model().eval()
    table = {k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in model.state_dict().items()}
    with open(os.path.join(GOLDEN, "hybrid_vision_keys.json"), "w") as f:
        json.dump(table, f, indent=0, sort_keys=True)
    reference_repaired.fill_by_name(model, 0)
    x = torch.randn(2, 3, 128, 128, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        res = model(x)
    out = {"x": x.numpy(), "final_features": res["final_features"].numpy(), "vit_features_mean": res["vit_features"].mean((2, 3)).numpy()}
    for s in range(3):
        out[f"pred{s}"] = res["predictions"][f"scale_{s}"].numpy()
    for k in ("fused_small", "fused_medium", "fused_large"):
        out[f"{k}_chanmean"] = res["fused_features"][k].mean((2, 3)).numpy()
    for k in ("scale_small", "scale_medium"):
        out[f"backbone_{k}_chanmean"] = res["backbone_features"][k].mean((2, 3)).numpy()
    n_mhc = sum(1 for m in model.modules() if type(m).__name__ == "ManifoldHyperConnection")
    say(f"hybrid_vision: reference model ({sum(p.numel() for p in model.parameters()) / 1e6:.1f} M parameters, {n_mhc} mHC modules, "
        f"{len(table)} state_dict entries) forward at 2x3x128x128 with name-seeded parameters: predictions std "
        f"{[round(float(out[f'pred{s}'].std()), 4) for s in range(3)]}")
    np.savez_compressed(os.path.join(GOLDEN, "hybrid_vision.npz"), **out)


def main():
    torch.set_num_threads(1)        # fixed reduction order -> reproducible fixtures
    os.makedirs(GOLDEN, exist_ok=True)
    ref = reference_repaired.load()
    say(f"torch {torch.__version__}; reference at {reference_repaired.REFERENCE_ROOT}")
    gen_sinkhorn(ref)
    gen_rmsnorm(ref)
    gen_module(ref)
    gen_stream(ref)
    head = gen_decode(ref)
    gen_nms(ref, head)
    gen_stability(ref)
    gen_loss(ref)
    gen_hybrid()
    with open(os.path.join(GOLDEN, "PINNING.txt"), "w") as f:
        f.write(_report.getvalue())
    say("golden fixtures written to tests/golden/")


if __name__ == "__main__":
    main()
