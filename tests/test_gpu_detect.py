"""Decode (tolerance) and NMS / post_process (bit-exact keep indices) on the GPU vs oracle and golden."""
import numpy as np
import pytest
import torch

from oracle import detect_ref

pytestmark = pytest.mark.gpu


def random_boxes_cxcywh(g, n):
    cx, cy = torch.rand(n, generator=g), torch.rand(n, generator=g)
    w, h = 0.05 + 0.3 * torch.rand(n, generator=g), 0.05 + 0.3 * torch.rand(n, generator=g)
    return torch.stack([cx, cy, w, h], -1)


def gpu_keep(boxes, scores, classes=None, **kw):
    import hvs_b200
    ki, ks, kc = hvs_b200.ops.nms(torch.as_tensor(boxes).cuda(), torch.as_tensor(scores).cuda(),
                                  None if classes is None else torch.as_tensor(classes).cuda(), **kw)
    n = int(kc[0])
    return ki[0, :n].cpu().tolist(), ks[0, :n].cpu().tolist()


def test_decode_golden_and_layouts(golden):
    import hvs_b200
    g = golden("decode")
    pred = torch.from_numpy(g["pred"]).cuda()                       # [B,A,H,W,85] channel-contiguous
    b, a, h, w, d = pred.shape
    nchw = pred.permute(0, 1, 4, 2, 3).contiguous()                 # conv layout [B,A,85,H,W]
    strided = nchw.permute(0, 1, 3, 4, 2)                           # the head's non-contiguous view
    assert not strided.is_contiguous()
    for s in range(3):
        awh = torch.from_numpy(g[f"s{s}/anchor_wh"]).cuda()
        for p in (pred, strided):
            out = hvs_b200.ops.yolo_decode(p, awh, want_scores=True)
            assert torch.allclose(out["boxes"].cpu(), torch.from_numpy(g[f"s{s}/boxes"]), rtol=2e-6, atol=2e-7)
            assert torch.allclose(out["class_scores"].cpu(), torch.from_numpy(g[f"s{s}/class_scores"]), rtol=2e-6, atol=1e-8)
            assert torch.allclose(out["objectness"].cpu(), torch.from_numpy(g[f"s{s}/objectness"]), rtol=2e-6)
            assert torch.equal(out["class_indices"].cpu(), torch.from_numpy(g[f"s{s}/class_indices"]))
            ref = detect_ref.yolo_decode(pred.cpu(), awh.cpu())
            assert torch.allclose(out["scores"].cpu(), ref["scores"], rtol=2e-6, atol=1e-8)
            # the detection path's call (no [cells, C] score tensor: the first-maximum kernel on the strided view)
            lean = hvs_b200.ops.yolo_decode(p, awh)
            for k in ("boxes", "class_scores", "class_indices", "objectness"):
                assert torch.equal(lean[k], out[k]), k


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_decode_half_inputs(dtype):
    import hvs_b200
    g = torch.Generator().manual_seed(7)
    pred = (torch.randn(2, 3, 20, 20, 85, generator=g) * 1.5).to(dtype)
    awh = detect_ref.anchors_wh(2)
    out = hvs_b200.ops.yolo_decode(pred.cuda(), awh.cuda())
    ref = detect_ref.yolo_decode(pred, awh)
    assert torch.allclose(out["boxes"].cpu(), ref["boxes"], rtol=2e-6, atol=2e-7)
    assert torch.allclose(out["class_scores"].cpu(), ref["class_scores"], rtol=2e-6, atol=1e-8)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_decode_first_maximum_kernel_is_bit_identical_to_the_all_scores_kernel(dtype):
    """decode_four_cells_first_max finds the winning class on the logits and evaluates one sigmoid per cell; the
    reference (yolo_head.py:282-285) takes max / first argmax over all obj * sigmoid(cls).  Same bits, same indices --
    on random logits and on the cases where rounding makes scores tie although the logits differ: saturated sigmoids,
    logits a few ulp apart, duplicates of the maximum, objectness so small that products fall into the denormals,
    NaN / +-inf logits."""
    import hvs_b200
    g = torch.Generator().manual_seed(11)
    b, a, h, w, c = 3, 3, 12, 16, 80
    pred = torch.randn(b, a, 5 + c, h, w, generator=g) * 2.0
    cls = pred[:, :, 5:]
    # cls[b, a, :, row] is the [C, W] slab of one grid row
    cls[0, 0, :, 0] = 15.0 + torch.randn(c, w, generator=g) * 4.0                        # saturated: ties between different logits
    cls[0, 1, :, 1] = 1.0 + torch.randint(0, 4, (c, w), generator=g) * 2.0 ** -20        # a few fp32 ulp apart
    cls[0, 2, :, 2] = torch.randint(-2, 3, (c, w), generator=g).float()                  # many duplicates of the maximum
    cls[1, 0, :, 3] = 0.0
    cls[1, 1, :, 4] = -95.0 + torch.randn(c, w, generator=g)                             # sigmoid in the denormals
    pred[1, 2, 4, 5] = -80.0                                                             # objectness ~ 1e-35: products underflow
    pred[1, 2, 4, 6] = -110.0                                                            # objectness 0: every score 0, index 0
    cls[2, 0, 7, 7] = float("nan")
    cls[2, 0, 0, 8] = float("nan")                                                       # NaN in class 0 wins (c == 0 is taken unseen)
    pred[2, 1, 4, 9] = float("nan")                                                      # NaN objectness
    cls[2, 1, 9, 10] = float("inf")
    cls[2, 1, 3, 10] = 60.0                                                              # sigmoid(60) == sigmoid(inf) == 1: first index wins
    cls[2, 2, :, 11] = float("-inf")
    cls[2, 2, 5, 0] = float("-inf")
    cls[2, 0, :, 9] = 100.0 * torch.randn(c, w, generator=g)
    pred = pred.to(dtype).cuda()
    view = pred.permute(0, 1, 3, 4, 2)                                                   # [B,A,H,W,85], channel stride H*W
    awh = detect_ref.anchors_wh(1).cuda()
    full = hvs_b200.ops.yolo_decode(view, awh, want_scores=True)
    lean = hvs_b200.ops.yolo_decode(view, awh)
    assert torch.equal(lean["class_indices"], full["class_indices"])
    assert torch.equal(lean["class_scores"].view(torch.int32), full["class_scores"].view(torch.int32))
    assert torch.equal(lean["boxes"].view(torch.int32), full["boxes"].view(torch.int32))
    assert torch.equal(lean["objectness"].view(torch.int32), full["objectness"].view(torch.int32))
    # and the all-scores kernel is the reference rule: max / first argmax of its own score tensor
    sc = full["scores"]
    finite = ~sc.isnan().any(-1)
    assert torch.equal(full["class_scores"][finite], sc.max(-1).values[finite])


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_decode_all_scales_in_one_launch_equals_scale_by_scale(dtype):
    """hvs_yolo_decode_scales (the head's decode loop, yolo_head.py:536-555, as one launch) against hvs_yolo_decode per
    scale: identical bits; also when a scale forces the scale-by-scale fallback (channel-contiguous layout)."""
    import hvs_b200
    g = torch.Generator().manual_seed(5)
    preds, awhs = [], []
    for s, hw in enumerate((40, 20, 12)):
        nchw = (torch.randn(3, 3 * 85, hw, hw, generator=g) * 1.5).to(dtype).cuda()
        preds.append(nchw.view(3, 3, 85, hw, hw).permute(0, 1, 3, 4, 2))
        awhs.append(detect_ref.anchors_wh(s).cuda())
    for variant in ("strided", "one contiguous"):
        ps = list(preds)
        if variant == "one contiguous":
            ps[1] = ps[1].contiguous()
        together = hvs_b200.ops.yolo_decode_scales(ps, awhs)
        for p, a, t in zip(ps, awhs, together):
            one = hvs_b200.ops.yolo_decode(p, a)
            for k in ("boxes", "class_scores", "objectness"):
                assert torch.equal(t[k].view(torch.int32), one[k].view(torch.int32)), (variant, k)
            assert torch.equal(t["class_indices"], one["class_indices"]), variant
            ref = detect_ref.yolo_decode(p.float().cpu(), a.cpu())
            assert torch.allclose(t["boxes"].cpu(), ref["boxes"], rtol=2e-6, atol=2e-7)
            assert torch.allclose(t["class_scores"].cpu(), ref["class_scores"], rtol=2e-6, atol=1e-8)
    with pytest.raises(hvs_b200.HvsError):
        hvs_b200.ops.yolo_decode_scales(preds, awhs[:2])
    with pytest.raises(hvs_b200.HvsError):
        hvs_b200.ops.yolo_decode_scales([preds[0], preds[1][:2]], awhs[:2])


def test_nms_golden_bit_exact(golden):
    g = golden("nms")
    ki, _ = gpu_keep(g["ka/boxes"], g["ka/scores"], g["ka/classes"], iou_threshold=0.5, class_aware=True, boxes_xyxy=False)
    assert ki == [0, 2]                                   # reference test_inference.py:361-379
    for tag in ("ag300", "ag1000", "ag64cap5", "ag1"):
        ki, _ = gpu_keep(g[f"{tag}/boxes"], g[f"{tag}/scores"], iou_threshold=float(g[f"{tag}/thr"]),
                         max_detections=int(g[f"{tag}/cap"]))
        assert ki == g[f"{tag}/keep"].tolist(), tag
    for tag in ("ca400", "ca600", "ca200cap1000"):
        ki, _ = gpu_keep(g[f"{tag}/boxes"], g[f"{tag}/scores"], g[f"{tag}/classes"], iou_threshold=float(g[f"{tag}/thr"]),
                         max_detections=int(g[f"{tag}/cap"]), class_aware=True, boxes_xyxy=False)
        assert ki == g[f"{tag}/keep"].tolist(), tag


@pytest.mark.parametrize("n,thr", [(2, 0.5), (33, 0.3), (2500, 0.45), (20000, 0.6)])
def test_nms_agnostic_vs_oracle(n, thr):
    g = torch.Generator().manual_seed(n)
    boxes = detect_ref.center_to_corner(random_boxes_cxcywh(g, n).numpy())
    scores = torch.rand(n, generator=g).numpy()
    want = detect_ref.nms_agnostic(boxes, scores, thr, 100).tolist()
    ki, ks = gpu_keep(boxes, scores, iou_threshold=thr, max_detections=100)
    assert ki == want and ks == want


def test_nms_threshold_compaction_indices():
    g = torch.Generator().manual_seed(1)
    n = 5000
    boxes = detect_ref.center_to_corner(random_boxes_cxcywh(g, n).numpy())
    scores = torch.rand(n, generator=g).numpy()
    mask = scores > np.float32(0.7)
    want = detect_ref.nms_agnostic(boxes[mask], scores[mask], 0.45, 50).tolist()
    ki, ks = gpu_keep(boxes, scores, iou_threshold=0.45, max_detections=50, score_threshold=0.7)
    assert ki == want                                     # index into the compacted list (yolo_head.py:620-629)
    assert ks == np.nonzero(mask)[0][want].tolist()       # index into the dense list


def test_nms_edge_cases():
    import hvs_b200
    # empty set
    ki, ks, kc = hvs_b200.ops.nms(torch.zeros(0, 4).cuda(), torch.zeros(0).cuda())
    assert int(kc[0]) == 0
    # all identical boxes
    b = np.tile(np.array([[0.1, 0.1, 0.5, 0.5]], np.float32), (40, 1))
    s = np.linspace(0.1, 0.9, 40).astype(np.float32)
    assert gpu_keep(b, s)[0] == [39]
    assert gpu_keep(b, s, np.arange(40) % 2, class_aware=True)[0] == [39, 38]
    # equal scores: lower index first
    b2 = detect_ref.center_to_corner(random_boxes_cxcywh(torch.Generator().manual_seed(3), 64).numpy())
    s2 = np.full(64, 0.5, np.float32)
    assert gpu_keep(b2, s2, iou_threshold=0.5)[0] == detect_ref.nms_agnostic(b2, s2, 0.5).tolist()
    # NaN box: IoU is NaN -> suppressed by the agnostic rule (iou < thr fails), kept by the class-aware one
    b3 = np.array([[0.1, 0.1, 0.4, 0.4], [np.nan, 0.1, 0.4, 0.4], [0.6, 0.6, 0.9, 0.9]], np.float32)
    s3 = np.array([0.9, 0.8, 0.7], np.float32)
    assert gpu_keep(b3, s3)[0] == detect_ref.nms_agnostic(b3, s3).tolist() == [0, 2]
    assert gpu_keep(b3, s3, np.zeros(3, np.int64), class_aware=True)[0] == \
        detect_ref.nms_class_aware(b3, s3, np.zeros(3), 0.5, boxes_are_corners=True).tolist() == [0, 1, 2]
    # nothing passes the score threshold
    assert gpu_keep(b3, s3, score_threshold=0.95)[0] == []


def test_nms_batched_offsets():
    import hvs_b200
    g = torch.Generator().manual_seed(4)
    sizes = [0, 17, 300, 1, 2048]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    n = int(offs[-1])
    boxes = detect_ref.center_to_corner(random_boxes_cxcywh(g, n).numpy())
    scores = torch.rand(n, generator=g).numpy()
    cls = torch.randint(0, 7, (n,), generator=g).numpy()
    ki, ks, kc = hvs_b200.ops.nms(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(), torch.from_numpy(cls).cuda(),
                                  iou_threshold=0.45, max_detections=100, class_aware=True, boxes_xyxy=True,
                                  offsets=torch.from_numpy(offs).cuda(), max_n=max(sizes))
    for p, sz in enumerate(sizes):
        lo, hi = offs[p], offs[p + 1]
        want = detect_ref.nms_class_aware(boxes[lo:hi], scores[lo:hi], cls[lo:hi], 0.45, 100, boxes_are_corners=True).tolist()
        assert ki[p, :int(kc[p])].cpu().tolist() == want


def test_post_process_golden_and_oracle(golden):
    import hvs_b200
    g = golden("nms")
    conf, thr, cap = float(g["pp/conf"]), float(g["pp/thr"]), int(g["pp/cap"])
    # same decoded tensors for both sides ("on the same inputs"): decode on the CPU oracle
    decoded = [detect_ref.yolo_decode(torch.from_numpy(g[f"pp/pred{s}"]), detect_ref.anchors_wh(s)) for s in range(3)]
    dev = [{k: v.cuda() for k, v in d.items() if k in ("boxes", "class_scores", "class_indices")} for d in decoded]
    db, ds, dl, dc = hvs_b200.ops.post_process(dev, conf, thr, cap)
    for b in range(2):
        k = int(dc[b])
        assert np.array_equal(db[b, :k].cpu().numpy(), g[f"pp/{b}/boxes"])
        assert np.array_equal(ds[b, :k].cpu().numpy(), g[f"pp/{b}/scores"])
        assert np.array_equal(dl[b, :k].cpu().numpy(), g[f"pp/{b}/labels"])


def test_post_process_full_size_worst_case():
    """640x640 grids (80/40/20), every candidate above the threshold (SURVEY D18 worst case)."""
    import hvs_b200
    g = torch.Generator().manual_seed(8)
    decoded = []
    for s, hw in enumerate((80, 40, 20)):
        pred = torch.randn(2, 3, hw, hw, 85, generator=g)
        pred[..., 2:4] *= 0.3
        decoded.append(detect_ref.yolo_decode(pred, detect_ref.anchors_wh(s)))
    want = detect_ref.post_process(decoded, 0.05, 0.45, 100)
    dev = [{k: v.cuda() for k, v in d.items() if k in ("boxes", "class_scores", "class_indices")} for d in decoded]
    db, ds, dl, dc = hvs_b200.ops.post_process(dev, 0.05, 0.45, 100)
    for b in range(2):
        k = int(dc[b])
        assert k == len(want[b]["scores"])
        assert np.array_equal(db[b, :k].cpu().numpy(), want[b]["boxes"])
        assert np.array_equal(dl[b, :k].cpu().numpy(), want[b]["labels"])


def test_fused_prediction_conv_decode_matches_oracle():
    """hvs_head_decode_fused: 1x1 prediction conv as a tcgen05 GEMM with the decode as its epilogue,
    against the oracle decode of the same logits (fp64 GEMM of the same bf16 operands)."""
    import hvs_b200
    for b, h, w, cin in ((2, 20, 20, 1024), (3, 40, 40, 512), (1, 80, 80, 256), (2, 7, 5, 64)):
        g = torch.Generator().manual_seed(h * w + cin)
        tok = (torch.randn(b * h * w, cin, generator=g)).to(torch.bfloat16)
        wt = (torch.randn(255, cin, generator=g) * (2.0 / cin ** 0.5)).to(torch.bfloat16)
        bias = torch.randn(255, generator=g) * 0.5
        bias[4::85] -= 1.0
        w256 = torch.zeros(256, cin, dtype=torch.bfloat16); w256[:255] = wt
        b256 = torch.zeros(256); b256[:255] = bias
        for s in range(3):
            awh = detect_ref.anchors_wh(s)
            got = hvs_b200.ops.head_decode_fused(tok.cuda(), w256.cuda(), b256.cuda(), awh.cuda(), b, h, w, want_objectness=True)
            logits = (tok.double() @ wt.double().t() + bias.double()).float()               # [B*H*W, 255]
            pred = logits.reshape(b, h, w, 3, 85).permute(0, 3, 1, 2, 4).contiguous()        # [B, A, H, W, 85]
            want = detect_ref.yolo_decode(pred, awh)
            assert torch.allclose(got["boxes"].cpu(), want["boxes"], rtol=3e-4, atol=1e-5)   # exp(tw) turns the logits' 1e-6 accumulation-order error into a relative one
            assert torch.allclose(got["class_scores"].cpu(), want["class_scores"], rtol=1e-4, atol=1e-7)
            assert torch.allclose(got["objectness"].cpu(), want["objectness"], rtol=1e-5, atol=1e-7)
            same = (got["class_indices"].cpu() == want["class_indices"]).float().mean().item()
            assert same > 0.999, same                       # fp32 accumulation order may flip an exact near-tie


def test_detection_head_fused_tail_equals_two_kernel_tail():
    """The head with fuse_pred_decode: same detections as conv -> decode kernel, no raw prediction tensor."""
    import hvs_b200
    torch.manual_seed(5)
    head = hvs_b200.YOLODetectionHead([64, 128, 256], num_classes=80, use_mhc=False).cuda().eval()
    with torch.no_grad():
        for hd in head.pred_heads:
            hd.pred_conv.weight.normal_(0, 0.2)
            # make conv outputs exactly representable paths comparable: run the reference path on bf16-rounded operands too
            hd.pred_conv.weight.copy_(hd.pred_conv.weight.to(torch.bfloat16).float())
    feats = {"scale_small": torch.randn(2, 64, 16, 16, device="cuda"), "scale_medium": torch.randn(2, 128, 8, 8, device="cuda"),
             "scale_large": torch.randn(2, 256, 4, 4, device="cuda")}
    with torch.no_grad():
        ref = head(feats)
        head.fuse_pred_decode = True
        before = hvs_b200._lib.launch_count()
        out = head(feats)
        assert hvs_b200._lib.launch_count() - before == 3 and out["predictions"] == {}
    for s in range(3):
        a, b = out["decoded"][f"scale_{s}"], ref["decoded"][f"scale_{s}"]
        # the fused path rounds the 1x1 conv's INPUT to bf16 (2^-9 relative on logits of magnitude ~3), the two-kernel path here is fp32
        ca, cb = (a["boxes"][..., :2] + a["boxes"][..., 2:]) / 2, (b["boxes"][..., :2] + b["boxes"][..., 2:]) / 2
        sa, sb = a["boxes"][..., 2:] - a["boxes"][..., :2], b["boxes"][..., 2:] - b["boxes"][..., :2]
        assert (ca - cb).abs().max() < 5e-3 and ((sa - sb).abs() / sb).max() < 0.1
        assert (a["class_scores"] - b["class_scores"]).abs().max() < 3e-2
    dets = head.post_process(out["decoded"], 0.2, 0.45, 50)
    assert len(dets) == 2 and all(d["boxes"].shape[1] == 4 for d in dets)


@pytest.mark.parametrize("n,cap,thr", [(2049, 100, 0.45), (5000, 100, 0.5), (19200, 100, 0.45), (24000, 1000, 0.3), (3000, 7, 0.45)])
def test_sorted_nms_kernel_large_sets_bit_exact(n, cap, thr):
    """nms_sorted_kernel (sets of more than 2048 candidates): keep lists identical to the oracle for both reference
    semantics, with a score threshold, duplicated scores (ties -> lower index) and NaN scores / boxes in the input."""
    import hvs_b200
    g = torch.Generator().manual_seed(n)
    c = torch.rand(n, 2, generator=g)
    wh = 0.02 + 0.2 * torch.rand(n, 2, generator=g)
    cxcywh = torch.cat([c, wh], 1)
    xyxy = torch.from_numpy(detect_ref.center_to_corner(cxcywh.numpy()))
    scores = torch.rand(n, generator=g)
    scores[n // 3: n // 3 + 50] = scores[5]                      # a run of exact ties
    scores[7] = float("nan")
    xyxy[11, 2] = float("nan")
    cls = torch.randint(0, 5, (n,), generator=g)
    st = 0.3
    keep_idx, keep_src, cnt = hvs_b200.ops.nms(xyxy.cuda(), scores.cuda(), None, thr, cap, score_threshold=st)
    m = scores > st
    want = detect_ref.nms_agnostic(xyxy[m].numpy(), scores[m].numpy(), thr, cap)
    k = int(cnt[0])
    assert keep_idx[0, :k].cpu().tolist() == want.tolist()
    assert keep_src[0, :k].cpu().tolist() == torch.nonzero(m).flatten()[torch.from_numpy(want)].tolist()
    kc, _, cc = hvs_b200.ops.nms(cxcywh.cuda(), scores.cuda(), cls.cuda(), thr, cap, score_threshold=st, class_aware=True, boxes_xyxy=False)
    wantc = detect_ref.nms_class_aware(cxcywh[m].numpy(), scores[m].numpy(), cls[m].numpy(), thr, cap)
    assert kc[0, :int(cc[0])].cpu().tolist() == wantc.tolist()
    a2 = hvs_b200.ops.nms(xyxy.cuda(), scores.cuda(), None, thr, cap, score_threshold=st)
    assert torch.equal(a2[0][0, :k], keep_idx[0, :k])
