"""Stage-by-stage comparison of the detection tail against the oracle on GPU-decoded tensors (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hvs_b200
from oracle import detect_ref

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
preds, awh = [], []
for s, hw in enumerate((80, 40, 20)):
    nchw = torch.randn(B, 3 * 85, hw, hw, generator=g, device=dev) * 0.5
    preds.append(nchw.view(B, 3, 85, hw, hw).permute(0, 1, 3, 4, 2))
    awh.append(detect_ref.anchors_wh(s).to(dev))
d_a = [hvs_b200.ops.yolo_decode(p, a, want_scores=False, want_objectness=False) for p, a in zip(preds, awh)]
d_b = [hvs_b200.ops.yolo_decode(p, a) for p, a in zip(preds, awh)]
for s in range(3):
    for k in ("boxes", "class_scores", "class_indices"):
        print("scale", s, k, "lean==full", bool((d_a[s][k] == d_b[s][k]).all()))
out = hvs_b200.ops.post_process(d_a, 0.25, 0.45, 100)
torch.cuda.synchronize()
sel = [0, 1, B - 1]
cpu = [{k: v[sel].cpu() for k, v in d.items()} for d in d_a]
want = detect_ref.post_process(cpu, 0.25, 0.45, 100)
for b, bb in enumerate(sel):
    k = int(out[3][bb])
    gb = out[0][bb, :k].cpu().numpy()
    wb = want[b]["boxes"]
    print("image", b, "k", k, len(wb), "equal", gb.shape == wb.shape and bool((gb == wb).all()))
    if gb.shape == wb.shape and not (gb == wb).all():
        bad = np.nonzero((gb != wb).any(axis=1))[0]
        print("  first differing ranks", bad[:10], "scores gpu", out[1][bb, bad[:5]].cpu().numpy(), "want", want[b]["scores"][bad[:5]])
    # stage 1 per scale
    for s in range(3):
        bx = cpu[s]["boxes"][b].reshape(-1, 4); sc = cpu[s]["class_scores"][b].reshape(-1)
        ki, ks, kc = hvs_b200.ops.nms(bx.to(dev), sc.to(dev), iou_threshold=0.45, max_detections=100, score_threshold=0.25)
        got = ki[0, :int(kc[0])].cpu().numpy()
        w = want[b]["scale_keep"][s]
        same = len(got) == len(w) and bool((got == w).all())
        print("   stage1 scale", s, "n>thr", int((sc > 0.25).sum()), "same", same)
        if not same:
            m = min(len(got), len(w)); d = np.nonzero(got[:m] != w[:m])[0]
            print("     first diff rank", d[:5], got[d[:5]], w[d[:5]])
            msk = (sc > 0.25).numpy(); cs = sc.numpy()[msk]
            for r in d[:3]:
                print("     scores", cs[got[r]], cs[w[r]], "ties with winner:", int((cs == cs[w[r]]).sum()))
