"""Development aid: the batch-64 decode of harness.detect_tail (three scales) run a few times -- the target of the
`ncu --set full -k regex:decode` capture of the decode kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hvs_b200 import ops
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
anchors = torch.tensor([[(10, 13), (16, 30), (33, 23)], [(30, 61), (62, 45), (59, 119)], [(116, 90), (156, 198), (373, 326)]],
                       dtype=torch.float32, device=dev) / 416.0
preds = []
for s, hw in enumerate((80, 40, 20)):
    nchw = torch.randn(64, 3 * 85, hw, hw, generator=g, device=dev) * 0.5
    preds.append(nchw.view(64, 3, 85, hw, hw).to(torch.bfloat16).permute(0, 1, 3, 4, 2))
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    ev = []
    for p, a in zip(preds, anchors):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.yolo_decode(p, a, want_scores=False, want_objectness=False); e1.record()
        ev.append((e0, e1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); dec = ops.yolo_decode_scales(preds, list(anchors), want_objectness=False); e1.record()
    ev.append((e0, e1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.post_process(dec, 0.25, 0.45, 100); e1.record()
    ev.append((e0, e1))
    torch.cuda.synchronize()
    print("ms: decode per scale | all scales in one launch | post_process:", " ".join(f"{a.elapsed_time(b):.4f}" for a, b in ev))
