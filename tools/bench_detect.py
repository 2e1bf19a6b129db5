"""Detection tail at BASELINE config-3 sizes: batch 64, 640x640 -> grids 80/40/20, A=3, 80 classes.
Times decode (reading the head's strided NCHW view in place) + the two-stage multi-scale NMS on the device
with CUDA events, and the CPU oracle on a bounded sample.  Prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hvs_b200
from oracle import detect_ref

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
obj_bias = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0      # 0.0 = SURVEY D18 worst case (all candidates ~0.25); -4 = realistic
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
preds, awh = [], []
for s, hw in enumerate((80, 40, 20)):
    nchw = torch.randn(B, 3 * 85, hw, hw, generator=g, device=dev) * 0.5          # pred_conv output layout
    v = nchw.view(B, 3, 85, hw, hw)
    v[:, :, 4] += obj_bias
    preds.append(v.permute(0, 1, 3, 4, 2))                                       # [B,A,H,W,85] strided view
    awh.append(detect_ref.anchors_wh(s).to(dev))

def step():
    dec = [hvs_b200.ops.yolo_decode(p, a, want_scores=False, want_objectness=False) for p, a in zip(preds, awh)]
    return hvs_b200.ops.post_process(dec, 0.25, 0.45, 100)

for _ in range(3): out = step()
torch.cuda.synchronize()
ts = []
for _ in range(10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); out = step(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
ts.sort()
ms = ts[len(ts) // 2]
in_bytes = sum(p.numel() * 4 for p in preds)
cand = int(sum((hvs_b200.ops.yolo_decode(p, a)["class_scores"] > 0.25).sum() for p, a in zip(preds, awh)))
# CPU oracle on 2 images
nb = 2
t0 = time.perf_counter()
dec_cpu = [detect_ref.yolo_decode(p[:nb].cpu(), a.cpu()) for p, a in zip(preds, awh)]
detect_ref.post_process(dec_cpu, 0.25, 0.45, 100)
cpu_s = (time.perf_counter() - t0) / nb
# keep-set parity is defined on the SAME decoded tensors (sigmoid differs in the last ulp between CPU and GPU)
dec_same = [{k: v[:nb].cpu() for k, v in hvs_b200.ops.yolo_decode(p, a).items()} for p, a in zip(preds, awh)]
want = detect_ref.post_process(dec_same, 0.25, 0.45, 100)
k = int(out[3][0])
same = bool((out[0][0, :k].cpu().numpy() == want[0]["boxes"]).all()) if k == len(want[0]["scores"]) else False
print(json.dumps({"workload": f"decode + two-stage NMS, batch {B}, 640x640 grids 80/40/20, 80 classes, conf 0.25 iou 0.45 max 100, objectness bias {obj_bias}",
                  "ms_per_batch": ms, "img_per_s": B / ms * 1e3, "decode_input_GB": in_bytes / 1e9,
                  "candidates_over_threshold_per_image": cand / B, "kept_image0": k,
                  "cpu_oracle_s_per_image": cpu_s, "cpu_img_per_s": 1 / cpu_s, "cpu_threads": torch.get_num_threads(),
                  "keep_set_matches_oracle_image0": same}))
