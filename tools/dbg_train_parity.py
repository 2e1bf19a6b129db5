"""Per-parameter gradient agreement of one whole-model training step: kernels vs library path vs library path again."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hvs_b200
from hvs_b200 import harness
from oracle import reference_repaired
DEV = "cuda:0"
size = int(sys.argv[1]) if len(sys.argv) > 1 else 320
torch.manual_seed(0)
m = hvs_b200.HybridVisionSystem({"num_classes": 80, "image_size": size})
reference_repaired.fill_by_name(m, 0)
m = m.to(DEV).train()
hvs_b200.hybrid_vision.to_channels_last(m)
for mod in m.modules():
    if isinstance(mod, (torch.nn.Dropout, torch.nn.Dropout2d)):
        mod.p = 0.0
g = torch.Generator().manual_seed(5)
x = torch.randn(2, 3, size, size, generator=g).to(DEV).contiguous(memory_format=torch.channels_last)
targets = harness.synthetic_targets(2, size, 0, DEV)
bn_state = {k: v.clone() for k, v in m.state_dict().items() if "running_" in k or "num_batches" in k}
acts = {}
def step(kernels, fp32=False):
    for mod in m.modules():
        if isinstance(mod, hvs_b200.ManifoldHyperConnection):
            mod.use_training_kernels = kernels
            mod.use_mixed_precision = not fp32
            mod.dtype = torch.float32 if fp32 else torch.bfloat16
    m.load_state_dict(bn_state, strict=False)
    for p in m.parameters():
        p.grad = None
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=not fp32):
        out = m(x, targets=targets, compute_loss=True)
        loss = out["loss"]["total_loss"] + 0.0 * out["final_features"].float().sum()
    loss.backward()
    return float(loss), {k: p.grad.detach().double().clone() for k, p in m.named_parameters() if p.grad is not None}, {k: v.detach().float().clone() for k, v in out["predictions"].items()}
def rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-300))
runs = {"kern": step(True), "lib": step(False), "lib2": step(False), "fp32": step(False, fp32=True)}
print("losses", {k: round(v[0], 4) for k, v in runs.items()})
for a, b in (("kern", "lib"), ("lib", "lib2"), ("kern", "fp32"), ("lib", "fp32")):
    print(a, "vs", b, "predictions:", {k: round(rel(runs[a][2][k].double(), runs[b][2][k].double()), 4) for k in runs[a][2]})
names = list(runs["kern"][1])
pick = [n for n in names if n.startswith("detection_head.pred_heads.0")] + names[:12]
for n in pick:
    print(f"{n:70s} kern~lib {rel(runs['kern'][1][n], runs['lib'][1][n]):.3f}  lib~lib2 {rel(runs['lib'][1][n], runs['lib2'][1][n]):.3f}  kern~fp32 {rel(runs['kern'][1][n], runs['fp32'][1][n]):.3f}  lib~fp32 {rel(runs['lib'][1][n], runs['fp32'][1][n]):.3f}  |g| {float(runs['fp32'][1][n].norm()):.3e}")
tot = lambda a, b: (sum(float((runs[a][1][n] - runs[b][1][n]).pow(2).sum()) for n in names) / sum(float(runs[b][1][n].pow(2).sum()) for n in names)) ** 0.5
print("all parameters:", {f"{a}~{b}": round(tot(a, b), 4) for a, b in (("kern", "lib"), ("lib", "lib2"), ("kern", "fp32"), ("lib", "fp32"))})
