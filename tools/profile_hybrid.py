"""Where the batch-64 hybrid_vision inference step spends its device time: torch.profiler (CUPTI) kernel table grouped
by kernel family.  Usage: python tools/profile_hybrid.py [batch] [train]  -> gpurun_out/hybrid_profile_<mode>.txt"""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import hvs_b200
from hvs_b200 import harness, ops

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
train = len(sys.argv) > 2 and sys.argv[2] == "train"
dev = torch.device("cuda", 0)
model = harness.build_model(dev)
if not train and os.environ.get("HVS_FOLD_BN", "1") == "1":
    from hvs_b200.hybrid_vision import to_channels_last
    harness.fold_batchnorm_for_inference(model.eval())
    to_channels_last(model)
    if os.environ.get("HVS_CAST_WEIGHTS", "1") == "1":
        harness.cast_weights_for_bf16_inference(model)
x = torch.randn(batch, 3, 640, 640, device=dev).to(torch.bfloat16 if not train else torch.float32).contiguous(memory_format=torch.channels_last)
if train:
    model.train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-5, fused=True)
    targets = harness.synthetic_targets(batch, 640, 0, dev)
    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = model(x, targets=targets, compute_loss=True)
            loss = out["loss"]["total_loss"] + 0.0 * out["final_features"].float().sum()
        loss.backward()
        opt.step()
else:
    model.eval()
    model.detection_head.want_scores = False
    def step():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            out = model(x)
            ops.post_process(list(out["decoded"].values()), 0.25, 0.45, 100)
for _ in range(2):
    step()
torch.cuda.synchronize()
STACKS = os.environ.get("HVS_PROF_STACKS", "0") == "1"
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], with_stack=STACKS) as prof:
    step()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total, e.count) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type == torch.autograd.DeviceType.CUDA]
if not rows:
    rows = [(e.key, getattr(e, "self_device_time_total", 0), e.count) for e in prof.key_averages() if getattr(e, "self_device_time_total", 0) > 0]
rows.sort(key=lambda r: -r[1])
total = sum(r[1] for r in rows)
fam = collections.Counter()
def family(k):
    kl = k.lower()
    for name, pats in (("hvs k2_gemm", ["k2_gemm"]), ("hvs layernorm/rmsnorm", ["layernorm_fwd", "rmsnorm"]), ("hvs coeffs", ["static_coeffs"]),
                       ("hvs decode/nms", ["decode_cell", "nms_kernel", "post_process", "gather", "compact"]),
                       ("conv (cudnn/cutlass)", ["conv", "cudnn", "implicit", "xmma", "sm90", "sm100", "cutlass", "wgrad", "dgrad", "fprop"]),
                       ("gemm (cublas)", ["gemm", "cublas", "nvjet"]), ("batchnorm", ["batch_norm", "bn_"]),
                       ("elementwise/copy", ["elementwise", "copy", "vectorized", "fill"]), ("reduce/pool/softmax", ["reduce", "pool", "softmax"])):
        if any(p in kl for p in pats):
            return name
    return "other"
for k, t, c in rows:
    fam[family(k)] += t
out = [f"mode={'train' if train else 'infer'} batch={batch}: total device time {total/1e3:.2f} ms over {sum(r[2] for r in rows)} kernels"]
for name, t in fam.most_common():
    out.append(f"  {name:28s} {t/1e3:9.2f} ms  {100*t/total:5.1f} %")
out.append("top kernels:")
for k, t, c in rows[:40]:
    out.append(f"  {t/1e3:9.3f} ms  x{c:<5d} {k[:150]}")
os.makedirs("gpurun_out", exist_ok=True)
path = f"gpurun_out/hybrid_profile_{'train' if train else 'infer'}_b{batch}.txt"
open(path, "w").write("\n".join(out) + "\n")
print("\n".join(out[:14]))
if STACKS:
    # which call sites launch the copy / elementwise kernels: aggregate operator device time by (op, innermost repo / torch frames)
    agg = collections.Counter(); cnt = collections.Counter()
    for e in prof.key_averages(group_by_stack_n=8):
        if e.key in ("aten::copy_", "aten::add", "aten::add_", "aten::mul", "aten::contiguous", "aten::to", "aten::_to_copy", "aten::sum", "aten::clone", "aten::linalg_vector_norm") and e.device_time_total > 0:
            frames = [f for f in e.stack if "hvs_b200" in f or "humanoid-vision" in f or "harness" in f][:3]
            k = (e.key, " <- ".join(f.strip()[-90:] for f in frames))
            agg[k] += e.self_device_time_total; cnt[k] += e.count
    rows2 = [f"  {t/1e3:8.3f} ms x{cnt[k]:<5d} {k[0]:24s} {k[1]}" for k, t in agg.most_common(40)]
    open(path.replace(".txt", "_stacks.txt"), "w").write("\n".join(rows2) + "\n")
    print("\n".join(rows2[:25]))
