"""Per-GEMM timing of the K2 training step (forward 4 + data-gradient 5 + weight-gradient 5 launches per layer) for every
(tokens, D, H) the model has at a given batch.  Usage: python tools/bench_k2_train.py [batch] [filter_D]
-> gpurun_out/k2_train_shapes.txt"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hvs_b200
from hvs_b200 import ops, _lib

SHAPES = [(102400, 32, 128, 2), (102400, 64, 256, 1), (25600, 32, 128, 1), (25600, 64, 256, 3), (6400, 64, 256, 2), (6400, 128, 512, 6),
          (6400, 256, 512, 2), (1600, 128, 512, 3), (1600, 256, 1024, 8), (1600, 256, 512, 1), (1600, 512, 1024, 1), (400, 256, 1024, 1),
          (400, 512, 2048, 4), (400, 256, 512, 2), (400, 512, 1024, 1), (400, 1024, 2048, 1), (401, 256, 512, 36)]
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
only_d = int(sys.argv[2]) if len(sys.argv) > 2 else None
dev = torch.device("cuda", 0)
bf = torch.bfloat16


def timeit(fn, reps=5):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


lines = []
tot = {}
for (t1, d, h, count) in SHAPES:
    if only_d is not None and d != only_d:
        continue
    t = t1 * batch
    r = lambda *s: torch.randn(*s, device=dev).to(bf)
    xn, h0, z1, a1, z2, a2, dpre = r(t, d), r(t, h), r(t, 2 * h), r(t, 2 * h), r(t, h), r(t, h), r(t, d)
    dz2, dz1, dh0 = r(t, h), r(t, 2 * h), r(t, h)
    hpre_t, hpost_t, hres_t, w1, w2 = r(h, d), r(d, h), r(d, d), r(2 * h, h), r(h, 2 * h)
    b1, b2 = torch.zeros(2 * h, device=dev), torch.zeros(h, device=dev)
    S, G = _lib.HVS_GEMM_EPI_BIAS_GELU_SAVE, _lib.HVS_GEMM_EPI_DGELU
    calls = [
        ("f.h0   xn@Hpre", 2 * t * d * h, lambda: ops.gemm_bf16(xn, hpre_t)),
        ("f.z1   gelu_save", 2 * t * h * 2 * h, lambda: ops.gemm_bf16_ex(h0, w1, bias=b1, epilogue=S, dropout_p=0.1, dropout_seed=1)),
        ("f.z2   gelu_save", 2 * t * h * 2 * h, lambda: ops.gemm_bf16_ex(a1, w2, bias=b2, epilogue=S, dropout_p=0.1, dropout_seed=2)),
        ("f.pre  dual", 2 * t * (h * d + d * d), lambda: ops.gemm_bf16(a2, hpost_t, xn, hres_t, out_dtype=bf)),
        ("d.dz2  dgelu", 2 * t * d * h, lambda: ops.gemm_bf16_ex(dpre, hpost_t, b_mn=True, epilogue=G, aux=z2, dropout_p=0.1, dropout_seed=2)),
        ("d.dxres", 2 * t * d * d, lambda: ops.gemm_bf16_ex(dpre, hres_t, b_mn=True)),
        ("d.dz1  dgelu", 2 * t * h * 2 * h, lambda: ops.gemm_bf16_ex(dz2, w2, b_mn=True, epilogue=G, aux=z1, dropout_p=0.1, dropout_seed=1)),
        ("d.dh0", 2 * t * h * 2 * h, lambda: ops.gemm_bf16_ex(dz1, w1, b_mn=True)),
        ("d.dxn", 2 * t * d * h, lambda: ops.gemm_bf16_ex(dh0, hpre_t, b_mn=True)),
        ("w.dHpost", 2 * t * d * h, lambda: ops.gemm_wgrad(a2, dpre)),
        ("w.dHres", 2 * t * d * d, lambda: ops.gemm_wgrad(xn, dpre)),
        ("w.dW2", 2 * t * h * 2 * h, lambda: ops.gemm_wgrad(dz2, a1)),
        ("w.dW1", 2 * t * h * 2 * h, lambda: ops.gemm_wgrad(dz1, h0)),
        ("w.dHpre", 2 * t * d * h, lambda: ops.gemm_wgrad(xn, dh0)),
        ("colsum dz1", 0, lambda: ops.colsum_bf16(dz1)),
        ("colsum dz2", 0, lambda: ops.colsum_bf16(dz2)),
    ]
    layer_ms, layer_fl = 0.0, 0.0
    lines.append(f"== T={t} D={d} H={h} x{count}")
    for name, fl, fn in calls:
        ms = timeit(fn)
        layer_ms += ms
        layer_fl += fl
        tot[name] = tot.get(name, 0.0) + ms * count
        lines.append(f"   {name:18s} {ms:8.3f} ms  {fl / ms / 1e9 if fl else 0:8.1f} TFLOP/s")
    lines.append(f"   layer total {layer_ms:8.3f} ms  {layer_fl / layer_ms / 1e9:8.1f} TFLOP/s   (x{count} = {layer_ms * count:.2f} ms)")
    del xn, h0, z1, a1, z2, a2, dpre, dz2, dz1, dh0
    torch.cuda.empty_cache()
lines.append("== whole model, per call site (ms per step):")
for k, v in tot.items():
    lines.append(f"   {k:18s} {v:8.2f}")
lines.append(f"   total {sum(tot.values()):.2f} ms")
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/k2_train_shapes.txt", "w").write("\n".join(lines) + "\n")
print("\n".join(lines[-20:]))
