"""Per-layer-shape timing of the fused K2 path at batch-64 token counts: ms, TFLOP/s per (T, D, H) and
the per-kernel split (LN | 4 GEMMs).  python tools/bench_k2_shapes.py [batch]"""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hvs_b200
from hvs_b200 import _lib, ops
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
SHAPES = [(102400, 32, 128, 2), (102400, 64, 256, 1), (25600, 32, 128, 1), (25600, 64, 256, 3), (6400, 64, 256, 2), (6400, 128, 512, 6),
          (6400, 256, 512, 2), (1600, 128, 512, 3), (1600, 256, 1024, 8), (1600, 256, 512, 1), (1600, 512, 1024, 1), (400, 256, 1024, 1),
          (400, 512, 2048, 4), (400, 256, 512, 2), (400, 512, 1024, 1), (400, 1024, 2048, 1), (401, 256, 512, 36)]
dev = "cuda:0"
tot_ms = tot_fl = 0.0
print(f"{'T':>9s} {'D':>5s} {'H':>5s} {'x':>3s} {'ms/layer':>9s} {'TFLOP/s':>8s} {'ms total':>9s} | LN, G1, G2(gelu), G3(gelu), G4+5(LN)  [ms]")
for t, d, h, cnt in SHAPES:
    T = t * batch
    mod = hvs_b200.ManifoldHyperConnection(d, hidden_dim=h).to(dev).eval()
    mod.output_dtype = torch.bfloat16
    x = torch.randn(T, d, device=dev, dtype=torch.bfloat16)
    with torch.no_grad():
        for _ in range(2): mod(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): mod(x)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        # per-kernel split
        st = mod._fresh_state(); w1, w2 = mod._mlp_bf16()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        evs[0].record()
        xn, _ = ops.layernorm_fwd(x, mod.norm_pre.weight.detach(), mod.norm_pre.bias.detach(), 1e-5, torch.bfloat16, d); evs[1].record()
        z = ops.gemm_bf16(xn, st.h_pre_t); evs[2].record()
        z = ops.gemm_bf16(z, w1, bias=mod.mlp[0].bias.detach(), epilogue=_lib.HVS_GEMM_EPI_BIAS_GELU); evs[3].record()
        z = ops.gemm_bf16(z, w2, bias=mod.mlp[3].bias.detach(), epilogue=_lib.HVS_GEMM_EPI_BIAS_GELU); evs[4].record()
        if d <= 512:
            ops.gemm_bf16(z, st.h_post_t, x, st.h_res_t, ln_weight=mod.norm_post.weight.detach(), ln_bias=mod.norm_post.bias.detach(), epilogue=_lib.HVS_GEMM_EPI_LAYERNORM)
        else:
            ops.gemm_bf16(z, st.h_post_t, x, st.h_res_t, out_dtype=torch.float32)
        evs[5].record(); torch.cuda.synchronize()
        parts = [evs[i].elapsed_time(evs[i + 1]) for i in range(5)]
    fl = 2.0 * (2 * d * h + 4 * h * h + d * d) * T
    tot_ms += ms * cnt; tot_fl += fl * cnt
    print(f"{T:9d} {d:5d} {h:5d} {cnt:3d} {ms:9.3f} {fl / ms / 1e9:8.1f} {ms * cnt:9.2f} | " + ", ".join(f"{p:.3f}" for p in parts))
    del mod, x
    torch.cuda.empty_cache()
print(f"all 75 layers: {tot_ms:.2f} ms, {tot_fl / 1e12:.2f} TFLOP, {tot_fl / tot_ms / 1e9:.1f} TFLOP/s")
