"""Development aid: forward / fused-backward kernel time against the number of Sinkhorn iterations (how much of
each kernel is the serial coefficient chain)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, hvs_b200
T = 1 << 20
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(T, 4, 512, generator=g, device=dev, dtype=torch.bfloat16)
dy = torch.randn(T, 4, 512, generator=g, device=dev, dtype=torch.bfloat16)
phi = torch.randn(2048, 24, generator=g, device=dev) * 0.02
bias = torch.zeros(24, device=dev); alpha = torch.full((3,), 0.01, device=dev); scale = torch.ones(2048, device=dev)
y = torch.empty_like(x); dx = torch.empty_like(x); saved = hvs_b200.ops.new_saved(x)
ws = torch.empty(int(hvs_b200._lib.load().hvs_mhc_stream_bwd_saved_workspace(T, 4, 512)), dtype=torch.uint8, device=dev)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for it in (0, 1, 2, 3, 0, 1, 2, 5, 10, 20, 24):
    f = t(lambda: hvs_b200.ops.mhc_stream_fwd(x, phi, bias, alpha, scale, it, 1e-8, 1e-8, out=y, saved=saved))
    b = t(lambda: hvs_b200.ops.mhc_stream_bwd_saved(x, dy, saved, phi, bias, alpha, scale, it, 1e-8, 1e-8, out=dx, workspace=ws))
    print(f"iters {it:2d}: fwd {f:.3f} ms  bwd {b:.3f} ms")
for it in (3, 20):
    f = t(lambda: hvs_b200.ops.mhc_stream_fwd(x, phi, bias, alpha, scale, it, 1e-8, 1e-8, out=y, saved=saved, adaptive=True))
    b = t(lambda: hvs_b200.ops.mhc_stream_bwd_saved(x, dy, saved, phi, bias, alpha, scale, it, 1e-8, 1e-8, out=dx, workspace=ws, adaptive=True))
    print(f"adaptive, limit {it:2d}: fwd {f:.3f} ms  bwd {b:.3f} ms")
