"""Development aid: A/B of library variants on one box.  HVS_VARIANT=<name> picks hvs_b200/build/variants/libhvs_b200_<name>.so.
Times 20 back-to-back training steps (forward saving statistics + fused backward) and each kernel alone."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, hvs_b200
from variants import use_variant; use_variant()
T = 1 << 20
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(T, 4, 512, generator=g, device=dev, dtype=torch.bfloat16)
dy = torch.randn(T, 4, 512, generator=g, device=dev, dtype=torch.bfloat16)
phi = torch.randn(2048, 24, generator=g, device=dev) * 0.02
bias = torch.zeros(24, device=dev); alpha = torch.full((3,), 0.01, device=dev); scale = torch.ones(2048, device=dev)
y = torch.empty_like(x); dx = torch.empty_like(x); saved = hvs_b200.ops.new_saved(x)
ws = torch.empty(int(hvs_b200._lib.load().hvs_mhc_stream_bwd_saved_workspace(T, 4, 512)), dtype=torch.uint8, device=dev)
def fwd(): hvs_b200.ops.mhc_stream_fwd(x, phi, bias, alpha, scale, out=y, saved=saved)
def bwd(): hvs_b200.ops.mhc_stream_bwd_saved(x, dy, saved, phi, bias, alpha, scale, out=dx, workspace=ws)
def step(): fwd(); bwd()
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
print(f"{os.environ.get('HVS_VARIANT', 'default'):10s} step {t(step):.3f} ms   fwd {t(fwd):.3f}   bwd {t(bwd):.3f}")
