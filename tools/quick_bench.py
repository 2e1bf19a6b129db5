"""Quick device-side timing of the K1 kernels (development aid; bench.py is the contract)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hvs_b200
from variants import use_variant; use_variant()

T = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(T, 4, 512, generator=g, device=dev, dtype=torch.bfloat16)
dy = torch.randn(T, 4, 512, generator=g, device=dev, dtype=torch.bfloat16)
phi = torch.randn(2048, 24, generator=g, device=dev) * 0.02
bias = torch.zeros(24, device=dev); alpha = torch.full((3,), 0.01, device=dev); scale = torch.ones(2048, device=dev)
y = torch.empty_like(x)
def fwd():
    hvs_b200.ops.mhc_stream_fwd(x, phi, bias, alpha, scale, out=y)
def bwd():
    hvs_b200.ops.mhc_stream_bwd(x, dy, phi, bias, alpha, scale)
def timeit(fn, n):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts)//2], ts[0]
res = {}
med, mn = timeit(fwd, iters)
res["fwd_ms_med"] = med; res["fwd_ms_min"] = mn
res["fwd_GBs_med"] = 8192 * T / med / 1e6
# plain copy for reference
med_c, mn_c = timeit(lambda: y.copy_(x), iters)
res["copy_ms_med"] = med_c; res["copy_GBs_med"] = 8192 * T / med_c / 1e6
try:
    med, mn = timeit(bwd, iters)
    res["bwd_ms_med"] = med; res["bwd_ms_min"] = mn
    res["bwd_GBs_med"] = 12288 * T / med / 1e6
except Exception as e:
    res["bwd"] = str(e)
try:
    saved = hvs_b200.ops.new_saved(x)
    dx = torch.empty_like(x)
    ws = torch.empty(int(hvs_b200._lib.load().hvs_mhc_stream_bwd_saved_workspace(T, 4, 512)), dtype=torch.uint8, device=dev)
    def fwd_s():
        hvs_b200.ops.mhc_stream_fwd(x, phi, bias, alpha, scale, out=y, saved=saved)
    def bwd_s():
        hvs_b200.ops.mhc_stream_bwd_saved(x, dy, saved, phi, bias, alpha, scale, out=dx, workspace=ws)
    med, mn = timeit(fwd_s, iters)
    res["fwd_save_ms_med"] = med; res["fwd_save_GBs"] = (8192 + 112) * T / med / 1e6
    med, mn = timeit(bwd_s, iters)
    res["bwd_fused_ms_med"] = med; res["bwd_fused_ms_min"] = mn
    res["bwd_fused_GBs_med"] = (12288 + 112) * T / med / 1e6
    res["fwd_bwd_fused_ms"] = res["fwd_save_ms_med"] + med
    res["fwd_bwd_frac_of_6548.8"] = 20480 * T / (res["fwd_bwd_fused_ms"]) / 1e6 / 6548.8
except Exception as e:
    res["bwd_fused"] = str(e)
print(json.dumps(res))
