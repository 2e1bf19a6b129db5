import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, hvs_b200
from variants import use_variant; use_variant()
T = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(T, 4, 512, generator=g, device=dev, dtype=torch.bfloat16)
dy = torch.randn(T, 4, 512, generator=g, device=dev, dtype=torch.bfloat16)
phi = torch.randn(2048, 24, generator=g, device=dev) * 0.02
bias = torch.zeros(24, device=dev); alpha = torch.full((3,), 0.01, device=dev); scale = torch.ones(2048, device=dev)
saved = hvs_b200.ops.new_saved(x); y = torch.empty_like(x)
hvs_b200.ops.mhc_stream_fwd(x, phi, bias, alpha, scale, out=y, saved=saved)
lib = hvs_b200._lib.load()
lib.hvs_debug_fused_timing.argtypes = [ctypes.c_void_p, ctypes.c_int]
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
buf = torch.zeros(148 * 4 * 8 + 64 * 12, dtype=torch.int64, device=dev)
hvs_b200.ops.mhc_stream_bwd_saved(x, dy, saved, phi, bias, alpha, scale)
lib.hvs_debug_fused_timing(buf.data_ptr(), mode)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); hvs_b200.ops.mhc_stream_bwd_saved(x, dy, saved, phi, bias, alpha, scale); b.record()
torch.cuda.synchronize()
lib.hvs_debug_fused_timing(None, 0)
ms = a.elapsed_time(b)
tt = buf[:148 * 4 * 8].view(148, 4, 8).double()
t = tt[:, :3]
tiles_per_warp = (T / 8) / 148 / 3
names = ["wait_full", "prologue", "fwd_loop", "wait_G", "mid(M,gates)", "bwd_loop", "epilogue", "-"]
print(f"kernel {ms:.3f} ms, tiles per coefficient warp {tiles_per_warp:.1f}")
for i, n in enumerate(names[:7]):
    print(f"  {n:14s} {t[:, :, i].mean().item() / tiles_per_warp:9.0f} cycles per tile")
print("  total per tile", t[:, :, :7].sum(-1).mean().item() / tiles_per_warp)

fn = ["wait_full", "GS issue", "wait_ed", "dW issue", "wait_dxr", "store+drain", "wait_dw", "load issue"]
tiles = (T / 8) / 148
print("front thread, cycles per tile:")
for i, n in enumerate(fn):
    print(f"  {n:14s} {tt[:, 3, i].mean().item() / tiles:9.0f}")
print("  total", tt[:, 3].sum(-1).mean().item() / tiles)

tr = buf[148 * 4 * 8:].view(64, 12).cpu().numpy()
base = tr[20, 0]
ev = ["load", "Gstart", "Gend", "fwd0", "fwd1", "Ggot", "cdone", "P3s", "P3e", "ed", "store", "freed"]
print("trace CTA 0, tiles 20..31, kcycles relative to load(20):")
print("tile " + " ".join(f"{e:>7s}" for e in ev))
for k in range(20, 32):
    print(f"{k:4d} " + " ".join(f"{(tr[k, e] - base) / 1000:7.1f}" for e in range(12)))
import numpy as np
sl = tr[8:60].astype(np.float64)
def dm(a, b): return float(np.mean(sl[:, b] - sl[:, a]))
print(f"SUMMARY {os.environ.get('HVS_VARIANT','')} kernel {ms:.3f} ms period {float(np.mean(np.diff(sl[:, 8]))):.0f} | load->Gstart {dm(0,1):.0f} Gstart->Gend {dm(1,2):.0f} Gend->Ggot {dm(2,5):.0f} "
      f"Ggot->cdone {dm(5,6):.0f} cdone->P3s {dm(6,7):.0f} P3 {dm(7,8):.0f} P3e->store {dm(8,10):.0f} store->freed {dm(10,11):.0f} fwd {dm(3,4):.0f}")
