"""Development aid: end-to-end (pinned host buffers) time of stream_mhc_fwd_bwd_host for several chunk sizes, next to
the raw PCIe copy rates of the box (H2D alone, D2H alone, both at once)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, hvs_b200
T = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
dev = torch.device("cuda:0")
layer = hvs_b200.StreamMHC(device=dev)
xh = torch.randn(T, 4, 512).to(torch.bfloat16).pin_memory(); dyh = torch.randn(T, 4, 512).to(torch.bfloat16).pin_memory()
yh = torch.empty_like(xh).pin_memory(); dxh = torch.empty_like(xh).pin_memory()
gb = xh.numel() * 2 / 1e9
d0 = torch.empty_like(xh, device=dev); d1 = torch.empty_like(xh, device=dev)
def timed(fn, n=2):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
def h2d():
    with torch.cuda.stream(s1): d0.copy_(xh, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): yh.copy_(d1, non_blocking=True)
def both(): h2d(); d2h()
print(f"H2D alone {gb / timed(h2d):.1f} GB/s, D2H alone {gb / timed(d2h):.1f} GB/s, both at once {gb / timed(both):.1f} GB/s each")
del d0, d1
torch.cuda.empty_cache()
for chunk in (8192, 16384, 32768, 65536, 131072):
    dt = timed(lambda: hvs_b200.stream_mhc_fwd_bwd_host(xh, dyh, layer, yh, dxh, chunk_tokens=chunk))
    print(f"chunk {chunk:7d}: {dt * 1e3:7.1f} ms  {T / dt / 1e6:5.2f} M tokens/s  ({2 * gb / dt:.1f} GB/s each way)")
