"""Development aid: hvs_se_gate_bf16 against the torch module sequence (pool, 1x1 conv, act, 1x1 conv, sigmoid under bf16
autocast) at the backbone's layer shapes, per call inside a CUDA graph replay.  python tools/bench_se_gate.py [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hvs_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = "cuda"
def timed(fn, n=20):
    for _ in range(3): fn()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
for hw, c in [(320 * 320, 32), (320 * 320, 64), (160 * 160, 64), (80 * 80, 128), (40 * 40, 256), (20 * 20, 512)]:
    s = int(hw ** 0.5)
    y = torch.randn(B, c, s, s, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    ca = torch.nn.Sequential(torch.nn.AdaptiveAvgPool2d(1), torch.nn.Conv2d(c, c // 4, 1), torch.nn.SiLU(), torch.nn.Conv2d(c // 4, c, 1),
                             torch.nn.Sigmoid()).to(dev).to(torch.bfloat16)
    ws = [None]
    def fused():
        with torch.no_grad():
            _, ws[0] = ops.se_gate(y, ca[1].weight, ca[1].bias, ca[3].weight, ca[3].bias, "silu", ws[0])
    def ref():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            ca(y)
    def pool():
        with torch.no_grad():
            torch.nn.functional.adaptive_avg_pool2d(y, 1)
    print(f"B={B} HW={hw:6d} C={c:4d}  bytes {y.numel() * 2 / 1e6:8.1f} MB   fused {timed(fused):8.1f} us   torch {timed(ref):8.1f} us (pool alone {timed(pool):7.1f})")
