"""Small driver for ncu captures of the round-2 kernels: K2 GEMMs (D=512, n=4, T=2^17), the single-kernel chain
(D=64, T=2^19), decode + NMS at batch 16.  python tools/ncu_targets.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hvs_b200
from hvs_b200 import harness
dev = torch.device("cuda", 0)
torch.manual_seed(0)
with torch.no_grad():
    m = hvs_b200.ManifoldHyperConnection(512, expansion_rate=4).to(dev).eval()
    m.output_dtype = torch.bfloat16
    x = torch.randn(1 << 17, 512, device=dev, dtype=torch.bfloat16)
    for _ in range(2):
        m(x)
    c = hvs_b200.ManifoldHyperConnection(64, expansion_rate=4).to(dev).eval()
    c.output_dtype = torch.bfloat16
    xc = torch.randn(1 << 19, 64, device=dev, dtype=torch.bfloat16)
    for _ in range(2):
        c(xc)
torch.cuda.synchronize()
print(harness.detect_tail(dev, 16, 0.0, steps=2, warmup=1))
