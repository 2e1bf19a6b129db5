"""Fused (saved-statistics) backward against the recompute backward and the oracle (debug aid, GPU only)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hvs_b200
from variants import use_variant; use_variant()
from oracle import mhc_ref

T = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
dev = "cuda:0"
g = torch.Generator().manual_seed(0)
x = torch.randn(T, 4, 512, generator=g).to(torch.bfloat16)
dy = torch.randn(T, 4, 512, generator=g).to(torch.bfloat16)
phi = torch.randn(2048, 24, generator=g) * 0.02
bias = torch.randn(24, generator=g) * 0.1
alpha = torch.tensor([0.5, 0.7, 0.9])
scale = 1.0 + 0.1 * torch.randn(2048, generator=g)
xd, dyd = x.to(dev), dy.to(dev)
P = [t.to(dev) for t in (phi, bias, alpha, scale)]
saved = hvs_b200.ops.new_saved(xd)
y, _, _ = hvs_b200.ops.mhc_stream_fwd(xd, *P, saved=saved)
torch.cuda.synchronize()
print("saved finite", bool(torch.isfinite(saved).all()), saved[0, :26].tolist()[:4], "ss", saved[0, 24].item(), (x[0].float() ** 2).sum().item())
old = hvs_b200.ops.mhc_stream_bwd(xd, dyd, *P)
torch.cuda.synchronize()
new = hvs_b200.ops.mhc_stream_bwd_saved(xd, dyd, saved, *P)
torch.cuda.synchronize()
for k in old:
    a, b = old[k].float(), new[k].float()
    print(k, "fused vs recompute: max abs", (a - b).abs().max().item(), "rel norm", ((a - b).norm() / a.norm()).item(), "nan", int(torch.isnan(b).sum()))
if T <= 4096:
    rg = mhc_ref.stream_mhc_backward(x, dy, phi, bias, alpha, scale)
    for k in ("dx", "dphi", "dbias", "dalpha", "dscale"):
        a, b = rg[k].float(), new[k].cpu().float()
        print(k, "fused vs oracle: rel norm", ((a - b).norm() / a.norm()).item(), "max abs", (a - b).abs().max().item(), "ref max", a.abs().max().item())
