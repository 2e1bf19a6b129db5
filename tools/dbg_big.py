import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, hvs_b200
from variants import use_variant; use_variant()
t = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
g = torch.Generator(device="cuda:0").manual_seed(11)
x = torch.randn(t, 4, 512, generator=g, device="cuda:0", dtype=torch.bfloat16)
dy = torch.randn(t, 4, 512, generator=g, device="cuda:0", dtype=torch.bfloat16)
gg = torch.Generator().manual_seed(1)
phi = (torch.randn(2048, 24, generator=gg) * 0.02).cuda(); bias = torch.zeros(24).cuda(); al = torch.full((3,), 0.3).cuda(); scale = torch.ones(2048).cuda()
P = [phi, bias, al, scale]
saved = hvs_b200.ops.new_saved(x)
hvs_b200.ops.mhc_stream_fwd(x, *P, saved=saved)
a = hvs_b200.ops.mhc_stream_bwd_saved(x, dy, saved, *P)
b = hvs_b200.ops.mhc_stream_bwd_saved(x, dy, saved, *P)
old = hvs_b200.ops.mhc_stream_bwd(x, dy, *P)
c = hvs_b200.ops.mhc_stream_bwd_saved(x, (dy.float() * 2).to(torch.bfloat16), saved, *P)
torch.cuda.synchronize()
for k in ("dx", "dphi", "dbias", "dalpha", "dscale"):
    A, B, O = a[k].float(), b[k].float(), old[k].float()
    print(k, "rerun equal", bool(torch.equal(A, B)), "max |a-b|", float((A - B).abs().max()),
          "vs two-kernel rel", float((A - O).norm() / O.norm()), "max abs", float((A - O).abs().max()), "ref max", float(O.abs().max()))
    if k != "dx":
        C = c[k].float()
        print("   linear: max |c-2a|", float((C - 2 * A).abs().max()), "n bad", int(((C - 2 * A).abs() > 1e-6 * A.abs().max()).sum()))
d = (a["dx"].float() - old["dx"].float()).abs().amax((1, 2))
bad = torch.nonzero(d > 0.05 * old["dx"].float().abs().max()).flatten()
print("tokens with large dx diff:", bad.numel(), bad[:20].tolist())
