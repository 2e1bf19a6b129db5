import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, hvs_b200
from oracle import mhc_ref
from test_gpu_mhc_stream import make_inputs, run_bwd, run_bwd_saved
for (alpha, phistd, bstd) in ((1.0, 0.05, 0.5), (1.0, 0.02, 0.5), (0.5, 0.03, 0.3)):
    inp = make_inputs(777, seed=41, alpha=alpha, phistd=phistd, bstd=bstd)
    x, phi, bias, al, scale = inp
    dy = torch.randn(777, 4, 512, generator=torch.Generator().manual_seed(9)).to(torch.bfloat16)
    ref = mhc_ref.stream_mhc_backward(x, dy, phi, bias, al, scale)
    fwd = mhc_ref.stream_mhc_forward(x, phi, bias, al, scale)
    hres = fwd["H_res"]
    print(f"alpha {alpha} phistd {phistd} bstd {bstd}: H_res row err {float((hres.sum(-1)-1).abs().max()):.2e} col err {float((hres.sum(-2)-1).abs().max()):.2e}")
    m = fwd["H_res"] + fwd["H_post"][:, :, None] * fwd["H_pre"][:, None, :]
    mag = torch.einsum("tij,tic->tjc", m.abs(), dy.float().abs()) + ref["dx"].abs()
    for name, got in (("recompute", run_bwd(x, dy, phi, bias, al, scale)), ("fused", run_bwd_saved(x, dy, phi, bias, al, scale))):
        ul = ((got["dx"].float() - ref["dx"]).abs() / mhc_ref.bf16_ulp(mag))
        worst = int(ul.flatten().argmax()) // 2048
        rels = {k: float(((got[k].double() - ref[k].double()).norm() / ref[k].double().norm())) for k in ("dphi", "dbias", "dalpha", "dscale")}
        print(f"  {name:10s} dx max ulp {float(ul.max()):7.2f} (token {worst}, tokens>2ulp {int((ul.amax((1,2))>2).sum())})", {k: f"{v:.1e}" for k, v in rels.items()})
