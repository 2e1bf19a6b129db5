"""At which iteration does the forward kernel's per-token Sinkhorn result stop changing bitwise?  (development aid for HVS_MHC_ADAPTIVE_ITERS)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, hvs_b200
dev = "cuda:0"
T = 1 << 15
g = torch.Generator(device=dev).manual_seed(1234)
x = torch.randn(T, 4, 512, generator=g, device=dev, dtype=torch.bfloat16)
gp = torch.Generator(device=dev).manual_seed(0)
phi = torch.randn(2048, 24, generator=gp, device=dev) * 0.02
bias = torch.zeros(24, device=dev); alpha = torch.full((3,), float(sys.argv[1]) if len(sys.argv) > 1 else 0.01, device=dev); scale = torch.ones(2048, device=dev)
tolk = {tol: torch.full((T,), -1, device=dev) for tol in (1e-6, 3e-7)}; prevf = None
prev = None; first_same = torch.full((T,), -1, device=dev); cyc2 = torch.full((T,), -1, device=dev); hist = []
for k in range(0, 25):
    _, _, co = hvs_b200.ops.mhc_stream_fwd(x, phi, bias, alpha, scale, sk_iters=k, want_y=False, want_coeffs=True)
    h = co[:, 8:].contiguous().view(torch.int32)
    if prev is not None:
        same = (h == prev).all(1)
        first_same = torch.where((first_same < 0) & same, torch.full_like(first_same, k), first_same)
    if len(hist) >= 2:
        c2 = (h == hist[-2]).all(1)
        cyc2 = torch.where((cyc2 < 0) & c2, torch.full_like(cyc2, k), cyc2)
    hf = co[:, 8:]
    if prevf is not None:
        relc = ((hf - prevf).abs() / hf.abs()).amax(1)
        for tol in tolk:
            tolk[tol] = torch.where((tolk[tol] < 0) & (relc <= tol), torch.full_like(tolk[tol], k), tolk[tol])
    prevf = hf.clone()
    hist.append(h); prev = h
fs = first_same.float()
print("P fixed (bitwise) by iteration: mean", fs[fs >= 0].mean().item(), "never", (first_same < 0).float().mean().item(),
      "quantiles", [int(torch.quantile(torch.where(fs < 0, torch.full_like(fs, 99.0), fs), q).item()) for q in (0.5, 0.9, 0.99, 0.999)])
w = torch.where(first_same < 0, torch.full_like(first_same, 99), first_same).view(-1, 8).max(1).values.float()
print("per 8-token warp: mean", w.clamp(max=25).mean().item(), "quantiles", [int(torch.quantile(w, q).item()) for q in (0.5, 0.9, 0.99)], "frac never", (w > 50).float().mean().item())
c = cyc2.float()
print("period-2 (or fixed) by iteration: never", (cyc2 < 0).float().mean().item(), "quantiles", [int(torch.quantile(torch.where(c < 0, torch.full_like(c, 99.0), c), q).item()) for q in (0.5, 0.9, 0.99, 0.999)])
w2 = torch.where(cyc2 < 0, torch.full_like(cyc2, 99), cyc2).view(-1, 8).max(1).values.float()
print("per warp period-2: quantiles", [int(torch.quantile(w2, q).item()) for q in (0.5, 0.9, 0.99)], "frac never", (w2 > 50).float().mean().item())
for tol, tk in tolk.items():
    w = torch.where(tk < 0, torch.full_like(tk, 99), tk).view(-1, 8).max(1).values.float()
    print(f"relative change <= {tol}: per token quantiles", [int(torch.quantile(torch.where(tk < 0, torch.full_like(tk, 99), tk).float(), q).item()) for q in (0.5, 0.99, 0.999)],
          "per 8-token warp quantiles", [int(torch.quantile(w, q).item()) for q in (0.5, 0.9, 0.99, 0.999)], "mean", w.clamp(max=25).mean().item())
