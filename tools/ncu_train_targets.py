"""Driver for the ncu captures of the round-2 TRAINING kernels: one forward + backward of ManifoldHyperConnection(512, 4)
at T = 2^15 and of ManifoldHyperConnection(64, 4) at T = 2^17 on _K2TokenPathFn (4 forward GEMMs, 5 data-gradient GEMMs with
the GELU' x dropout epilogue, 5 weight-gradient GEMMs with both operands MN-major + split-K, column sums, LayerNorm backward),
and one general-shape K1 backward (n = 2, C = 256, T = 2^16).  python tools/ncu_train_targets.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hvs_b200
dev = torch.device("cuda", 0)
torch.manual_seed(0)
for d, t in ((512, 1 << 15), (64, 1 << 17)):
    m = hvs_b200.ManifoldHyperConnection(d, expansion_rate=4).to(dev).train()
    m.output_dtype = torch.bfloat16
    x = torch.randn(t, d, device=dev, dtype=torch.bfloat16, requires_grad=True)
    y = m(x)
    y.backward(torch.randn_like(y))
    torch.cuda.synchronize()
layer = hvs_b200.StreamMHC(n_streams=2, channels=256, device=dev)
xs = torch.randn(1 << 16, 2, 256, device=dev).to(torch.bfloat16).requires_grad_(True)
for _ in range(1):
    layer(xs).backward(torch.randn(1 << 16, 2, 256, device=dev).to(torch.bfloat16))
torch.cuda.synchronize()
print("ok", hvs_b200._lib.launch_count())
