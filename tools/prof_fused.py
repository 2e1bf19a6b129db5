import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, hvs_b200
T = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(T, 4, 512, generator=g, device=dev, dtype=torch.bfloat16)
dy = torch.randn(T, 4, 512, generator=g, device=dev, dtype=torch.bfloat16)
phi = torch.randn(2048, 24, generator=g, device=dev) * 0.02
bias = torch.zeros(24, device=dev); alpha = torch.full((3,), 0.01, device=dev); scale = torch.ones(2048, device=dev)
saved = hvs_b200.ops.new_saved(x)
y = torch.empty_like(x)
for _ in range(2):
    hvs_b200.ops.mhc_stream_fwd(x, phi, bias, alpha, scale, out=y, saved=saved)
    hvs_b200.ops.mhc_stream_bwd_saved(x, dy, saved, phi, bias, alpha, scale)
torch.cuda.synchronize()
