import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, hvs_b200
torch.manual_seed(0)
layer = hvs_b200.StreamMHC(alpha=0.2, device="cuda")
x = torch.randn(3, 50, 4, 512, device="cuda").to(torch.bfloat16).requires_grad_(True)
y = layer(x)
dy = torch.randn_like(y)
xf = x.detach().reshape(-1, 4, 512)
print("ptrs", xf.data_ptr() % 256, dy.data_ptr() % 256, dy.is_contiguous(), dy.shape, dy.stride())
for name, t in (("direct", dy.reshape(-1, 4, 512)),):
    try:
        g = hvs_b200.ops.mhc_stream_bwd(xf, t, layer.phi.detach(), layer.bias.detach(), layer.alpha.detach(), layer.rms_scale.detach())
        print(name, "ok")
    except Exception as e:
        print(name, "FAIL", e)
try:
    y.backward(dy)
    print("autograd ok")
except Exception as e:
    print("autograd FAIL", e)
class F(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a):
        return a.clone()
    @staticmethod
    def backward(ctx, g):
        print("in backward: g ptr%256", g.data_ptr() % 256, g.is_contiguous(), g.shape, g.stride(), g.dtype)
        return g
z = F.apply(x)
z.backward(dy)
