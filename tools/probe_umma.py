"""Checks the tcgen05 building blocks of the fused backward against torch (development aid, GPU only)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hvs_b200
from hvs_b200 import _lib

lib = _lib.load()
fn = lib.hvs_debug_umma_probe
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int64, ctypes.c_int, ctypes.c_void_p]
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(3)
T = 8
x = torch.randn(T, 4, 512, generator=g, device=dev).to(torch.bfloat16)
dy = torch.randn(T, 4, 512, generator=g, device=dev).to(torch.bfloat16)
E = torch.randn(T, 24, generator=g, device=dev)
xf, dyf = x.float(), dy.float()
# expected G/S: rows r = j*8+tok of [x; dy] against x rows
xr = xf.permute(1, 0, 2).reshape(32, 512)       # row = j*8 + tok
dr = dyf.permute(1, 0, 2).reshape(32, 512)
want_gs = torch.cat([xr, dr]) @ xr.t()          # [64, 32]
Eh = E.to(torch.bfloat16).float()
El = (E - Eh).to(torch.bfloat16).float()
want_dw = torch.einsum("tk,tl->kl", xf.reshape(T, 2048), Eh + El)   # [2048, 24]
for mode in (0, 1):
    out_gs = torch.full((128, 32), float("nan"), device=dev)
    out_dw = torch.full((128, 384), float("nan"), device=dev)
    rc = fn(x.data_ptr(), dy.data_ptr(), E.data_ptr(), out_gs.data_ptr(), out_dw.data_ptr(), T, mode, None)
    torch.cuda.synchronize()
    print("mode", mode, "rc", rc)
    lanes = torch.tensor([(m % 16) + 32 * (m // 16) for m in range(64)], device=dev)
    got_gs = out_gs[lanes]
    err = (got_gs - want_gs).abs().max().item()
    print("  GS max abs err", err, "ref max", want_gs.abs().max().item())
    got_dw = torch.empty(2048, 24, device=dev)
    for b in range(32):
        cb, j = b >> 2, b & 3
        rows = lanes + 16 * (b & 1)
        got_dw[j * 512 + cb * 64: j * 512 + cb * 64 + 64] = out_dw[rows][:, (b >> 1) * 24:(b >> 1) * 24 + 24]
    err = (got_dw - want_dw).abs().max().item()
    print("  dW max abs err", err, "ref max", want_dw.abs().max().item(), "nan", int(torch.isnan(got_dw).sum()))
    if err > 1e-2:
        # diagnostics: which blocks are right
        for b in range(4):
            cb, j = b >> 2, b & 3
            sl = slice(j * 512 + cb * 64, j * 512 + cb * 64 + 64)
            print("   block", b, "err", (got_dw[sl] - want_dw[sl]).abs().max().item())
