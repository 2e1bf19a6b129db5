"""tcgen05 building blocks of the fused backward, checked in isolation against torch fp32 (development probe
humanoid-vision-system_b200/csrc/umma_probe.cu): the K-major G = [x ; dy] x^T MMA (M = 64 accumulator lane layout),
the MN-major dW = x^T E MMA reading the token tile in place with the bf16 hi/lo terms of E along K (stride-0 K step,
and the two-MMA form without it), the no-swizzle operand tile of E, interleaved M = 64 accumulators, 3-D TMA boxes."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", [0, 1])
def test_umma_probe_matches_torch(mode):
    import hvs_b200
    lib = hvs_b200.load_library()
    fn = lib.hvs_debug_umma_probe
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int64, ctypes.c_int, ctypes.c_void_p]
    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(3 + mode)
    t = 8
    x = torch.randn(t, 4, 512, generator=g, device=dev).to(torch.bfloat16)
    dy = torch.randn(t, 4, 512, generator=g, device=dev).to(torch.bfloat16)
    e = torch.randn(t, 24, generator=g, device=dev)
    out_gs = torch.full((128, 32), float("nan"), device=dev)
    out_dw = torch.full((128, 384), float("nan"), device=dev)
    assert fn(x.data_ptr(), dy.data_ptr(), e.data_ptr(), out_gs.data_ptr(), out_dw.data_ptr(), t, mode, None) == 0
    torch.cuda.synchronize()
    xr = x.float().permute(1, 0, 2).reshape(32, 512)           # row = 8 * stream + token
    dr = dy.float().permute(1, 0, 2).reshape(32, 512)
    want_gs = torch.cat([xr, dr]) @ xr.t()
    lanes = torch.tensor([(m % 16) + 32 * (m // 16) for m in range(64)], device=dev)     # M = 64 accumulator rows -> lanes
    assert torch.allclose(out_gs[lanes], want_gs, rtol=1e-5, atol=1e-3)
    eh = e.to(torch.bfloat16).float()
    el = (e - eh).to(torch.bfloat16).float()
    want_dw = torch.einsum("tk,tl->kl", x.float().reshape(t, 2048), eh + el)
    got = torch.empty(2048, 24, device=dev)
    for b in range(32):
        cb, j = b >> 2, b & 3
        got[j * 512 + cb * 64: j * 512 + cb * 64 + 64] = out_dw[lanes + 16 * (b & 1)][:, (b >> 1) * 24:(b >> 1) * 24 + 24]
    assert torch.allclose(got, want_dw, rtol=1e-5, atol=1e-4)
