"""Stand-alone Sinkhorn-Knopp and constrained_matrices on the GPU vs the oracle / golden vectors."""
import numpy as np
import pytest
import torch

from oracle import mhc_ref

pytestmark = pytest.mark.gpu


def test_sinkhorn_golden(golden):
    import hvs_b200
    g = golden("sinkhorn")
    for k in "abcd":
        m = torch.from_numpy(g[f"{k}_in"]).cuda()
        iters = int(g[f"{k}_iters"])
        hist = torch.zeros(iters, device="cuda")
        out = hvs_b200.ops.sinkhorn(m, iters, history=hist).cpu()
        want = torch.from_numpy(g[f"{k}_out"])
        assert ((out - want).abs() / want.abs()).max() < 1e-5
        assert torch.allclose(hist.cpu(), torch.from_numpy(g[f"{k}_hist"]), atol=2e-6)
        assert (out >= 0).all()


@pytest.mark.parametrize("shape", [(1000, 4, 4), (7, 8, 8), (3, 16, 12), (2, 32, 32), (64, 64), (257, 257), (512, 512), (1024, 1024),
                                   (1792, 1792), (40, 1300)])
def test_sinkhorn_shapes(shape):
    import hvs_b200
    torch.manual_seed(sum(shape))
    m = torch.randn(*shape) * 0.3
    out = hvs_b200.ops.sinkhorn(m.cuda(), 20).cpu()
    want = mhc_ref.sinkhorn_knopp(m, 20)
    assert ((out - want).abs() / want.abs()).max() < 1e-5
    if shape[-1] == shape[-2]:
        assert (out.sum(-1) - 1).abs().max() < 1e-4 and (out.sum(-2) - 1).abs().max() < 1e-4


def test_constrained_matrices_golden(golden):
    import hvs_b200
    g = golden("mhc_module")
    for tag in ("d64n4", "d32n2"):
        raw = [torch.from_numpy(g[f"{tag}/p/{k}"]).cuda() for k in ("H_pre_raw", "H_post_raw", "H_res_raw")]
        hp, hq, hr = hvs_b200.ops.constrained_matrices(*raw)
        for got, name in ((hp, "H_pre"), (hq, "H_post"), (hr, "H_res")):
            want = torch.from_numpy(g[f"{tag}/{name}"])
            assert ((got.cpu() - want).abs() / want.abs()).max() < 1e-5, name


def test_sinkhorn_deterministic():
    import hvs_b200
    m = torch.randn(300, 300, generator=torch.Generator().manual_seed(0)).cuda()
    a = hvs_b200.ops.sinkhorn(m, 20)
    b = hvs_b200.ops.sinkhorn(m, 20)
    assert torch.equal(a, b)
