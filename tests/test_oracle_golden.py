"""The CPU oracle reproduces the golden vectors written from the reference (oracle/make_golden.py)."""
import numpy as np
import torch

from oracle import detect_ref, mhc_ref


def bits_to_bf16(a):
    return torch.from_numpy(a.astype(np.int16)).view(torch.bfloat16)


def test_sinkhorn_golden(golden):
    g = golden("sinkhorn")
    for k in "abcd":
        m = torch.from_numpy(g[f"{k}_in"])
        out, hist = mhc_ref.sinkhorn_knopp(m, int(g[f"{k}_iters"]), return_history=True)
        assert torch.allclose(out, torch.from_numpy(g[f"{k}_out"]), rtol=1e-6, atol=1e-8)
        assert torch.allclose(hist, torch.from_numpy(g[f"{k}_hist"]), atol=1e-6)
        # reference test_models.py:44-53 properties
        assert (out >= 0).all()
        if k != "d":
            assert torch.allclose(out.sum(-1), torch.ones_like(out.sum(-1)), rtol=1e-4)
            assert torch.allclose(out.sum(-2), torch.ones_like(out.sum(-2)), rtol=1e-4)


def test_rmsnorm_golden(golden):
    g = golden("rmsnorm")
    out = mhc_ref.rms_norm(torch.from_numpy(g["x"]), torch.from_numpy(g["scale"]))
    assert torch.allclose(out, torch.from_numpy(g["out"]), rtol=1e-6, atol=1e-7)


def test_module_golden(golden):
    g = golden("mhc_module")
    for tag in ("d64n4", "d32n2"):
        p = {k.split("/p/")[1]: torch.from_numpy(g[k]) for k in g.files if k.startswith(tag + "/p/")}
        y = mhc_ref.mhc_module_forward(torch.from_numpy(g[f"{tag}/x"]), p)
        assert torch.allclose(y, torch.from_numpy(g[f"{tag}/y"]), rtol=1e-5, atol=1e-5)
        hp, hq, hr = mhc_ref.constrained_matrices(p["H_pre_raw"], p["H_post_raw"], p["H_res_raw"])
        assert torch.allclose(hr, torch.from_numpy(g[f"{tag}/H_res"]), rtol=1e-6)
        # reference test_models.py:145-159: gate ranges, H_res sums
        assert (hp >= 0).all() and (hp <= 1).all() and (hq >= 0).all() and (hq <= 2).all()
        assert torch.allclose(hr.sum(0), torch.ones(hr.shape[0]), rtol=1e-3)
        assert torch.allclose(hr.sum(1), torch.ones(hr.shape[0]), rtol=1e-3)


def test_stream_golden(golden):
    g = golden("stream_mhc")
    for tag in ("n4c512", "n4c512_hot"):
        x = bits_to_bf16(g[f"{tag}/x_bits"])
        out = mhc_ref.stream_mhc_forward(x, torch.from_numpy(g[f"{tag}/phi"]), torch.from_numpy(g[f"{tag}/bias"]),
                                         torch.from_numpy(g[f"{tag}/alpha"]), torch.from_numpy(g[f"{tag}/scale"]))
        for name in ("H_pre", "H_post", "H_res"):
            ref = torch.from_numpy(g[f"{tag}/{name}"])
            assert ((out[name] - ref).abs() / ref.abs()).max() < 1e-5
        y_ref = torch.from_numpy(g[f"{tag}/y"])
        mag = mhc_ref.mixing_condition_magnitude(x, out["H_pre"], out["H_post"], out["H_res"])
        assert ((out["y"].float() - y_ref).abs() <= 2 * mhc_ref.bf16_ulp(mag)).all()


def test_stream_backward_matches_finite_difference():
    torch.manual_seed(0)
    t, n, c = 3, 4, 512
    x = torch.randn(t, n, c).to(torch.bfloat16)
    phi = (torch.randn(n * c, 24) * 0.02).to(torch.bfloat16).float()   # bf16-valued: no rounding kink
    bias = torch.randn(24) * 0.1
    alpha = torch.tensor([0.3, 0.2, 0.5])
    scale = torch.ones(n * c)
    dy = torch.randn(t, n, c)
    g = mhc_ref.stream_mhc_backward(x, dy, phi, bias, alpha, scale)
    def loss(b_, a_):
        o = mhc_ref.stream_mhc_forward(x, phi, b_, a_, scale, round_output=False, dtype=torch.float64)
        return (o["y"].double() * dy.double()).sum()
    # double-precision central differences on bias and alpha
    for i in (0, 5, 13, 23):
        e = torch.zeros(24, dtype=torch.float64); e[i] = 1e-4
        fd = (loss(bias.double() + e, alpha.double()) - loss(bias.double() - e, alpha.double())) / 2e-4
        assert abs(fd.item() - g["dbias"][i].item()) <= 2e-3 * max(1.0, abs(fd.item()))
    for i in range(3):
        e = torch.zeros(3, dtype=torch.float64); e[i] = 1e-4
        fd = (loss(bias.double(), alpha.double() + e) - loss(bias.double(), alpha.double() - e)) / 2e-4
        assert abs(fd.item() - g["dalpha"][i].item()) <= 2e-3 * max(1.0, abs(fd.item()))


def test_decode_golden(golden):
    g = golden("decode")
    pred = torch.from_numpy(g["pred"])
    for s in range(3):
        d = detect_ref.yolo_decode(pred, torch.from_numpy(g[f"s{s}/anchor_wh"]))
        assert torch.allclose(d["boxes"], torch.from_numpy(g[f"s{s}/boxes"]), rtol=1e-6, atol=1e-7)
        assert torch.equal(d["class_indices"], torch.from_numpy(g[f"s{s}/class_indices"]))


def test_nms_golden(golden):
    g = golden("nms")
    assert detect_ref.nms_class_aware(g["ka/boxes"], g["ka/scores"], g["ka/classes"], 0.5).tolist() == [0, 2]
    for tag in ("ag300", "ag1000", "ag64cap5", "ag1"):
        keep = detect_ref.nms_agnostic(g[f"{tag}/boxes"], g[f"{tag}/scores"], float(g[f"{tag}/thr"]), int(g[f"{tag}/cap"]))
        assert keep.tolist() == g[f"{tag}/keep"].tolist()
    for tag in ("ca400", "ca600", "ca200cap1000"):
        keep = detect_ref.nms_class_aware(g[f"{tag}/boxes"], g[f"{tag}/scores"], g[f"{tag}/classes"],
                                          float(g[f"{tag}/thr"]), int(g[f"{tag}/cap"]))
        assert keep.tolist() == g[f"{tag}/keep"].tolist()


def test_post_process_golden(golden):
    g = golden("nms")
    decoded = [detect_ref.yolo_decode(torch.from_numpy(g[f"pp/pred{s}"]), detect_ref.anchors_wh(s)) for s in range(3)]
    out = detect_ref.post_process(decoded, float(g["pp/conf"]), float(g["pp/thr"]), int(g["pp/cap"]))
    for b in range(2):
        assert np.array_equal(out[b]["boxes"], g[f"pp/{b}/boxes"])
        assert np.array_equal(out[b]["scores"], g[f"pp/{b}/scores"])
        assert np.array_equal(out[b]["labels"], g[f"pp/{b}/labels"])


def test_nms_edge_cases():
    assert detect_ref.nms_agnostic(np.zeros((0, 4)), np.zeros(0)).tolist() == []
    assert detect_ref.nms_class_aware(np.zeros((0, 4)), np.zeros(0), np.zeros(0)).tolist() == []
    # identical boxes: only the best survives (agnostic); different classes survive class-aware
    b = np.tile(np.array([[0.1, 0.1, 0.5, 0.5]], np.float32), (4, 1))
    s = np.array([0.2, 0.9, 0.5, 0.7], np.float32)
    assert detect_ref.nms_agnostic(b, s).tolist() == [1]
    assert detect_ref.nms_class_aware(b, s, np.array([0, 0, 1, 1]), boxes_are_corners=True).tolist() == [1, 3]
