"""Host-side checks of the drop-in claim that need no GPU: state_dict compatibility of the host model with the
reference's HybridVisionSystem (fixture generated from the reference itself), YOLOLoss against the reference's values,
the dense-target rule, and -- where /root/reference exists (the build container) -- (a) the reference's own
HybridVisionSystem constructed with hvs_b200's classes patched in, (b) hvs_b200's host model composed with the
reference's leaf modules reproducing the reference forward on CPU."""
import importlib
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import reference_repaired

needs_reference = pytest.mark.skipif(not reference_repaired.available(), reason="needs /root/reference (build container only)")


def test_host_model_state_dict_matches_reference_table():
    import hvs_b200
    table = json.load(open(os.path.join(GOLDEN, "hybrid_vision_keys.json")))
    model = hvs_b200.HybridVisionSystem({"num_classes": 80, "image_size": 640})
    mine = {k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in model.state_dict().items()}
    assert mine == table                                  # 1574 entries, names, shapes and dtypes: reference checkpoints load
    assert sum(1 for m in model.modules() if isinstance(m, hvs_b200.ManifoldHyperConnection)) == 76
    assert abs(sum(p.numel() for p in model.parameters()) - 353.8e6) < 0.1e6
    for bad in ({"use_rag": True}, {"has_depth": True}, {"use_fpn": False}):
        with pytest.raises(NotImplementedError):
            hvs_b200.HybridVisionSystem(bad)


def test_yolo_loss_matches_reference_values(golden):
    import hvs_b200
    g = golden("yolo_loss")
    preds = {f"scale_{s}": torch.from_numpy(g[f"pred{s}"]).requires_grad_(True) for s in range(3)}
    tgts = [torch.from_numpy(g[f"tgt{s}"]) for s in range(3)]
    out = hvs_b200.YOLOLoss(num_classes=80)(preds, tgts)
    for k in ("coord_loss", "obj_loss", "noobj_loss", "cls_loss", "total_loss"):
        assert abs(float(out[k]) - float(g[k])) <= 2e-5 * abs(float(g[k])), k
    out["total_loss"].backward()
    assert preds["scale_1"].grad.abs().max() == 0           # the scale without objects contributes nothing (:411-413)
    assert preds["scale_0"].grad.abs().max() > 0
    # no objects anywhere: total is exactly zero, still differentiable
    z = hvs_b200.YOLOLoss()(preds, [torch.zeros_like(t) for t in tgts])
    assert float(z["total_loss"]) == 0.0


def test_dense_targets_rule():
    import hvs_b200
    boxes = [torch.tensor([[0.5, 0.25, 0.025, 0.03], [0.99, 0.99, 0.8, 0.7]]), torch.zeros(0, 4)]
    labels = [torch.tensor([3, 79]), torch.zeros(0, dtype=torch.long)]
    t = hvs_b200.dense_targets_from_boxes(boxes, labels, [(80, 80), (40, 40), (20, 20)])
    assert [tuple(x.shape) for x in t] == [(2, 3, 80, 80, 85), (2, 3, 40, 40, 85), (2, 3, 20, 20, 85)]
    for s, g in enumerate((80, 40, 20)):
        assert t[s][0, ..., 4].sum() == 2 and t[s][1].abs().sum() == 0
        a = int(t[s][0, :, int(0.25 * g), int(0.5 * g), 4].argmax())
        assert t[s][0, a, int(0.25 * g), int(0.5 * g), 5 + 3] == 1
        assert t[s][0, :, g - 1, g - 1, 4].sum() == 1        # clamped to the last cell
    assert int(t[0][0, :, 20, 40, 4].argmax()) == 0 and int(t[2][0, :, 19, 19, 4].argmax()) == 2   # small box -> small anchor


@needs_reference
def test_reference_system_constructs_with_hvs_b200_classes_patched_in():
    """INTEGRATION.md section 1 as a test: the reference's HybridVisionSystem, its leaf classes replaced by hvs_b200's
    in the namespaces the reference modules import them into, constructs and exposes the same state_dict."""
    import hvs_b200
    reference_repaired.load_full_model()                   # imports + repairs
    mods = {n: importlib.import_module(f"src.models.{n}") for n in ("vision_backbone", "vit_encoder_decoder", "feature_fusion",
                                                                     "yolo_head", "hybrid_vision", "manifold_layers")}
    saved = []
    try:
        for m in mods.values():
            for name, repl in (("ManifoldHyperConnection", hvs_b200.ManifoldHyperConnection), ("RMSNorm", hvs_b200.RMSNorm),
                               ("YOLODetectionHead", hvs_b200.YOLODetectionHead)):
                if hasattr(m, name):
                    saved.append((m, name, getattr(m, name)))
                    setattr(m, name, repl)
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            model = mods["hybrid_vision"].HybridVisionSystem({"num_classes": 80, "image_size": 640})
    finally:
        for m, name, old in saved:
            setattr(m, name, old)
    n_mine = sum(1 for m in model.modules() if isinstance(m, hvs_b200.ManifoldHyperConnection))
    assert n_mine == 76 and isinstance(model.detection_head, hvs_b200.YOLODetectionHead)
    table = json.load(open(os.path.join(GOLDEN, "hybrid_vision_keys.json")))
    assert {k: list(v.shape) for k, v in model.state_dict().items()} == {k: v[0] for k, v in table.items()}


@needs_reference
def test_host_model_composition_reproduces_reference_forward_on_cpu(golden):
    """hvs_b200.hybrid_vision with the REFERENCE's mHC / RMSNorm / detection head as leaves (they run on CPU), filled
    with the fixture's name-seeded parameters, must reproduce the reference model's outputs: this pins the host
    composition (backbone / ViT / FPN wiring, repairs R3 / R4 / R8) independently of the CUDA kernels."""
    import hvs_b200
    from hvs_b200 import hybrid_vision as hv
    ref = reference_repaired.load()
    g = golden("hybrid_vision")
    saved = (hv.ManifoldHyperConnection, hv.RMSNorm, hv.YOLODetectionHead, hv.refresh_static_coefficients)
    try:
        hv.ManifoldHyperConnection, hv.RMSNorm = ref.ManifoldHyperConnection, ref.RMSNorm
        hv.YOLODetectionHead = ref.YOLODetectionHead
        hv.refresh_static_coefficients = lambda model: 0
        model = hv.HybridVisionSystem({"num_classes": 80, "image_size": 640}).eval()
    finally:
        hv.ManifoldHyperConnection, hv.RMSNorm, hv.YOLODetectionHead, hv.refresh_static_coefficients = saved
    reference_repaired.fill_by_name(model, 0)
    hv_mhc = ref.ManifoldHyperConnection
    # the functions test isinstance(..., ManifoldHyperConnection) at call time: keep the reference class visible while running
    hv.ManifoldHyperConnection = hv_mhc
    try:
        with torch.no_grad():
            out = model(torch.from_numpy(g["x"]))
    finally:
        hv.ManifoldHyperConnection = saved[0]
    # fp32 on both sides; the fixture was written single-threaded, so only the reduction order differs (x 76 layers)
    for s in range(3):
        assert torch.allclose(out["predictions"][f"scale_{s}"], torch.from_numpy(g[f"pred{s}"]), rtol=2e-3, atol=2e-3), s
    assert torch.allclose(out["final_features"], torch.from_numpy(g["final_features"]), rtol=2e-3, atol=2e-3)
    assert torch.allclose(out["vit_features"].mean((2, 3)), torch.from_numpy(g["vit_features_mean"]), rtol=2e-3, atol=2e-3)


def test_inference_preparation_helpers_restructure_the_host_model_as_documented():
    """harness.fold_batchnorm_for_inference / cast_weights_for_bf16_inference on the CPU (module surgery only, no forward):
    every conv -> BatchNorm pair is folded, the conv -> BatchNorm -> activation stacks of the FPN and the prediction heads
    hand their bias to a FoldedBiasAct, the mHC modules keep fp32 master parameters, everything else that autocast would
    cast per call is stored in bf16."""
    import torch
    import hvs_b200
    from hvs_b200 import harness
    m = harness.build_model(torch.device("cpu"), seed=0).eval()
    n_bn = sum(isinstance(x, torch.nn.BatchNorm2d) for x in m.modules())
    assert harness.fold_batchnorm_for_inference(m) == n_bn >= 40
    assert not any(isinstance(x, torch.nn.BatchNorm2d) for x in m.modules())
    fused = [x for x in m.modules() if isinstance(x, harness.FoldedBiasAct)]
    assert len(fused) >= 12 and all(f.bias.dtype == torch.float32 for f in fused)
    assert {f.kind for f in fused} == {"relu", "leaky_relu_0.1"}
    n_cast = harness.cast_weights_for_bf16_inference(m)
    assert n_cast > 100
    for mod in m.modules():
        if isinstance(mod, hvs_b200.ManifoldHyperConnection):
            assert all(p.dtype == torch.float32 for p in mod.parameters())
    inside = {id(x) for mod in m.modules() if isinstance(mod, hvs_b200.ManifoldHyperConnection) for x in mod.modules()}
    for mod in m.modules():
        if isinstance(mod, (torch.nn.Conv2d, torch.nn.Linear)) and id(mod) not in inside:
            assert mod.weight.dtype == torch.bfloat16
            assert mod.bias is None or mod.bias.dtype == torch.bfloat16
