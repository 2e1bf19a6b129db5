"""CPU whole-model baseline (TEST INFRASTRUCTURE, see oracle/__init__.py): the host wiring of
hvs_b200/hybrid_vision.py (plain torch composition, pinned against the reference's HybridVisionSystem on CPU by
tests/test_hybrid_cpu.py) with the ORACLE's CPU arithmetic as leaves -- mhc_ref.mhc_module_forward for every
ManifoldHyperConnection (manifold_layers.py:223-280), mhc_ref.rms_norm (:449-456), detect_ref decode / post_process
(yolo_head.py:220-294, :571-676).  It exists because /root/reference does not travel to the GPU box and the product has
no CPU path; bench.py's cpu_baseline leg times it on the host cores ("kind": "port").
"""
from __future__ import annotations

import contextlib
import math
import time
from typing import Dict

import torch
import torch.nn as nn

from . import detect_ref, mhc_ref


class OracleMHC(nn.Module):
    """Reference constructor / parameter names (manifold_layers.py:129-189); forward = the oracle restatement."""

    def __init__(self, input_dim, expansion_rate=4, hidden_dim=None, alpha=0.01, sk_iterations=20, use_mixed_precision=True,
                 dropout_rate=0.1):
        super().__init__()
        self.input_dim, self.hidden_dim = input_dim, hidden_dim or input_dim * expansion_rate
        self.sk_iterations = sk_iterations
        h = self.hidden_dim
        self.H_pre_raw = nn.Parameter(torch.empty(input_dim, h))
        self.H_post_raw = nn.Parameter(torch.empty(h, input_dim))
        self.H_res_raw = nn.Parameter(torch.empty(input_dim, input_dim))
        self.mlp = nn.Sequential(nn.Linear(h, 2 * h), nn.GELU(), nn.Dropout(dropout_rate), nn.Linear(2 * h, h), nn.GELU(), nn.Dropout(dropout_rate))
        self.norm_pre, self.norm_post = nn.LayerNorm(input_dim), nn.LayerNorm(input_dim)
        for w in (self.H_pre_raw, self.H_post_raw, self.H_res_raw):
            nn.init.xavier_uniform_(w, gain=0.1)
        for m in self.mlp:
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight, gain=math.sqrt(2))
                nn.init.zeros_(m.bias)

    def forward(self, x):
        p = {k: v for k, v in self.named_parameters()}
        return mhc_ref.mhc_module_forward(x, p, self.sk_iterations)     # recomputes the Sinkhorn projection per call, like the reference


class OracleRMSNorm(nn.Module):
    def __init__(self, dim, eps=1e-8):
        super().__init__()
        self.scale, self.eps = nn.Parameter(torch.ones(dim)), eps

    def forward(self, x):
        return mhc_ref.rms_norm(x, self.scale, self.eps)


class _OracleDecoder(nn.Module):
    def forward(self, predictions, anchors, grid_size=None, want_scores=True):
        a = predictions.shape[1]
        return detect_ref.yolo_decode(predictions, anchors.reshape(a, -1, 4)[:, 0, 2:4])


@contextlib.contextmanager
def _oracle_leaves():
    from hvs_b200 import detection, hybrid_vision
    saved = (hybrid_vision.ManifoldHyperConnection, hybrid_vision.RMSNorm, detection.ManifoldHyperConnection)
    hybrid_vision.ManifoldHyperConnection, hybrid_vision.RMSNorm = OracleMHC, OracleRMSNorm
    detection.ManifoldHyperConnection = OracleMHC
    try:
        yield
    finally:
        hybrid_vision.ManifoldHyperConnection, hybrid_vision.RMSNorm, detection.ManifoldHyperConnection = saved


def build_cpu_model(seed: int = 0):
    from hvs_b200 import hybrid_vision
    torch.manual_seed(seed)
    with _oracle_leaves():
        model = hybrid_vision.HybridVisionSystem({"num_classes": 80, "image_size": 640}).eval()
    model.detection_head.decoder = _OracleDecoder()
    return model


def cpu_forward(model, x: torch.Tensor) -> Dict:
    with _oracle_leaves(), torch.no_grad():          # the host functions test isinstance(..., ManifoldHyperConnection) at call time
        return model(x)


def cpu_inference_img_per_s(batch: int = 1, image: int = 640, steps: int = 2, warmup: int = 1, with_nms: bool = True):
    """BASELINE config 1: CPU forward (+ decode + two-stage NMS), fp32, all host threads.  Returns (img/s, s/step, threads)."""
    torch.set_num_threads(max(1, torch.get_num_threads()))
    model = build_cpu_model()
    x = torch.randn(batch, 3, image, image, generator=torch.Generator().manual_seed(0))
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        out = cpu_forward(model, x)
        if with_nms:
            detect_ref.post_process(list(out["decoded"].values()), 0.25, 0.45, 100)
        if i >= warmup:
            ts.append(time.perf_counter() - t0)
    dt = sum(ts) / len(ts)
    return batch / dt, dt, torch.get_num_threads()
