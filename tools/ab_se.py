import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hvs_b200 import harness
from hvs_b200.hybrid_vision import to_channels_last, ConvMHCLayer
dev = torch.device("cuda", 0)
m = harness.build_model(dev, seed=0).eval()
harness.fold_batchnorm_for_inference(m); to_channels_last(m); harness.cast_weights_for_bf16_inference(m)
def setf(v):
    for mod in m.modules():
        if isinstance(mod, ConvMHCLayer): mod.fuse_se_gate = v
for rep in range(2):
    for v in (False, True):
        setf(v)
        b = harness.streaming_latency(m, dev, frames=150)
        r = harness.inference_sharded(m, dev, 1, 0, 64, 640)
        print(f"fuse_se_gate={v}: streaming p50 {b['p50_ms']:.3f} ms   batch 64 {r['ms_per_step']:.2f} ms = {64 / r['ms_per_step'] * 1e3:.1f} img/s", flush=True)
