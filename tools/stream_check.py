"""Development aid: streaming (batch-1, one CUDA graph) p50 / p99 and the batch-64 step of the inference-prepared model in
one process -- run it under HVS_PDL=0 / 1 (or any other library switch) for an A/B on one box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hvs_b200 import harness
from hvs_b200.hybrid_vision import to_channels_last
dev = torch.device("cuda", 0)
m = harness.build_model(dev, seed=0).eval()
harness.fold_batchnorm_for_inference(m); to_channels_last(m); harness.cast_weights_for_bf16_inference(m)
b = harness.streaming_latency(m, dev, frames=200)
r = harness.inference_sharded(m, dev, 1, 0, 64, 640)
print("HVS_PDL", os.environ.get("HVS_PDL"), "streaming:", {k: round(b[k], 3) for k in b if "p50" in k or "p99" in k},
      b.get("bitwise_identical_repeats"), "batch64 ms", round(r["ms_per_step"], 2))
