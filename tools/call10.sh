mkdir -p gpurun_out
python -m pytest tests/test_gpu_edges.py tests/test_gpu_hybrid.py -m gpu -x -q > gpurun_out/t_edges.log 2>&1; echo "edges+hybrid tests rc=$?"; tail -4 gpurun_out/t_edges.log | cut -c1-300
python - <<'PY' 2>&1 | tail -4
import torch, copy, json
from hvs_b200 import harness
from hvs_b200.hybrid_vision import to_channels_last
dev = torch.device("cuda", 0)
m = harness.build_model(dev, seed=0).eval()
harness.fold_batchnorm_for_inference(m); to_channels_last(m); harness.cast_weights_for_bf16_inference(m)
b = harness.streaming_latency(m, dev, frames=200)
print("streaming:", {k: b[k] for k in b if "p50" in k or "p99" in k})
r = harness.inference_sharded(m, dev, 1, 0, 64, 640)
print("batch 64:", r["ms_per_step"], 64 / r["ms_per_step"] * 1e3)
PY
