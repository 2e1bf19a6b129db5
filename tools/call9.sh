mkdir -p gpurun_out
python -m pytest tests/test_gpu_k2.py tests/test_gpu_hybrid.py tests/test_gpu_modules.py -m gpu -x -q > gpurun_out/t_k2.log 2>&1; echo "k2+hybrid+modules tests rc=$?"; tail -4 gpurun_out/t_k2.log
python - <<'PY' 2>&1 | tail -4
import torch, copy, json
from hvs_b200 import harness
from hvs_b200.hybrid_vision import to_channels_last
dev = torch.device("cuda", 0)
m = harness.build_model(dev, seed=0).eval()
harness.fold_batchnorm_for_inference(m); to_channels_last(m)
a = harness.streaming_latency(m, dev, frames=200)
print("before cast:", {k: a[k] for k in a if "p50" in k or "p99" in k or "launch" in k})
harness.cast_weights_for_bf16_inference(m)
b = harness.streaming_latency(m, dev, frames=200)
print("after cast :", {k: b[k] for k in b if "p50" in k or "p99" in k or "launch" in k})
r = harness.inference_sharded(m, dev, 1, 0, 64, 640)
print("batch 64:", r["ms_per_step"], 64 / r["ms_per_step"] * 1e3)
PY
