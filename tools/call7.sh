mkdir -p gpurun_out
python -m pytest tests/test_gpu_detect.py tests/test_gpu_modules.py tests/test_gpu_hybrid.py -m gpu -x -q > gpurun_out/t_detect.log 2>&1; echo "detect+modules+hybrid tests rc=$?"; tail -3 gpurun_out/t_detect.log
python -c "
from hvs_b200 import harness; import json
for b in (0.0, -4.0):
    d = harness.detect_tail('cuda:0', objectness_bias=b); print(json.dumps({k: d[k] for k in ('ms_per_batch','decode_ms','nms_ms','decode_GBps')}), d['graph_replay']['decode_ms'], d['graph_replay']['nms_ms'])
" 2>&1 | tail -3
