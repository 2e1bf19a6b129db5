mkdir -p gpurun_out
python -m pytest tests/test_gpu_detect.py -m gpu -x -q > gpurun_out/t_detect.log 2>&1; echo "detect tests rc=$?"; tail -3 gpurun_out/t_detect.log
python -c "
from hvs_b200 import harness; import json
for b in (0.0, -4.0):
    d = harness.detect_tail('cuda:0', objectness_bias=b); print(json.dumps({k: d[k] for k in ('ms_per_batch','decode_ms','nms_ms','decode_GBps')}))
" 2>&1 | tail -3
for v in base late warp both nohmma nofence base; do HVS_VARIANT=$v timeout 120 python tools/ab.py 2>&1 | tail -1; sleep 2; done
