mkdir -p gpurun_out
python -m pytest tests/test_gpu_detect.py -m gpu -x -q > gpurun_out/t_detect.log 2>&1; echo "detect tests rc=$?"; tail -3 gpurun_out/t_detect.log
python -c "
from hvs_b200 import harness; import json
for b in (0.0, -4.0):
    d = harness.detect_tail('cuda:0', objectness_bias=b); print(json.dumps({k: d[k] for k in ('ms_per_batch','decode_ms','nms_ms','decode_GBps','graph_replay')}))
" 2>&1 | tail -3
for v in trace tracenohmma; do HVS_VARIANT=$v timeout 120 python tools/time_fused.py 1048576 > gpurun_out/trace_$v.txt 2>&1; tail -1 gpurun_out/trace_$v.txt; done
