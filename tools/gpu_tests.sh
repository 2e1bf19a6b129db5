#!/bin/bash
# Run every GPU test file in its own process under a hard timeout (a hung kernel must not take the box down);
# logs to gpurun_out/tests_<file>.log, one summary line per file on stdout.
mkdir -p gpurun_out
files="${@:-tests/test_gpu_k2.py tests/test_gpu_sinkhorn.py tests/test_gpu_modules.py tests/test_gpu_detect.py tests/test_gpu_mhc_stream.py}"
rc_all=0
for f in $files; do
  name=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu --timeout 300 --timeout-method=thread -x --no-header -rf > gpurun_out/tests_$name.log 2>&1
  rc=$?
  [ $rc -ne 0 ] && rc_all=1
  echo "== $f rc=$rc: $(tail -1 gpurun_out/tests_$name.log)"
done
exit $rc_all
