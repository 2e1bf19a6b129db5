mkdir -p gpurun_out
python -m pytest tests/test_gpu_hybrid.py -m gpu -x -q > gpurun_out/t_hybrid.log 2>&1; echo "hybrid tests rc=$?"; tail -5 gpurun_out/t_hybrid.log
python tools/profile_hybrid.py 1 > gpurun_out/ph1.log 2>&1; head -3 gpurun_out/hybrid_profile_infer_b1.txt
python tools/profile_hybrid.py 64 > gpurun_out/ph64.log 2>&1; head -12 gpurun_out/hybrid_profile_infer_b64.txt
