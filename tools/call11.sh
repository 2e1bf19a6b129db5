python -m pytest tests/test_gpu_edges.py -m gpu -x -q 2>&1 | tail -3 | cut -c1-300
python tools/bench_se_gate.py 64 2>&1 | tail -6; python tools/bench_se_gate.py 1 2>&1 | tail -6
