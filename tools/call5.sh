mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/t_all.log 2>&1; echo "gpu tests rc=$?"; tail -6 gpurun_out/t_all.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
( time python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err ) 2>&1 | grep real; echo "bench rc=$?"; wc -c gpurun_out/bench.json
ncu --set full --clock-control none --import-source on -k regex:decode_four -c 3 -o gpurun_out/r2_decode -f python tools/prof_decode.py 1 > gpurun_out/ncu_decode.log 2>&1; tail -1 gpurun_out/ncu_decode.log
python tools/prof_decode.py 4 2>&1 | tail -2
