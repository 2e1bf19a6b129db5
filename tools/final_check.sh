# Round-end validation on one B200: GPU tests, smoke, the default bench line (add NCU=1 for the ncu captures the profiles/ summaries come from).
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/t_all.log 2>&1; echo "gpu tests rc=$?"; tail -6 gpurun_out/t_all.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
( time python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err ) 2>&1 | grep real; wc -c gpurun_out/bench.json
if [ "$NCU" = "1" ]; then
ncu --set full --clock-control none --import-source on -k regex:"decode|nms|gather" -c 8 -f -o gpurun_out/r2_detect python tools/prof_decode.py 1 > gpurun_out/ncu_detect.log 2>&1; tail -1 gpurun_out/ncu_detect.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline --skip-hybrid > gpurun_out/ncu_ll.log 2>&1; tail -c 300 gpurun_out/ncu_ll.log
fi
