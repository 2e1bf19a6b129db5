mkdir -p gpurun_out
python tools/prof_decode.py 4 2>&1 | tail -4
ncu --set full --clock-control none --import-source on -k regex:decode_four -c 3 -o gpurun_out/r2_decode python tools/prof_decode.py 1 > gpurun_out/ncu_decode.log 2>&1; tail -2 gpurun_out/ncu_decode.log
ncu -i gpurun_out/r2_decode.ncu-rep --page raw --csv > gpurun_out/r2_decode_raw.csv 2>/dev/null; ls -la gpurun_out/r2_decode*
