mkdir -p gpurun_out
python tools/prof_decode.py 3 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:"decode|nms|gather" -c 12 -f -o gpurun_out/r2_detect python tools/prof_decode.py 1 > gpurun_out/ncu_detect.log 2>&1; tail -1 gpurun_out/ncu_detect.log
ncu -i gpurun_out/r2_detect.ncu-rep --page raw --csv > gpurun_out/r2_detect_raw.csv 2>/dev/null; ls -la gpurun_out/r2_detect*
