#!/bin/bash
mkdir -p gpurun_out
python tools/small_call_target.py > gpurun_out/small_plain.log 2>&1 && \
ncu --set full --cache-control none --clock-control none --import-source on -k regex:k_small -s 3 -c 1 -f -o gpurun_out/prof_small \
    python tools/small_call_target.py > gpurun_out/ncu_small.log 2>&1
ncu -i gpurun_out/prof_small.ncu-rep --page raw --csv > gpurun_out/small_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_small.ncu-rep --page source --csv > gpurun_out/small_src.csv 2>/dev/null
ls -la gpurun_out/prof_small.ncu-rep; tail -3 gpurun_out/ncu_small.log
