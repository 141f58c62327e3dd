#!/bin/bash
mkdir -p gpurun_out
python tools/small_call_target.py > gpurun_out/small_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/small_launches.csv \
    python tools/small_call_target.py > gpurun_out/small_ncu.log 2>&1
grep -v "^==" gpurun_out/small_launches.csv | awk -F'","' '{print $5, $NF}' | cut -c1-200 | tail -30
python -m pytest tests -m gpu -q -x -k "batch_invariances or empty_and_large or small" 2>&1 | tail -3
