#!/bin/bash
# ncu full capture of k_multipoles for the given profile_target.py arguments
mkdir -p gpurun_out
python tools/profile_target.py --passes 3 "$@" > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_multipoles -s 1 -c 1 -f -o gpurun_out/prof_k1 python tools/profile_target.py --passes 3 "$@" > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/plain2.log; tail -3 gpurun_out/ncu_full.log
