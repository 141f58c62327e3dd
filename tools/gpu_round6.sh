#!/bin/bash
mkdir -p gpurun_out
A="--passes 3 --expdeg 5 --cache 0"
B="--passes 3 --expdeg 5 --cache 1 --ilp 4"
python tools/profile_target.py $A > gpurun_out/plainA.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_multipoles -s 1 -c 1 -f -o gpurun_out/prof_k1_A python tools/profile_target.py $A > gpurun_out/ncuA.log 2>&1
python tools/profile_target.py $B > gpurun_out/plainB.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_multipoles -s 1 -c 1 -f -o gpurun_out/prof_k1_B python tools/profile_target.py $B > gpurun_out/ncuB.log 2>&1
cat gpurun_out/plainA.log gpurun_out/plainB.log
