#!/bin/bash
# ncu full capture of k_multipoles_general for the given profile_target.py arguments
mkdir -p gpurun_out
python tools/profile_target.py --passes 3 --batch 8192 --theory 0 "$@" > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_multipoles_general -s 1 -c 1 -f -o gpurun_out/prof_k1g \
    python tools/profile_target.py --passes 3 --batch 8192 --theory 0 "$@" > gpurun_out/ncu_full_g.log 2>&1
ncu -i gpurun_out/prof_k1g.ncu-rep --page raw --csv > gpurun_out/k1g_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_k1g.ncu-rep --page source --csv > gpurun_out/k1g_src.csv 2>/dev/null
cat gpurun_out/plain3.log | cut -c1-30,120-300; tail -3 gpurun_out/ncu_full_g.log
