#!/bin/bash
set -x
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" > gpurun_out/summary.txt
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
tail -15 gpurun_out/pytest_gpu.log
rm -f gpurun_out/variants.log
for v in "--fast 1 --threads 256" "--fast 0 --threads 256" "--fast 1 --threads 128" "--fast 1 --threads 192" "--fast 1 --threads 256 --batch 1 --passes 5" ; do
  python tools/profile_target.py $v >> gpurun_out/variants.log 2>&1
done
cat gpurun_out/variants.log
python tools/profile_target.py --passes 3 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_multipoles -s 1 -c 1 -o gpurun_out/prof_k1 python tools/profile_target.py --passes 3 > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/summary.txt
