#!/bin/bash
# First light on the B200: smoke, GPU parity tests, bench, variant timings, then ncu.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
tail -5 gpurun_out/pytest_gpu.log
for v in "--fast 1 --threads 256" "--fast 0 --threads 256" "--fast 1 --threads 128" "--fast 1 --threads 192" "--fast 1 --threads 256 --batch 1 --passes 5" "--fast 1 --threads 256 --batch 256 --passes 5"; do
  python tools/profile_target.py $v >> gpurun_out/variants.log 2>&1
done
cat gpurun_out/variants.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/bench.json
python tools/profile_target.py --passes 3 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_target.py --passes 3 > gpurun_out/ncu_list.log 2>&1
python tools/profile_target.py --passes 3 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_multipoles -s 1 -c 1 -o gpurun_out/prof_k1 python tools/profile_target.py --passes 3 > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/summary.txt
