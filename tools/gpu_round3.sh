#!/bin/bash
# One GPU call: smoke + GPU tests + bench + ncu launch list + ncu full captures (K1, K2).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" > gpurun_out/summary.txt
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
tail -15 gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench_ref rc=$?" >> gpurun_out/summary.txt
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_short.json 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_list.log 2>&1
python tools/profile_target.py --passes 3 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_multipoles -s 1 -c 1 -f -o gpurun_out/prof_k1 \
    python tools/profile_target.py --passes 3 > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_chi2 -s 1 -c 1 -f -o gpurun_out/prof_k2 \
    python tools/profile_target.py --passes 3 > gpurun_out/ncu_full_k2.log 2>&1
cat gpurun_out/summary.txt; cat gpurun_out/bench.json
