#!/bin/bash
# final evidence for the committed kernels: bench line, ncu launch list of the same command, full capture of K1
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_short.json 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_list.log 2>&1
python tools/profile_target.py --passes 3 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_multipoles -s 1 -c 1 -f -o gpurun_out/prof_k1 \
    python tools/profile_target.py --passes 3 > gpurun_out/ncu_full.log 2>&1
python tools/profile_target.py --passes 3 --rsd dispersion --batch 8192 > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_multipoles_general -s 1 -c 1 -f -o gpurun_out/prof_k1g \
    python tools/profile_target.py --passes 3 --rsd dispersion --batch 8192 > gpurun_out/ncu_full_g.log 2>&1
cat gpurun_out/bench.json | cut -c1-300; cat gpurun_out/plain2.log gpurun_out/plain3.log | cut -c1-50,150-300
