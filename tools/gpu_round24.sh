#!/bin/bash
# evidence for the committed kernels: smoke, GPU tests, bench line + reference arm, ncu launch list of the same
# bench command, full captures of K1 (tuned) and of the general kernel (dispersion)
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_short.json 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_list.log 2>&1
python tools/profile_target.py --passes 3 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_multipoles -s 1 -c 1 -f -o gpurun_out/prof_k1 \
    python tools/profile_target.py --passes 3 > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/prof_k1.ncu-rep --page raw --csv > gpurun_out/k1_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_k1.ncu-rep --page source --csv > gpurun_out/k1_src.csv 2>/dev/null
cat gpurun_out/bench.json | cut -c1-300; cat gpurun_out/plain2.log | cut -c1-50,150-300
