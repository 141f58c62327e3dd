#!/bin/bash
# round 2, fifth pass: GPU tests (k_small), n = 1 latency, non-finite dispersion rows, short bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
python tools/probe_latency.py > gpurun_out/latency.txt 2>&1; cat gpurun_out/latency.txt
python tools/check_nonfinite.py > gpurun_out/nonfinite.txt 2>&1; tail -20 gpurun_out/nonfinite.txt | cut -c1-400
python bench.py --workload mcmc --steps 10 --warmup 2 > gpurun_out/bench_mcmc.json 2> gpurun_out/bench_mcmc.err; cut -c1-600 gpurun_out/bench_mcmc.json
python tools/profile_target.py --passes 4 --theory 0 2>&1 | cut -c1-20,60-400
