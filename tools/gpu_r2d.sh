#!/bin/bash
# round 2, fourth pass: GPU tests with the new defaults, first bench line with `extra`, load-path probe, parity margins
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
python tools/probe_loads.py > gpurun_out/probe_loads.txt 2>&1; cat gpurun_out/probe_loads.txt
rm -f gpurun_out/variants.log
T="python tools/profile_target.py --passes 4 --theory 0"
$T >> gpurun_out/variants.log 2>&1
$T --batch 16384 --rsd dispersion >> gpurun_out/variants.log 2>&1
$T --batch 16384 --rsd dispersion --newton 2 >> gpurun_out/variants.log 2>&1
$T --batch 16384 --rsd dispersion --ilp 2 >> gpurun_out/variants.log 2>&1
$T --batch 16384 --aniso 1 >> gpurun_out/variants.log 2>&1
cut -c1-20,60-400 gpurun_out/variants.log
python tools/parity_report.py > gpurun_out/parity_report.jsonl 2> gpurun_out/parity_report.err; echo "parity rc=$?"
cut -c1-420 gpurun_out/parity_report.jsonl
cut -c1-1500 gpurun_out/bench.json
