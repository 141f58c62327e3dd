#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" > gpurun_out/summary.txt; tail -6 gpurun_out/pytest_gpu.log
python tools/sanitize_target.py > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/sanitize_plain.log; exit 1; }
timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/memcheck.log python tools/sanitize_target.py > gpurun_out/sanitize_run.log 2>&1
echo "sanitizer rc=$?" >> gpurun_out/summary.txt
tail -3 gpurun_out/sanitize_run.log; tail -6 gpurun_out/memcheck.log; cat gpurun_out/summary.txt
