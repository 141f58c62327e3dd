#!/bin/bash
# quick check: smoke + GPU tests + variant timings (no profiler)
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" > gpurun_out/summary.txt
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
tail -5 gpurun_out/pytest_gpu.log
rm -f gpurun_out/variants.log
for v in "--fast 1 --threads 256" "--fast 1 --threads 128" "--fast 1 --threads 192" "--fast 0 --threads 256" "$@"; do
  python tools/profile_target.py $v >> gpurun_out/variants.log 2>&1
done
cat gpurun_out/variants.log gpurun_out/summary.txt
