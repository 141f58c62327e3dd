#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" > gpurun_out/summary.txt
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
tail -30 gpurun_out/pytest_gpu.log
python tools/profile_target.py --passes 3 > gpurun_out/plain2.log 2>&1; cat gpurun_out/plain2.log
cat gpurun_out/summary.txt
