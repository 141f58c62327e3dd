#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "small or batch_invariances or empty_and_large" 2>&1 | tail -3
python tools/probe_small_kernels.py > gpurun_out/small_kernels.txt 2>&1; cat gpurun_out/small_kernels.txt | tail -14
python tools/probe_latency.py > gpurun_out/latency.txt 2>&1; cat gpurun_out/latency.txt
