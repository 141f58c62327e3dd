#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/variants10.log
python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
python tools/parity_report.py 2>&1 | head -3
for v in "" "--newton 2" ""; do python tools/profile_target.py --passes 3 $v >> gpurun_out/variants10.log 2>&1; done
grep -o "newton=[23]\|evals/s=[0-9.e+]*\|ms=\[[^]]*\]" gpurun_out/variants10.log | paste - - -
