#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/variants11.log
python tools/parity_report.py 2>&1 | grep -E "default|sloppy"
for v in "" "--sloppy 1" "" "--sloppy 1"; do python tools/profile_target.py --passes 3 $v >> gpurun_out/variants11.log 2>&1; done
grep -o "sloppy=[01]\|evals/s=[0-9.e+]*\|ms=\[[^]]*\]" gpurun_out/variants11.log | paste - - -
