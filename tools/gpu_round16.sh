#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/variants9.log
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "variants_hold or example_config" 2>&1 | tail -3
for v in "--f2i 0" "--f2i 1" "--f2i 0" "--f2i 1"; do python tools/profile_target.py --passes 3 $v >> gpurun_out/variants9.log 2>&1; done
grep -o "f2i=[01]\|evals/s=[0-9.e+]*\|ms=\[[^]]*\]" gpurun_out/variants9.log | paste - - -
