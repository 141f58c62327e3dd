#!/bin/bash
# round 2, second pass: all GPU tests, exp-table variants of K1
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
rm -f gpurun_out/variants.log
T="python tools/profile_target.py --passes 4 --theory 0"
for v in "--expdeg 5 --newton 3" "--expdeg 5 --newton 2" "--expdeg 3 --newton 3" "--expdeg 3 --newton 2" "--expdeg 30 --newton 2" \
         "--expdeg 3 --newton 2 --theory 1"; do
  $T $v >> gpurun_out/variants.log 2>&1
done
cut -c1-20,60-400 gpurun_out/variants.log
