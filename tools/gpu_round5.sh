#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "variants" > gpurun_out/pytest_variants.log 2>&1; tail -5 gpurun_out/pytest_variants.log
rm -f gpurun_out/variants3.log
for v in "--expdeg 5 --cache 0" "--expdeg 5 --cache 1 --ilp 4" "--expdeg 5 --cache 1 --ilp 2" "--expdeg 5 --cache 1 --ilp 1" "--expdeg 5 --cache 2 --ilp 4" "--expdeg 5 --cache 2 --ilp 1" "--expdeg 5 --cache 1 --ilp 4 --threads 192" "--expdeg 5 --cache 1 --ilp 2 --threads 128"; do
  python tools/profile_target.py --passes 3 $v >> gpurun_out/variants3.log 2>&1
done
cat gpurun_out/variants3.log
