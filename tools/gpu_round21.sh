#!/bin/bash
# fused likelihood epilogue: parity tests, then timing fused vs separate K2, with and without theory output
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
rm -f gpurun_out/variants_fuse.log
for v in "--fuse 0 --theory 1" "--fuse 1 --theory 1" "--fuse 1 --theory 0" "--fuse 0 --theory 0" "--fuse 1 --theory 0 --rsd dispersion" "--fuse 0 --theory 0 --rsd dispersion"; do
  python tools/profile_target.py --passes 5 $v >> gpurun_out/variants_fuse.log 2>&1
done
cat gpurun_out/variants_fuse.log
