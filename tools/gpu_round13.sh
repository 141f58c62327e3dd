#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/variants7.log
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for v in "" "--aniso 1" "--rsd dispersion --batch 16384" "--rsd kaiser" "--rsd dispersion --aniso 1 --batch 16384"; do
  python tools/profile_target.py --passes 3 $v >> gpurun_out/variants7.log 2>&1
done
cat gpurun_out/variants7.log
