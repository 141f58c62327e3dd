#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/variants8.log
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for v in "--aniso 1" "--aniso 1 --fast 0" "--rsd dispersion --batch 16384" "--rsd dispersion --batch 16384 --fast 0" "--rsd kaiser" "--rsd kaiser --fast 0"; do
  python tools/profile_target.py --passes 3 $v >> gpurun_out/variants8.log 2>&1
done
cut -c1-60,140-300 gpurun_out/variants8.log
