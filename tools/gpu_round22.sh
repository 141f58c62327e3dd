#!/bin/bash
# general kernel: parity tests, then timings of its variants
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
rm -f gpurun_out/variants_general.log
for v in "--rsd dispersion" "--aniso 1" "--rsd kaiser" "--rsd dispersion --aniso 1" "$@"; do
  python tools/profile_target.py --passes 4 --theory 0 $v >> gpurun_out/variants_general.log 2>&1
done
cut -c1-30,120-400 gpurun_out/variants_general.log
