#!/bin/bash
# round 2, first pass: GPU tests on the new tuned kernels, then variant timings (no profiler)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
rm -f gpurun_out/variants.log
T="python tools/profile_target.py --passes 4 --theory 0"
for v in "--expdeg 5 --newton 3" "--expdeg 5 --newton 2" "--expdeg 52 --newton 3" "--expdeg 52 --newton 2" \
         "--expdeg 53 --newton 3" "--expdeg 53 --newton 2" "--expdeg 5 --newton 3"; do
  $T $v >> gpurun_out/variants.log 2>&1
done
VICTOR_B200_LIB=$PWD/build/libvb_tail0.so $T --expdeg 5 --newton 3 >> gpurun_out/variants.log 2>&1
VICTOR_B200_LIB=$PWD/build/libvb_tail0.so $T --expdeg 53 --newton 2 >> gpurun_out/variants.log 2>&1
for v in "--rsd dispersion --tuned 0" "--rsd dispersion --tuned 1 --ilp 4" "--rsd dispersion --tuned 1 --ilp 2" \
         "--aniso 1 --tuned 0" "--aniso 1 --tuned 1 --ilp 4" "--aniso 1 --tuned 1 --ilp 2" \
         "--rsd dispersion --aniso 1 --tuned 0" "--rsd dispersion --aniso 1 --tuned 1"; do
  $T --batch 16384 $v >> gpurun_out/variants.log 2>&1
done
cut -c1-20,60-400 gpurun_out/variants.log
