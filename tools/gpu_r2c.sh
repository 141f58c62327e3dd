#!/bin/bash
# round 2, third pass: GPU tests, LDS diagnostic build, parity margins of the candidate defaults
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
rm -f gpurun_out/variants.log
T="python tools/profile_target.py --passes 4 --theory 0"
$T --expdeg 5 --newton 2 >> gpurun_out/variants.log 2>&1
VICTOR_B200_LIB=$PWD/build/libvb_diag.so $T --expdeg 5 --newton 2 >> gpurun_out/variants.log 2>&1
VICTOR_B200_LIB=$PWD/build/libvb_diag.so $T --expdeg 3 --newton 2 >> gpurun_out/variants.log 2>&1
$T --batch 16384 --rsd dispersion >> gpurun_out/variants.log 2>&1
VICTOR_B200_LIB=$PWD/build/libvb_diag.so $T --batch 16384 --rsd dispersion >> gpurun_out/variants.log 2>&1
cut -c1-20,60-400 gpurun_out/variants.log
python tools/parity_report.py > gpurun_out/parity_report.jsonl 2> gpurun_out/parity_report.err; echo "parity rc=$?"
cat gpurun_out/parity_report.jsonl
