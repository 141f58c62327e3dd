#!/bin/bash
mkdir -p gpurun_out
python tools/probe_pipes.py > gpurun_out/probe_pipes.json 2>&1
cat gpurun_out/probe_pipes.json
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "variants or fast_math" > gpurun_out/pytest_variants.log 2>&1; tail -5 gpurun_out/pytest_variants.log
rm -f gpurun_out/variants2.log
for v in "--expdeg 6 --group 0" "--expdeg 5 --group 0" "--expdeg 6 --group 1" "--expdeg 5 --group 1" "--expdeg 5 --group 1 --threads 128" "--expdeg 5 --group 1 --threads 192"; do
  python tools/profile_target.py --passes 3 $v >> gpurun_out/variants2.log 2>&1
done
cat gpurun_out/variants2.log
