#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/variants5.log
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "streaming_golden or variants_hold" 2>&1 | tail -3
for v in "--minblocks 4" "--minblocks 3" "--minblocks 2" "--minblocks 1 --ilp 8"; do
  python tools/profile_target.py --passes 3 $v >> gpurun_out/variants5.log 2>&1
done
cat gpurun_out/variants5.log
