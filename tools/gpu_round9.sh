#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/variants4.log
for v in "" "--sigma-v 1.0" "--sigma-v 100" "--sigma-v 500" "--ilp 2" "--ilp 2 --sigma-v 1.0"; do
  python tools/profile_target.py --passes 3 $v >> gpurun_out/variants4.log 2>&1
done
cat gpurun_out/variants4.log
python bench.py --workload mcmc --steps 5 --warmup 3 2>/dev/null | cut -c1-700
