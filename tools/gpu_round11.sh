#!/bin/bash
mkdir -p gpurun_out
python tools/parity_report.py > gpurun_out/parity_report.jsonl 2>&1; cat gpurun_out/parity_report.jsonl
rm -f gpurun_out/variants6.log
for v in "--newton 3" "--newton 2" "--newton 3" "--newton 2"; do python tools/profile_target.py --passes 3 $v >> gpurun_out/variants6.log 2>&1; done
cat gpurun_out/variants6.log
