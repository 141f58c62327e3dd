#!/bin/bash
# BASELINE.json configs[3] as written: one 1,048,576-row dense-grid sweep sharded over 2 / 4 / 8 GPUs (strong scaling)
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
for N in 2 4 8; do
  if [ $N -le $NG ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700+N)) \
      bench.py --workload dense --sweep 1048576 --gpus $N --steps 3 --warmup 1 > gpurun_out/sweep_n$N.json 2> gpurun_out/sweep_n$N.err
    echo "N=$N rc=$?"
  fi
done
for f in gpurun_out/sweep_n*.json; do grep "^{" $f | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['n_gpus'], 'value=%.4g' % d['value'], 'e2e=%.4g' % d['e2e']['value'], 'ms/step=%.1f' % d['ms_per_step'], d['scaling'], d["config"].get("rows_total"))"; done | tee gpurun_out/sweep_summary.txt
