#!/bin/bash
# one 8-GPU box, final state: the bench line at N = 1 and N = 8 (CPU leg skipped: the reference arm is in r02s / r02w)
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/scale8_n1.json 2> gpurun_out/scale8_n1.err; echo "N=1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29608 \
  bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu > gpurun_out/scale8_n8.json 2> gpurun_out/scale8_n8.err; echo "N=8 rc=$?"
python - <<'P'
import json
for n in (1, 8):
    d = json.loads([l for l in open(f"gpurun_out/scale8_n{n}.json") if l.startswith("{")][0])
    x = d["extra"]
    print(n, "value=%.4g" % d["value"], "e2e=%.4g" % d["e2e"]["value"], "ms=%.3f" % d["ms_per_step"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"],
          {k: (round(v["value"]), v.get("ms_per_table") or v.get("ms_per_sweep") or (v.get("latency_us") or {}).get("median")) for k, v in x.items()})
P
