#!/bin/bash
# scaling run on one box: bench.py (headline + extras) at N = 1, 2, 4, 8 (whatever is visible), plus the reference arm
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
echo "gpus visible: $NG" > gpurun_out/scale_summary.txt
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/scale_ref.json 2>/dev/null
python bench.py --steps 10 --warmup 3 > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
for N in 2 4 8; do
  if [ $N -le $NG ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) \
      bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
    echo "N=$N rc=$?"
  fi
done
python - <<'P' | tee -a gpurun_out/scale_summary.txt
import glob, json
for fn in sorted(glob.glob("gpurun_out/scale_n*.json")) + ["gpurun_out/scale_ref.json"]:
    for ln in open(fn):
        if not ln.startswith("{"):
            continue
        d = json.loads(ln)
        x = d.get("extra") or {}
        print(fn, d.get("n_gpus"), "value=%.4g" % d["value"], "e2e=%.4g" % d["e2e"]["value"], "ms/step=%.3f" % d["ms_per_step"],
              (d.get("clocks") or {}).get("sm_mhz"), (d.get("clocks") or {}).get("reasons"),
              {k: (round(v["value"]), v.get("ms_per_table") or v.get("ms_per_sweep") or (v.get("latency_us") or {}).get("median"))
               for k, v in x.items()})
P
