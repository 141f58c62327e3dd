#!/bin/bash
# two GPUs: NCCL gather test, bench at N = 1 and N = 2 with the extra block
mkdir -p gpurun_out
python -m pytest tests/test_multi_gpu.py tests/test_gpu_parity.py -m gpu -q -k "two_ranks or multi_device" > gpurun_out/pytest_n2.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_n2.log
python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench n1 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/bench_n2.err
python - <<'PY'
import json
for n in (1, 2):
    d = json.load(open(f"gpurun_out/bench_n{n}.json"))
    x = d["extra"]
    print(n, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 4),
          "| disp", round(x["dispersion"]["value"]), "dense", round(x["dense_sweep"]["value"]), "ms", round(x["dense_sweep"]["ms_per_sweep"], 1),
          "mcmc", round(x["mcmc"]["value"]), x["mcmc"]["latency_us"]["median"], "strong64k", round(x["strong_64k"]["value"]), round(x["strong_64k"]["ms_per_table"], 2),
          "sustained", round(x["sustained"]["value"]), x["sustained"]["clocks"])
PY
