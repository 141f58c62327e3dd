#!/bin/bash
# 2-GPU check after the ABI change: NCCL test, all three bench workloads at N = 1 and N = 2
mkdir -p gpurun_out
python -m pytest tests/test_multi_gpu.py -m gpu -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/pytest_multi.log
for W in boss dense mcmc; do
  python bench.py --workload $W --no-cpu > gpurun_out/bench_${W}_n1.json 2> gpurun_out/bench_${W}_n1.err; echo "$W n1 rc=$?"
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 \
      bench.py --workload $W --gpus 2 --no-cpu > gpurun_out/bench_${W}_n2.json 2> gpurun_out/bench_${W}_n2.err; echo "$W n2 rc=$?"
done
for f in gpurun_out/bench_*_n[12].json; do echo $f; cut -c1-330 $f; done
