#!/bin/bash
# 2-GPU call: GPU tests (incl. NCCL world_size 2), bench at N=1 and N=2, dense and mcmc workloads
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi_L.txt
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" > gpurun_out/summary.txt
tail -8 gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench1 rc=$?" >> gpurun_out/summary.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?" >> gpurun_out/summary.txt
python bench.py --workload dense --batch 32768 --steps 5 --warmup 3 > gpurun_out/bench_dense_n1.json 2> gpurun_out/bench_dense_n1.err; echo "dense1 rc=$?" >> gpurun_out/summary.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload dense --batch 32768 --steps 5 --warmup 3 > gpurun_out/bench_dense_n2.json 2> gpurun_out/bench_dense_n2.err; echo "dense2 rc=$?" >> gpurun_out/summary.txt
python bench.py --workload mcmc --steps 5 --warmup 3 > gpurun_out/bench_mcmc_n1.json 2> gpurun_out/bench_mcmc_n1.err; echo "mcmc1 rc=$?" >> gpurun_out/summary.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload mcmc --steps 5 --warmup 3 > gpurun_out/bench_mcmc_n2.json 2> gpurun_out/bench_mcmc_n2.err; echo "mcmc2 rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
for f in bench_n1 bench_n2 bench_dense_n1 bench_dense_n2 bench_mcmc_n1 bench_mcmc_n2; do echo "== $f"; cut -c1-420 gpurun_out/$f.json; tail -3 gpurun_out/$f.err; done
