#!/bin/bash
# Evidence for the committed kernels, one GPU: smoke, GPU tests, bench line + reference arm, ncu launch list of the
# bench command, `ncu --set full` captures of the tuned kernel (streaming, dispersion) and of k_small.
# Profiling passes only run after the same command has exited 0 without ncu; numbers printed under ncu are not used.
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log | cut -c1-200
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "reference arm rc=$?"
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
SHORT="bench.py --steps 2 --warmup 3 --no-cpu --sustain 0 --sweep 131072"
python $SHORT > gpurun_out/bench_short.json 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python $SHORT > gpurun_out/ncu_list.log 2>&1
capture() {   # name, kernel regex, skip, profile_target / small_call_target arguments...
  local name=$1 regex=$2 skip=$3 target=$4; shift 4
  python tools/$target "$@" > gpurun_out/plain_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -f -o gpurun_out/prof_$name \
      python tools/$target "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i gpurun_out/prof_$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_$name.ncu-rep --page source --csv > gpurun_out/${name}_src.csv 2>/dev/null
  rm -f gpurun_out/prof_$name.ncu-rep
}
capture k1 k_multipoles 1 profile_target.py --passes 3
capture disp k_multipoles 1 profile_target.py --passes 3 --batch 16384 --rsd dispersion --theory 0
capture aniso k_multipoles 1 profile_target.py --passes 3 --batch 16384 --aniso 1 --theory 0
capture small k_small 3 small_call_target.py
capture k2 k_chi2 1 profile_target.py --passes 3
cut -c1-400 gpurun_out/bench.json
