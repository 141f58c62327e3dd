#!/bin/bash
# NB: this pool answers compute-sanitizer with rc 86 ("closed on this pool"); the plain run of the target still
# exercises every kernel variant once and prints values to compare with the goldens.
mkdir -p gpurun_out
python tools/sanitize_target.py > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/sanitize_plain.log; exit 1; }
timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/memcheck.log python tools/sanitize_target.py > gpurun_out/sanitize_run.log 2>&1
echo "sanitizer rc=$?"
tail -5 gpurun_out/sanitize_run.log; tail -8 gpurun_out/memcheck.log
timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/racecheck.log python tools/sanitize_target.py > gpurun_out/sanitize_race.log 2>&1
echo "racecheck rc=$?"
tail -6 gpurun_out/racecheck.log
