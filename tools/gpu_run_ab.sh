#!/bin/bash
# A/B of library builds on one box: usage gpu_run_ab.sh <tag> <lib1> <lib2> ...   ("default" = the in-tree library)
set -u
mkdir -p gpurun_out
T=$1; shift
for L in "$@"; do
  if [ "$L" = "default" ]; then unset DGAN_LIB; N=default; else export DGAN_LIB=$PWD/$L; N=$(basename $L .so); fi
  timeout 600 python bench.py --steps 8 --warmup 3 --no_extra --cpu_sample 0 --no_profile > gpurun_out/${T}_bench_${N}.json 2> gpurun_out/${T}_bench_${N}.err
  timeout 300 python tools/loop_stalls.py mnist 256 50 > gpurun_out/${T}_stalls_${N}.log 2>&1
done
