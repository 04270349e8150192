#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-r2j}
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
rc=$?
echo "smoke rc=$rc" >> gpurun_out/${T}_smoke.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${T}_tests_quick.log
timeout 300 python tools/loop_stalls.py mnist 256 50 > gpurun_out/${T}_stalls.log 2>&1
timeout 300 python tools/loop_trace.py mnist 256 30 gpurun_out/${T}_trace.npz > gpurun_out/${T}_trace.log 2>&1
for B in 256 50 512 1024; do
  timeout 600 python bench.py --steps 6 --warmup 3 --cpu_sample 0 --no_profile --no_extra --batch $B > gpurun_out/${T}_bench_b${B}.json 2> gpurun_out/${T}_bench_b${B}.err
done
timeout 300 python tools/loop_stalls.py mnist 50 50 > gpurun_out/${T}_stalls_b50.log 2>&1
timeout 300 python tools/loop_stalls.py mnist 1024 30 > gpurun_out/${T}_stalls_b1024.log 2>&1
