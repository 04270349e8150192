#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-r2g}
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
rc=$?
echo "smoke rc=$rc" >> gpurun_out/${T}_smoke.log
if [ $rc -ne 0 ]; then
  timeout 600 compute-sanitizer --tool memcheck --print-limit 30 python tools/profile_step.py 2 fp16 mnist 16 > gpurun_out/${T}_memcheck.log 2>&1
  exit 0
fi
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${T}_tests_quick.log
timeout 300 python tools/loop_stalls.py mnist 256 50 > gpurun_out/${T}_stalls.log 2>&1
timeout 300 python tools/loop_trace.py mnist 256 30 gpurun_out/${T}_trace.npz > gpurun_out/${T}_trace.log 2>&1
timeout 600 python bench.py --steps 8 --warmup 3 --cpu_sample 0 --no_profile > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | tail -40 > gpurun_out/${T}_tests.log
