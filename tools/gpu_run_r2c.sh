#!/bin/bash
# Round-2 GPU call C: loop kernel with store warps / dependency prefetch - smoke, stall counters, tests, bench.
set -u
mkdir -p gpurun_out
T=${1:-r2c}
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
rc=$?
echo "smoke rc=$rc" >> gpurun_out/${T}_smoke.log
if [ $rc -ne 0 ]; then
  timeout 900 compute-sanitizer --tool memcheck --print-limit 30 python tools/profile_step.py 2 fp16 mnist 8 > gpurun_out/${T}_memcheck.log 2>&1
  echo "memcheck rc=$?" >> gpurun_out/${T}_memcheck.log
  exit 0
fi
timeout 300 python tools/loop_stalls.py mnist 256 50 > gpurun_out/${T}_stalls.log 2>&1
timeout 300 python tools/loop_stalls.py celeba 128 20 >> gpurun_out/${T}_stalls.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench rc=$?" >> gpurun_out/${T}_bench.err
timeout 1800 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -60 > gpurun_out/${T}_tests.log
echo "tests rc=${PIPESTATUS[0]}" >> gpurun_out/${T}_tests.log
