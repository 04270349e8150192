#!/bin/bash
# Round-2 GPU call B: first light of the persistent loop kernel - smoke, then the GPU test-suite, then a short bench.
set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke.log 2>&1
rc=$?
echo "smoke rc=$rc" >> gpurun_out/r2b_smoke.log
if [ $rc -ne 0 ]; then
  timeout 900 compute-sanitizer --tool memcheck --print-limit 30 python tools/profile_step.py 2 fp16 mnist 8 > gpurun_out/r2b_memcheck.log 2>&1
  echo "memcheck rc=$?" >> gpurun_out/r2b_memcheck.log
  exit 0
fi
timeout 1800 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -80 > gpurun_out/r2b_tests.log
echo "tests rc=${PIPESTATUS[0]}" >> gpurun_out/r2b_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
echo "bench rc=$?" >> gpurun_out/r2b_bench.err
