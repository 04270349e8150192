#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-r2s}
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${T}_tests_quick.log
timeout 900 python bench.py --steps 10 --warmup 3 --cpu_sample 0 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/profile_step.py 3 fp16 mnist 16 > gpurun_out/${T}_memcheck.log 2>&1
