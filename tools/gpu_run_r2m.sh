#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-r2m}
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_wide.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${T}_tests_quick.log
for CS in 4 2; do
  for B in 256 512; do
    DGAN_CS=$CS timeout 600 python bench.py --steps 6 --warmup 3 --cpu_sample 0 --no_profile --no_extra --batch $B > gpurun_out/${T}_bench_cs${CS}_b${B}.json 2> gpurun_out/${T}_bench_cs${CS}_b${B}.err
  done
done
DGAN_CS=4 timeout 600 python bench.py --steps 4 --warmup 3 --cpu_sample 0 --no_extra --batch 256 > gpurun_out/${T}_bench_cs4_prof.json 2> gpurun_out/${T}_bench_cs4_prof.err
DGAN_CS=2 timeout 600 python bench.py --steps 4 --warmup 3 --cpu_sample 0 --no_extra --batch 256 > gpurun_out/${T}_bench_cs2_prof.json 2> gpurun_out/${T}_bench_cs2_prof.err
