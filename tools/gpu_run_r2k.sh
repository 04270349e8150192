#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-r2k}
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 2400 python -m pytest tests -m gpu -q -s 2>&1 | tail -40 > gpurun_out/${T}_tests.log
