#!/bin/bash
# evidence run (what profiles/r2_* were produced with): launch list of the bench command, ncu --set full of one L-step, bench lines, GPU tests
set -u
mkdir -p gpurun_out
T=${1:-r3}
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log
timeout 1200 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
timeout 600 python bench.py --precision fp32 --steps 3 --warmup 3 --no_extra --cpu_sample 0 > gpurun_out/${T}_bench_fp32.json 2> gpurun_out/${T}_bench_fp32.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no_extra --cpu_sample 0 --no_profile > gpurun_out/${T}_ncu_bench.log 2>&1
timeout 1500 ncu --set full --metrics sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --import-source on -s 17 -c 8 -f -o gpurun_out/${T}_full python tools/profile_step.py 6 > gpurun_out/${T}_ncu_full.log 2>&1
timeout 2400 python -m pytest tests -m gpu -q -s 2>&1 | tail -40 > gpurun_out/${T}_tests.log
# in-kernel timeline (probe build of the library, built by __graft_entry__.build())
DGAN_LIB=$PWD/defensegan_b200/libdefensegan_b200_probe.so timeout 300 python tools/probe_step.py mnist 256 50 > gpurun_out/${T}_probe_mnist.txt 2>&1
DGAN_LIB=$PWD/defensegan_b200/libdefensegan_b200_probe.so timeout 300 python tools/probe_step.py celeba 128 50 > gpurun_out/${T}_probe_celeba.txt 2>&1
