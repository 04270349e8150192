#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-trace}
timeout 300 python tools/loop_trace.py mnist 256 30 gpurun_out/${T}_mnist.npz > gpurun_out/${T}_mnist.log 2>&1
timeout 300 python tools/loop_stalls.py mnist 256 50 > gpurun_out/${T}_stalls.log 2>&1
