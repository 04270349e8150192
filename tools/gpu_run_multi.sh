#!/bin/bash
# multi-GPU check: the 2-rank NCCL parity test, then the bench at every N the box has (weak scaling: 512 images per GPU)
set -u
mkdir -p gpurun_out
T=${1:-r2p}
NG=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests/test_gpu_parity_wide.py -m gpu -q -k "nccl or shard" -s 2>&1 | tail -8 > gpurun_out/${T}_nccl_test.log
for N in 2 4 8; do
  if [ $N -le $NG ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) bench.py --gpus $N --steps 5 --warmup 3 --no_extra --cpu_sample 0 --no_profile > gpurun_out/${T}_bench_n${N}.json 2> gpurun_out/${T}_bench_n${N}.err
  fi
done
