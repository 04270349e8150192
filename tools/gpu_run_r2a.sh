#!/bin/bash
# Round-2 GPU call A: parity tests (incl. the wide ones), a bench line, and the tensor-pipe counter calibration
# (cuBLAS bf16 GEMM vs our tcgen05 kernels) asked for by VERDICT item 6.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
nproc > gpurun_out/r2a_nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -80 > gpurun_out/r2a_tests.log
echo "tests rc=$?" >> gpurun_out/r2a_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
M="sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor_subpipe_hmma.sum,sm__inst_executed_pipe_tensor.sum,sm__ops_path_tensor_op_hmma_src_fp16_dst_fp32.sum,sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32.sum,sm__ops_path_tensor_op_hmma_src_fp16_dst_fp32.sum.pct_of_peak_sustained_elapsed,sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32.sum.pct_of_peak_sustained_elapsed,sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_uniform.sum,sm__cycles_elapsed.max,gpu__time_duration.sum,l1tex__m_xbar2l1tex_read_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed"
timeout 600 ncu --metrics $M --clock-control none -k regex:"gemm|cutlass|nvjet|xmma" -c 4 --csv --log-file gpurun_out/r2a_calib_cublas.csv python tools/cublas_gemm.py 8192 bf16 > gpurun_out/r2a_calib_cublas.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:"gemm|cutlass|nvjet|xmma" -c 4 --csv --log-file gpurun_out/r2a_calib_cublas_fp16.csv python tools/cublas_gemm.py 8192 fp16 > gpurun_out/r2a_calib_cublas_fp16.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:tc_bsgemm2 -s 18 -c 18 --csv --log-file gpurun_out/r2a_calib_ours.csv python tools/profile_step.py 4 fp16 mnist 256 > gpurun_out/r2a_calib_ours.log 2>&1
python tools/cublas_gemm.py 8192 bf16 > gpurun_out/r2a_cublas_noprof.log 2>&1
ls -la gpurun_out | tail -20
