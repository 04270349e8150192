"""cuBLAS bf16 8192^3 GEMM (what MEASURED_PEAKS.json's bf16_tflops is measured with): the calibration kernel for
tensor-pipe counters.  Prints its own CUDA-event TFLOP/s so the counters can be compared with a known utilisation."""
import sys
import torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dt = torch.float16 if (len(sys.argv) > 2 and sys.argv[2] == "fp16") else torch.bfloat16
a = torch.randn(n, n, device="cuda", dtype=dt)
b = torch.randn(n, n, device="cuda", dtype=dt)
for _ in range(3):
    c = a @ b
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for _ in range(5):
    e0.record()
    c = a @ b
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print("cublas %s %d^3: %.3f ms best -> %.1f TFLOP/s" % (dt, n, best, 2.0 * n ** 3 / best / 1e9))
