"""Developer aid, needs a library built with -DDGAN_PROBE (DGAN_LIB=...): per tensor-core kernel instantiation, the
distribution over CTAs of the cycles from the PDL wait to the end of the CTA's work (mean over the launches of one
projection): a wide distribution = the static item assignment leaves SMs idle at the kernel boundary.
Usage: DGAN_LIB=build_ab/probe.so python tools/probe_step.py [mnist|celeba] [batch] [L]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from defensegan_b200.models.gan import dataset_gan_dict

dataset = sys.argv[1] if len(sys.argv) > 1 else "mnist"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
L = int(sys.argv[3]) if len(sys.argv) > 3 else 50
gan = dataset_gan_dict[dataset](test_mode=True, verbose=False, precision="fp16", batch_size=50)
gan.rec_rr, gan.rec_iters = 10, L
g = torch.Generator().manual_seed(0)
x = torch.rand(B, *gan.image_dim, generator=g).cuda()
z0 = (torch.randn(B * 10, 128, generator=g) * 128 ** -0.5).cuda()
lib = gan._native.lib if hasattr(gan, "_native") and gan._native is not None else None
gan.reconstruct(x, z_init_val=z0)
torch.cuda.synchronize()
lib = gan._native.lib
buf = (ctypes.c_ulonglong * (48 * 160 * 4))()
lib.dgan_debug_probe_read.restype = ctypes.c_int
lib.dgan_debug_probe_read.argtypes = [ctypes.POINTER(ctypes.c_ulonglong)]
assert lib.dgan_debug_probe_read(buf) == 0          # discard the first call (schedule upload, graph capture)
gan.reconstruct(x, z_init_val=z0)
torch.cuda.synchronize()
assert lib.dgan_debug_probe_read(buf) == 0
a = np.frombuffer(buf, dtype=np.uint64).reshape(48, 160, 4).astype(np.float64)
NT = [256, 128, 64, 48, 16]
EP = ["bias+relu", "bias", "mask", "none", "final-sigmoid", "final-tanh", "float-out", "?"]
print("kernel <N, epilogue> | launches | cycles from PDL wait to CTA end: mean / min / max over CTAs | (max-mean)/max | entry->wait mean | MMA operand wait mean (leaders)")
for k in range(48):
    cnt = a[k, :, 1]
    act = cnt > 0
    if not act.any():
        continue
    dur = a[k, act, 0] / cnt[act]
    pre = a[k, act, 2] / cnt[act]
    wf = a[k, act, 3] / cnt[act]
    lead = wf > 0
    print("<%d, %s> | %d | %.0f / %.0f / %.0f | %.3f | %.0f | %.0f" % (NT[k // 8], EP[k % 8], int(cnt[act].max()), dur.mean(), dur.min(), dur.max(),
                                                             (dur.max() - dur.mean()) / dur.max(), pre.mean(), wf[lead].mean() if lead.any() else 0))
    np.save("gpurun_out/probe_%s_%d_%s.npy" % (dataset, NT[k // 8], EP[k % 8]), a[k])
