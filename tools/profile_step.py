"""Small driver for ncu: BASELINE configs[1] shape (MNIST, B=256, R=10) at a short horizon so that
a profiler replaying every kernel ~40x stays cheap.  Usage: python tools/profile_step.py [L] [precision] [dataset] [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from defensegan_b200.models.gan import dataset_gan_dict

L = int(sys.argv[1]) if len(sys.argv) > 1 else 3
precision = sys.argv[2] if len(sys.argv) > 2 else "fp16"
dataset = sys.argv[3] if len(sys.argv) > 3 else "mnist"
B = int(sys.argv[4]) if len(sys.argv) > 4 else 256
R = 10
gan = dataset_gan_dict[dataset](test_mode=True, verbose=False, precision=precision, batch_size=50)
gan.rec_rr, gan.rec_iters = R, L
g = torch.Generator().manual_seed(0)
x = torch.rand(B, *gan.image_dim, generator=g).cuda()
z0 = (torch.randn(B * R, 128, generator=g) * 128 ** -0.5).cuda()
rec = gan.reconstruct(x, z_init_val=z0)
torch.cuda.synchronize()
print("done", float(rec.mean()), "launches", gan._native.last_launch_count)
