"""How long does the host take to enqueue one projection (no sync) vs. how long the GPU takes to run it?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from defensegan_b200.models.gan import dataset_gan_dict
B, R, L = 256, 10, 200
gan = dataset_gan_dict["mnist"](test_mode=True, verbose=False, precision="fp16", batch_size=50)
gan.rec_rr, gan.rec_iters = R, L
g = torch.Generator().manual_seed(0)
x = torch.rand(B, 28, 28, 1, generator=g).cuda()
z0 = (torch.randn(B * R, 128, generator=g) * 128 ** -0.5).cuda()
for _ in range(2):
    gan.reconstruct(x, z_init_val=z0)
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter()
    gan.reconstruct(x, z_init_val=z0)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("chains=%s enqueue %.1f ms, total %.1f ms, launches %d -> %.2f us/launch host" % (
        os.environ.get("DGAN_CHAINS", "default"), 1e3 * (t1 - t0), 1e3 * (t2 - t0), gan._native.last_launch_count,
        1e6 * (t1 - t0) / gan._native.last_launch_count))
