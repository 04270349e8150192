"""Where the roles of the persistent loop kernel spend their time (developer aid).
Runs one projection under profile level 1 and prints, per role, the share of its loop spent waiting:
producer (ready queue empty / ring space), MMA issuer (mailbox / operands / accumulator buffers), epilogue (accumulators /
staging tile), store warp (tiles / store completion).   Usage: python tools/loop_stalls.py [dataset] [B] [L]"""
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
R = 10
gan = dataset_gan_dict[dataset](test_mode=True, verbose=False, precision="fp16", batch_size=50)
gan.rec_rr, gan.rec_iters = R, L
g = torch.Generator().manual_seed(0)
x = torch.rand(B, *gan.image_dim, generator=g).cuda()
z0 = (torch.randn(B * R, 128, generator=g) * 128 ** -0.5).cuda()
gan.reconstruct(x, z_init_val=z0)
nat = gan._native
nat.profile_enable(1)
gan.reconstruct(x, z_init_val=z0)
torch.cuda.synchronize()
prof = nat.profile_read()
lib = nat.lib
lib.dgan_debug_loop_stalls.restype = ctypes.c_int
lib.dgan_debug_loop_stalls.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64), ctypes.c_int]
buf = (ctypes.c_uint64 * (16 * 512))()
n = lib.dgan_debug_loop_stalls(nat._handle, buf, 512)
a = np.frombuffer(buf, dtype=np.uint64).reshape(512, 16)[:n].astype(np.float64)
names = ["P_POP", "P_RING", "P_TOTAL", "M_FULL", "M_ACC", "M_TOTAL", "E_ACC", "E_TILE", "E_TOTAL", "S_TILE", "S_DONE", "S_TOTAL", "P_ITEMS", "M_MAIL", "F_QUEUE", "F_CREDIT"]
lead = a[0::2]
print("CTAs", n, "L", L, "status", nat.last_status())
for k in prof:
    if k["launches"]:
        print("  %-60s %9.1f us x %d" % (k["name"], 1e3 * k["ms"] / k["launches"], k["launches"]))
tot = a[:, 2].mean()
print("ticks per L-step (producer loop): %.0f" % (tot / L))
for grp, keys, total in (("producer (leader CTAs)", (0, 1), 2), ("MMA (leader CTAs)", (13, 3, 4), 5), ("epilogue", (6, 7), 8), ("store warp", (9, 10), 11)):
    src = lead if "leader" in grp else a
    t = src[:, total].mean()
    print("%-18s total %10.0f ticks  " % (grp, t) + "  ".join("%s %5.1f%% (max %5.1f%%)" % (names[k], 100 * src[:, k].mean() / t, 100 * (src[:, k] / src[:, total]).max()) for k in keys))
peer = a[1::2]
print("fetcher (peer CTAs): waiting for a ready item %5.1f%%, for credit / a free mailbox slot %5.1f%% of the kernel" % (
    100 * peer[:, 14].mean() / lead[:, 2].mean(), 100 * peer[:, 15].mean() / lead[:, 2].mean()))
items = lead[:, 12]
print("items per CTA pair: mean %.0f  min %.0f  max %.0f   (per L-step: %.1f)" % (items.mean(), items.min(), items.max(), items.mean() / L))
