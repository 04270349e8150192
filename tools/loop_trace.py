"""Trace of one steady-state L-step of the persistent loop kernel (developer aid): for every item when a CTA pair took it
from the ready queue, when its accumulator buffer was granted, and when its epilogue began / ended; prints per segment
the ready -> taken -> done latencies and, per CTA pair, how much of the step it spent without an item in flight.
Usage: python tools/loop_trace.py [dataset] [B] [L] [out.npz]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from defensegan_b200.models.gan import dataset_gan_dict

dataset = sys.argv[1] if len(sys.argv) > 1 else "mnist"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
L = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out_path = sys.argv[4] if len(sys.argv) > 4 else None
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
names = [k["name"] for k in nat.profile_read()]
lib = nat.lib
MAXI, MAXD = 1 << 16, 48
lib.dgan_debug_loop_trace.restype = ctypes.c_int
lib.dgan_debug_loop_trace.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_int64), ctypes.c_int, ctypes.c_int]
buf = (ctypes.c_uint64 * (8 * MAXI))()
dbuf = (ctypes.c_int64 * (MAXD * MAXI))()
n = lib.dgan_debug_loop_trace(nat._handle, buf, dbuf, MAXD, MAXI)
a = np.frombuffer(buf, dtype=np.uint64).reshape(MAXI, 8)[:n].astype(np.int64)
deps = np.frombuffer(dbuf, dtype=np.int64).reshape(MAXI, MAXD)[:n]
if out_path:
    np.savez_compressed(out_path, items=a, deps=deps, names=np.array(names))
keep = a[:, 7] > 0                      # items of the traced L-step
idx_of = -np.ones(n, dtype=np.int64)
idx_of[np.nonzero(keep)[0]] = np.arange(keep.sum())
a = a[keep]
deps = deps[keep]
deps = np.where(deps >= 0, idx_of[np.clip(deps, 0, n - 1)], deps)
n = len(a)
t0 = a[:, 4][a[:, 4] > 0].min() & ((1 << 48) - 1)
pair, seg, win, mp = a[:, 0], a[:, 1], a[:, 2], a[:, 3]
pop, acc, eb, ee = [((a[:, 4 + k] & ((1 << 48) - 1)) - t0) / 1e3 for k in range(4)]           # us
# ready time of an item = latest epilogue end among the items it waits for (store completion and queue hops come on top)
ready = np.full(n, np.nan)
for e in range(n):
    d = deps[e][deps[e] >= 0]
    if len(d):
        ready[e] = ee[d].max()
lat = pop - ready
print("items %d on %d CTA pairs, traced step spans %.1f us (first pop -> last epilogue end)" % (n, len(set(pair)), ee.max()))
print("%-28s %5s %9s %9s %9s %9s %9s" % ("segment", "items", "pop-ready", "acc-pop", "epi-acc", "epi", "pop..end"))
for sg in sorted(set(seg)):
    m = seg == sg
    print("%-28s %5d %9.1f %9.1f %9.1f %9.1f %9.1f   pops %7.1f..%7.1f  ends %7.1f..%7.1f" % (
        names[sg][:28], m.sum(), np.nanmean(lat[m]) if np.isfinite(lat[m]).any() else float("nan"), (acc[m] - pop[m]).mean(),
        (eb[m] - acc[m]).mean(), (ee[m] - eb[m]).mean(), (ee[m] - pop[m]).mean(), pop[m].min(), pop[m].max(), ee[m].min(), ee[m].max()))
# per CTA pair: the union of [pop, epilogue end] intervals vs the span of the traced step
span = ee.max() - pop.min()
idle = []
for p in sorted(set(pair)):
    m = np.nonzero(pair == p)[0]
    iv = sorted(zip(pop[m], ee[m]))
    covered, cur_b, cur_e = 0.0, iv[0][0], iv[0][1]
    for b_, e_ in iv[1:]:
        if b_ > cur_e:
            covered += cur_e - cur_b
            cur_b, cur_e = b_, e_
        else:
            cur_e = max(cur_e, e_)
    covered += cur_e - cur_b
    idle.append(span - covered)
idle = np.array(idle)
print("time without an item in flight per CTA pair: min %.1f median %.1f max %.1f us of %.1f us" % (idle.min(), np.median(idle), idle.max(), span))
cnt = np.bincount(pair.astype(np.int64))
print("items per CTA pair: min %d median %d max %d" % (cnt[cnt > 0].min(), np.median(cnt[cnt > 0]), cnt.max()))
for m_ in sorted(set(mp)):
    m = mp == m_
    print("  row pair %2d: step runs %7.1f .. %7.1f us" % (m_, pop[m].min(), ee[m].max()))
