"""Trace of one steady-state L-step of the persistent loop kernel (developer aid): for every item when its producer started
waiting for its dependencies, when the wait ended, and when its epilogue began / ended; prints where the dependency
waits are and which producing item each long wait was for.   Usage: python tools/loop_trace.py [dataset] [B] [L] [out.npz]"""
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
keep = a[:, 7] > 0                      # items of the two traced program entries (sections 1 and 2)
idx_of = -np.ones(n, dtype=np.int64)
idx_of[np.nonzero(keep)[0]] = np.arange(keep.sum())
a = a[keep]
deps = deps[keep]
deps = np.where(deps >= 0, idx_of[np.clip(deps, 0, n - 1)], deps)
n = len(a)
t0 = a[:, 4][a[:, 4] > 0].min()
pair, seg, win, mp = a[:, 0], a[:, 1], a[:, 2], a[:, 3]
wb, we, eb, ee = [(a[:, 4 + k] - t0) / 1e3 for k in range(4)]           # us
wait = we - wb
print("items %d, traced step spans %.1f us (first wait begin -> last epilogue end)" % (n, ee.max()))
print("dependency wait: total %.1f us over %d CTA pairs = %.1f us per pair; items waiting > 1 us: %d" % (
    wait.sum(), len(set(pair)), wait.sum() / len(set(pair)), (wait > 1).sum()))
for s in sorted(set(seg)):
    m = seg == s
    nph = len(names) - 1
    print("  vseg %2d %s.%-26s items %4d  wait/pair %6.1f us  begin %7.1f..%7.1f  epilogue end %7.1f..%7.1f" % (
        s, "AB"[s // nph], names[s % nph][:26], m.sum(), wait[m].sum() / len(set(pair)), wb[m].min(), wb[m].max(), ee[m].min(), ee[m].max()))
order = np.argsort(-wait)[:25]
print("longest waits: item (pair seg win mp) waited us | released by dep item (pair seg win mp) whose epilogue ended at, flag seen at")
for e in order:
    d = deps[e][deps[e] >= 0]
    if len(d) == 0:
        print("  %5d (%2d %d %3d %2d) %6.1f us | z update of the previous L-step" % (e, pair[e], seg[e], win[e], mp[e], wait[e]))
        continue
    last = d[np.argmax(ee[d])]
    print("  %5d (%2d %d %3d %2d) %6.1f us [%.1f -> %.1f] | %5d (%2d %d %3d %2d) epilogue %.1f..%.1f" % (
        e, pair[e], seg[e], win[e], mp[e], wait[e], wb[e], we[e], last, pair[last], seg[last], win[last], mp[last], eb[last], ee[last]))
# per pair timeline summary
busy = np.zeros(int(pair.max()) + 1)
for p in range(len(busy)):
    m = pair == p
    busy[p] = wait[m].sum()
print("per-pair dependency wait: min %.1f median %.1f max %.1f us" % (busy.min(), np.median(busy), busy.max()))
