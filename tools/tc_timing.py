"""Developer aid: per-role wait/work cycles of the CTA-pair tensor-core kernels (DGAN_TC_DEBUG=1).
Prints, for each of the first launches of one projection, the mean over CTAs of: producer wait on ring space,
MMA wait on 'full', MMA wait on 'acc_empty', MMA issue, whole MMA loop, epilogue work, CTA utilisation, the
%globaltimer timeline of the launch, and (leader CTAs) when the MMA loop started / first operands landed / loop ended.
DGAN_TC_DBGFLAGS: 8 = no per-step clocks (unperturbed loop time), 16 = issue no MMAs (delivery-only time),
1 = skip epilogue stores, 2 = skip mask loads, 32 = skip last-layer target loads."""
import ctypes
import os
import sys

os.environ["DGAN_TC_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from defensegan_b200.models.gan import dataset_gan_dict

dataset = sys.argv[1] if len(sys.argv) > 1 else "mnist"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
R, L = 10, 3
gan = dataset_gan_dict[dataset](test_mode=True, verbose=False, precision="fp16", batch_size=50)
gan.rec_rr, gan.rec_iters = R, L
g = torch.Generator().manual_seed(0)
x = torch.rand(B, *gan.image_dim, generator=g).cuda()
z0 = (torch.randn(B * R, 128, generator=g) * 128 ** -0.5).cuda()
gan.reconstruct(x, z_init_val=z0)
torch.cuda.synchronize()
nat = gan._native
buf = (ctypes.c_ulonglong * (64 * 160 * 16))()
nat.lib.dgan_debug_tc_timing.restype = ctypes.c_int
nat.lib.dgan_debug_tc_timing.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
n = nat.lib.dgan_debug_tc_timing(nat._handle, buf, 64)
a = np.frombuffer(buf, dtype=np.uint64).reshape(64, 160, 16)[:n].astype(np.float64)
names = ["prod_wait_empty", "mma_wait_full", "mma_wait_acc", "mma_issue", "mma_loop_total", "epi_work", "CTA util %"]
print("launch | " + " | ".join(names) + " | first start us | last start us | first end us | last end us | gap to next us"
      "   (cycles: mean over active CTAs; MMA columns: leader CTAs only; times: %globaltimer, PDL on)")
raw = np.frombuffer(buf, dtype=np.uint64).reshape(64, 160, 16)[:n]
t0 = None
rows = []
extras = {}
for i in range(n):
    act = raw[i][:, 7] > 0
    if not act.any():
        continue
    lead = act & (a[i][:, 4] > 0)
    vals = []
    for k in range(6):
        m = lead if k in (1, 2, 3, 4) else act
        vals.append(a[i][m, k].mean() if m.any() else 0.0)
    vals[4] = a[i][lead, 4].mean() if lead.any() else 0.0
    st, en = raw[i][act, 6].astype(np.int64), raw[i][act, 7].astype(np.int64)
    if t0 is None:
        t0 = st.min()
    util = float((en - st).mean()) / max(1.0, float(en.max() - st.min()))
    vals.append(100.0 * util)
    lm = lead & (raw[i][:, 8] > 0)
    if lm.any():   # leader CTAs: us from the PDL wait to MMA-loop start / first operands landed / loop end / CTA end
        g0 = raw[i][lm, 6].astype(np.int64)
        extra = "  [mma start +%.1f, first full +%.1f, loop end +%.1f, cta end +%.1f us]" % tuple(
            float((raw[i][lm, k].astype(np.int64) - g0).mean()) / 1e3 for k in (8, 10, 9, 7))
    else:
        extra = ""
    extras[i] = extra
    rows.append((i, vals, (st.min() - t0) / 1e3, (st.max() - t0) / 1e3, (en.min() - t0) / 1e3, (en.max() - t0) / 1e3))
for j, (i, vals, s0, s1, e0, e1) in enumerate(rows):
    gap = rows[j + 1][2] - e1 if j + 1 < len(rows) else float("nan")
    print("%2d | " % i + " | ".join("%9.0f" % v for v in vals) + " | %8.1f | %8.1f | %8.1f | %8.1f | %6.1f" % (s0, s1, e0, e1, gap) + extras.get(i, ""))
