"""Regenerates the golden fixtures in this directory from the CPU oracle (oracle/defensegan_oracle.py).

    python tests/golden/make_golden.py

The reference (TF1/py2) cannot run here and ships no fixtures for this path, so
these vectors pin the *oracle restatement* (fp32 run = reference-precision stand-in, fp64 run =
truth used to bound drift); parity stays "unpinned" with respect to TensorFlow itself.
Weights are not stored: they are re-drawn from seed by oracle.init_generator_weights and
checked against the stored SHA-256.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import defensegan_oracle as O  # noqa: E402


def weights_digest(w):
    h = hashlib.sha256()
    for k, v in w.items():
        h.update(k.encode())
        h.update(np.ascontiguousarray(v).tobytes())
    return h.hexdigest()


def make_case(name, arch, B, R, L, kind="S1", lr=10.0, random_bias=False, seed_img=O.IMAGE_SEED, seed_z=O.Z0_SEED,
              use_bn=False):
    w = O.init_generator_weights(arch, random_bias=random_bias, use_bn=use_bn)
    imgs = O.synthetic_images(arch, w, B, kind=kind, seed=seed_img)   # S1 images come from the plain (no-BN) forward
    z0 = O.sample_z0(B * R, 128, seed=seed_z)
    r32 = O.reconstruct(arch, w, imgs, R, L, rec_lr=lr, z_init_val=z0, dtype=torch.float32, use_bn=use_bn)
    r64 = O.reconstruct(arch, w, imgs, R, L, rec_lr=lr, z_init_val=z0, dtype=torch.float64, use_bn=use_bn)
    y, loss, grad = O.loss_and_grad(arch, w, imgs, z0, R, dtype=torch.float64, use_bn=use_bn)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        arch=arch, B=B, R=R, L=L, lr=lr, random_bias=int(random_bias), use_bn=int(use_bn), weights_sha256=weights_digest(w),
        images=imgs, z0=z0,
        rec32=r32["rec"], loss_min32=r32["loss_min"], idx32=r32["idx"], loss_all32=r32["loss_all"],
        rec64=r64["rec"].astype("float32"), loss_min64=r64["loss_min"], idx64=r64["idx"], loss_all64=r64["loss_all"],
        y0_64=y.astype("float32"), loss0_64=loss, grad0_64=grad)
    print(name, "loss_min32", r32["loss_min"][:4], "idx", r32["idx"][:8],
          "max|rec32-rec64|", float(np.abs(r32["rec"] - r64["rec"]).max()))


if __name__ == "__main__":
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    make_case("mnist_c1", "mnist", 16, 2, 10)                       # BASELINE configs[0]
    make_case("mnist_ragged_bias", "mnist", 5, 3, 6, kind="S2", random_bias=True)
    make_case("celeba_small", "celeba", 3, 2, 4, random_bias=True)
    # opt-in batch-statistics BatchNorm (use_bn=True): all rows of a call are coupled
    make_case("mnist_bn", "mnist", 8, 2, 3, random_bias=True, use_bn=True, lr=0.5)
    make_case("celeba_bn", "celeba", 2, 2, 3, random_bias=True, use_bn=True, lr=0.5)
