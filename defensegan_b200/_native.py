"""ctypes binding of the C-ABI in include/defensegan_b200.h.

PyTorch is used only for device memory and streams.  There is no CPU fallback: if the shared
library is missing or no sm_100 GPU is present, every compute entry point raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys
from typing import Dict, List, Optional, Sequence

import torch

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_NAME = "libdefensegan_b200.so"
LIB_PATH = os.environ.get("DGAN_LIB", os.path.join(_PKG_DIR, LIB_NAME))   # DGAN_LIB: A/B-test another build
CSRC_DIR = os.path.join(_PKG_DIR, "csrc")
INCLUDE_DIR = os.path.join(os.path.dirname(_PKG_DIR), "include")

ARCH_IDS = {"mnist": 0, "f-mnist": 0, "fmnist": 0, "celeba": 1}
PRECISIONS = {"fp32": 0, "fp16": 1}

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]

# Every symbol include/defensegan_b200.h declares.
ABI_SYMBOLS = [
    "dgan_abi_version", "dgan_last_error", "dgan_num_weights", "dgan_create", "dgan_destroy",
    "dgan_workspace_bytes", "dgan_reconstruct", "dgan_sample_z0", "dgan_forward", "dgan_loss_grad",
    "dgan_last_launch_count", "dgan_last_enqueue_count", "dgan_macs_per_row", "dgan_profile_enable", "dgan_profile_num_kinds",
    "dgan_profile_kind_name", "dgan_profile_read",
]


class dgan_desc(ctypes.Structure):
    _fields_ = [("abi_version", ctypes.c_int32), ("arch", ctypes.c_int32), ("latent_dim", ctypes.c_int32),
                ("net_dim", ctypes.c_int32), ("use_bn", ctypes.c_int32), ("precision", ctypes.c_int32)]


class dgan_rec_params(ctypes.Structure):
    _fields_ = [("batch", ctypes.c_int32), ("rec_rr", ctypes.c_int32), ("rec_iters", ctypes.c_int32),
                ("rec_lr", ctypes.c_float), ("momentum", ctypes.c_float), ("decay_lr", ctypes.c_int32),
                ("seed", ctypes.c_uint64), ("z_row_offset", ctypes.c_uint64)]


ABI_VERSION = 2


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/ for sm_100a with nvcc into the in-tree shared library (cross-compiles
    without a GPU).  Rebuilds when any source is newer than the library."""
    srcs = [os.path.join(CSRC_DIR, f) for f in sorted(os.listdir(CSRC_DIR))]
    srcs.append(os.path.join(INCLUDE_DIR, "defensegan_b200.h"))
    if not force and os.path.exists(LIB_PATH):
        lib_m = os.path.getmtime(LIB_PATH)
        if all(os.path.getmtime(s) <= lib_m for s in srcs):
            return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + [os.path.join(CSRC_DIR, "dgan_api.cu"), "-o", LIB_PATH]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout)
    return LIB_PATH


PROBE_LIB_PATH = os.path.join(_PKG_DIR, "libdefensegan_b200_probe.so")


def build_probe_library(force: bool = False) -> str:
    """The same sources with -DDGAN_PROBE: per-CTA clock / %globaltimer counters in the tensor-core kernels and
    dgan_debug_probe_read().  Measurement aid only (tools/probe_step.py, the `timeline` pass of bench.py run it in a
    process of its own through DGAN_LIB); the product library carries none of it."""
    srcs = [os.path.join(CSRC_DIR, f) for f in sorted(os.listdir(CSRC_DIR))]
    if not force and os.path.exists(PROBE_LIB_PATH) and all(os.path.getmtime(s) <= os.path.getmtime(PROBE_LIB_PATH) for s in srcs):
        return PROBE_LIB_PATH
    cmd = [os.environ.get("NVCC", "nvcc")] + NVCC_FLAGS + ["-DDGAN_PROBE", os.path.join(CSRC_DIR, "dgan_api.cu"), "-o", PROBE_LIB_PATH]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout)
    return PROBE_LIB_PATH


_lib = None


def load_library() -> ctypes.CDLL:
    """dlopen the in-tree library and declare the signatures of the header."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "%s not found: run `python -c 'import __graft_entry__ as g; g.build()'` (or "
            "defensegan_b200._native.build_library()) first. There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    lib.dgan_abi_version.restype = ctypes.c_int
    if lib.dgan_abi_version() != ABI_VERSION:
        raise RuntimeError("%s has ABI version %d, this binding needs %d: rebuild it (__graft_entry__.build())"
                           % (LIB_PATH, lib.dgan_abi_version(), ABI_VERSION))
    vp, i32, u64, f32, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_float, ctypes.c_size_t
    lib.dgan_abi_version.restype = i32
    lib.dgan_abi_version.argtypes = []
    lib.dgan_last_error.restype = ctypes.c_char_p
    lib.dgan_last_error.argtypes = []
    lib.dgan_num_weights.restype = i32
    lib.dgan_num_weights.argtypes = [ctypes.POINTER(dgan_desc)]
    lib.dgan_create.restype = i32
    lib.dgan_create.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(dgan_desc), ctypes.POINTER(vp), i32, vp]
    lib.dgan_destroy.restype = i32
    lib.dgan_destroy.argtypes = [vp]
    lib.dgan_workspace_bytes.restype = sz
    lib.dgan_workspace_bytes.argtypes = [vp, i32, i32]
    lib.dgan_reconstruct.restype = i32
    lib.dgan_reconstruct.argtypes = [vp, ctypes.POINTER(dgan_rec_params), vp, vp, vp, vp, vp, vp, sz, vp]
    lib.dgan_sample_z0.restype = i32
    lib.dgan_sample_z0.argtypes = [vp, u64, u64, i32, vp, vp]
    lib.dgan_forward.restype = i32
    lib.dgan_forward.argtypes = [vp, vp, i32, vp, vp, sz, vp]
    lib.dgan_loss_grad.restype = i32
    lib.dgan_loss_grad.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp, vp, sz, vp]
    lib.dgan_last_launch_count.restype = ctypes.c_int64
    lib.dgan_last_launch_count.argtypes = [vp]
    lib.dgan_last_enqueue_count.restype = ctypes.c_int64
    lib.dgan_last_enqueue_count.argtypes = [vp]
    lib.dgan_macs_per_row.restype = ctypes.c_int64
    lib.dgan_macs_per_row.argtypes = [vp]
    lib.dgan_profile_enable.restype = i32
    lib.dgan_profile_enable.argtypes = [vp, i32]
    lib.dgan_profile_num_kinds.restype = i32
    lib.dgan_profile_num_kinds.argtypes = [vp]
    lib.dgan_profile_kind_name.restype = ctypes.c_char_p
    lib.dgan_profile_kind_name.argtypes = [vp, i32]
    lib.dgan_profile_read.restype = i32
    lib.dgan_profile_read.argtypes = [vp, i32, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64),
                                      ctypes.POINTER(ctypes.c_double)]
    _lib = lib
    return lib


def _check(lib, rc: int, what: str):
    if rc != 0:
        msg = lib.dgan_last_error()
        raise RuntimeError("%s failed (status %d): %s" % (what, rc, msg.decode() if msg else "?"))


def _ptr(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _require_cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor (there is no CPU path)" % name)
    if t.dtype != torch.float32:
        t = t.to(torch.float32)
    return t.contiguous()


class NativeGenerator:
    """Owns one dgan_handle (the generator's re-laid-out weights on one GPU)."""

    def __init__(self, arch: str, weights: Sequence[torch.Tensor], latent_dim: int = 128, net_dim: int = 64,
                 use_bn: bool = False, precision: str = "fp32", device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("defensegan_b200 needs a CUDA (sm_100) device; there is no CPU fallback")
        if arch not in ARCH_IDS:
            raise ValueError("unknown arch %r" % (arch,))
        if precision not in PRECISIONS:
            raise ValueError("precision must be one of %s" % sorted(PRECISIONS))
        self.lib = load_library()
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.arch, self.precision = arch, precision
        self.latent_dim, self.net_dim = int(latent_dim), int(net_dim)
        self.image_dim = (64, 64, 3) if ARCH_IDS[arch] == 1 else (28, 28, 1)
        self.hwc = self.image_dim[0] * self.image_dim[1] * self.image_dim[2]
        self._handle = ctypes.c_void_p(0)
        self._ws = None
        desc = dgan_desc(ABI_VERSION, ARCH_IDS[arch], self.latent_dim, self.net_dim, int(bool(use_bn)), PRECISIONS[precision])
        with torch.cuda.device(self.device):
            # the handle copies the weights (on `stream`): they only have to outlive those copies
            ws = [_require_cuda_f32(w.to(self.device), "weight") for w in weights]
            n_expected = self.lib.dgan_num_weights(ctypes.byref(desc))
            if len(ws) != n_expected:
                raise ValueError("expected %d weight tensors, got %d" % (n_expected, len(ws)))
            arr = (ctypes.c_void_p * len(ws))(*[w.data_ptr() for w in ws])
            stream = torch.cuda.current_stream(self.device)
            h = ctypes.c_void_p(0)
            _check(self.lib, self.lib.dgan_create(ctypes.byref(h), ctypes.byref(desc), arr, len(ws),
                                                  ctypes.c_void_p(stream.cuda_stream)), "dgan_create")
            self._handle = h
            stream.synchronize()
            del ws

    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle.value:
            torch.cuda.synchronize(self.device)
            self.lib.dgan_destroy(self._handle)
            self._handle = ctypes.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers -------------------------------------------------------------------------
    def _workspace(self, batch: int, rec_rr: int):
        need = int(self.lib.dgan_workspace_bytes(self._handle, batch, rec_rr))
        if need == 0:
            raise RuntimeError("dgan_workspace_bytes returned 0 (invalid batch / rec_rr)")
        if self._ws is None or self._ws.numel() < need + 1024:
            if self._ws is not None:
                # earlier calls (possibly on other streams) may still use the old block: let them finish before the
                # caching allocator can hand it to somebody else
                torch.cuda.synchronize(self.device)
            self._ws = None
            self._ws = torch.empty(need + 1024, dtype=torch.uint8, device=self.device)
        base = self._ws.data_ptr()
        aligned = (base + 1023) // 1024 * 1024
        return ctypes.c_void_p(aligned), need

    @property
    def macs_per_row(self) -> int:
        return int(self.lib.dgan_macs_per_row(self._handle))

    @property
    def last_launch_count(self) -> int:
        return int(self.lib.dgan_last_launch_count(self._handle))

    @property
    def last_enqueue_count(self) -> int:
        """Stream operations the host issued for the last reconstruct (the L-step loop is one CUDA-graph launch)."""
        return int(self.lib.dgan_last_enqueue_count(self._handle))

    def profile_enable(self, on: bool) -> None:
        _check(self.lib, self.lib.dgan_profile_enable(self._handle, int(bool(on))), "dgan_profile_enable")

    def profile_read(self):
        """[{name, ms, launches, flops_per_launch}] for the launches recorded since profile_enable(True)."""
        nk = int(self.lib.dgan_profile_num_kinds(self._handle))
        ms = (ctypes.c_double * nk)()
        cnt = (ctypes.c_int64 * nk)()
        fl = (ctypes.c_double * nk)()
        _check(self.lib, self.lib.dgan_profile_read(self._handle, nk, ms, cnt, fl), "dgan_profile_read")
        return [dict(name=self.lib.dgan_profile_kind_name(self._handle, k).decode(), ms=float(ms[k]),
                     launches=int(cnt[k]), flops_per_launch=float(fl[k])) for k in range(nk)]

    # -- entry points ----------------------------------------------------------------------
    def reconstruct(self, images: torch.Tensor, rec_rr: int, rec_iters: int, rec_lr: float = 10.0,
                    z_init_val: Optional[torch.Tensor] = None, seed: int = 0, momentum: float = 0.7,
                    decay_lr: bool = False, out: Optional[torch.Tensor] = None, return_aux: bool = False,
                    z_row_offset: int = 0):
        x = _require_cuda_f32(images, "images")
        batch = x.shape[0]
        if x.numel() != batch * self.hwc:
            raise ValueError("images must be [B,%d,%d,%d]" % self.image_dim)
        if rec_rr <= 0 or rec_iters <= 0 or batch <= 0:
            raise ValueError("batch, rec_rr and rec_iters must be positive")
        z0 = None
        if z_init_val is not None:
            z0 = _require_cuda_f32(z_init_val, "z_init_val")
            if z0.numel() != batch * rec_rr * self.latent_dim:
                raise ValueError("z_init_val must be [B*rec_rr, latent_dim]")
        with torch.cuda.device(self.device):
            rec = out if out is not None else torch.empty_like(x)
            if not (rec.is_cuda and rec.dtype == torch.float32 and rec.is_contiguous() and rec.numel() == x.numel()):
                raise ValueError("out must be a contiguous CUDA float32 tensor shaped like images")
            loss = torch.empty(batch, dtype=torch.float32, device=self.device)
            idx = torch.empty(batch, dtype=torch.int32, device=self.device)
            ws, need = self._workspace(batch, rec_rr)
            stream = torch.cuda.current_stream(self.device).cuda_stream
            prm = dgan_rec_params(batch, int(rec_rr), int(rec_iters), float(rec_lr), float(momentum), int(bool(decay_lr)),
                                  seed & (2 ** 64 - 1), int(z_row_offset))
            rc = self.lib.dgan_reconstruct(self._handle, ctypes.byref(prm), _ptr(x), _ptr(z0), _ptr(rec), _ptr(loss),
                                           _ptr(idx), ws, need, ctypes.c_void_p(stream))
            _check(self.lib, rc, "dgan_reconstruct")
        rec = rec.view(images.shape) if out is None else rec
        if return_aux:
            return rec, loss, idx
        return rec

    def sample_z0(self, n_rows: int, seed: int, z_row_offset: int = 0) -> torch.Tensor:
        """Rows [z_row_offset, z_row_offset + n_rows) of the N(0, 1/latent_dim) Philox stream `seed` - the z0 that
        reconstruct(..., z_init_val=None, seed=seed) starts from (reference models/gan.py:370-377)."""
        with torch.cuda.device(self.device):
            z = torch.empty(int(n_rows), self.latent_dim, dtype=torch.float32, device=self.device)
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _check(self.lib, self.lib.dgan_sample_z0(self._handle, ctypes.c_uint64(seed & (2 ** 64 - 1)),
                                                     ctypes.c_uint64(int(z_row_offset)), int(n_rows), _ptr(z),
                                                     ctypes.c_void_p(stream)), "dgan_sample_z0")
        return z

    def forward(self, z: torch.Tensor) -> torch.Tensor:
        zc = _require_cuda_f32(z, "z")
        n = zc.shape[0]
        with torch.cuda.device(self.device):
            y = torch.empty((n,) + self.image_dim, dtype=torch.float32, device=self.device)
            ws, need = self._workspace(n, 1)
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _check(self.lib, self.lib.dgan_forward(self._handle, _ptr(zc), n, _ptr(y), ws, need, ctypes.c_void_p(stream)),
                   "dgan_forward")
        return y

    def loss_grad(self, images: torch.Tensor, z: torch.Tensor, rec_rr: int):
        x = _require_cuda_f32(images, "images")
        zc = _require_cuda_f32(z, "z")
        batch = x.shape[0]
        n = batch * rec_rr
        if zc.shape[0] != n:
            raise ValueError("z must have batch*rec_rr rows")
        with torch.cuda.device(self.device):
            y = torch.empty((n,) + self.image_dim, dtype=torch.float32, device=self.device)
            loss = torch.empty(n, dtype=torch.float32, device=self.device)
            grad = torch.empty(n, self.latent_dim, dtype=torch.float32, device=self.device)
            ws, need = self._workspace(batch, rec_rr)
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _check(self.lib, self.lib.dgan_loss_grad(self._handle, _ptr(x), batch, rec_rr, _ptr(zc), _ptr(y), _ptr(loss),
                                                     _ptr(grad), ws, need, ctypes.c_void_p(stream)), "dgan_loss_grad")
        return y, loss, grad
