"""Host-side mirror of the reference's `models/gan.py` model classes for the projection loop.

Same class names, attribute names, defaults, argument meaning and error behaviour as
`DefenseGANBase` and its dataset subclasses (reference models/gan.py:39-135,333-449,649-765),
with the TF1 graph machinery replaced by eager calls into the native sm_100a library:

    gan = MnistDefenseGAN(cfg=cfg, test_mode=True)
    gan.load_generator()                      # reference models/gan.py:86-87
    gan.rec_rr, gan.rec_lr, gan.rec_iters = 10, 10.0, 200   # callers set these (blackbox.py:649-658)
    x_hat = gan.reconstruct(images)           # torch CUDA tensor [B,H,W,C] in, same shape out

Graph-mode -> eager mapping: `images` is a float32 CUDA tensor (NHWC,
already input-transformed) instead of a symbolic placeholder; the result is a detached tensor
(the reference graph has no gradient path from `images` to the output either);
`batch_size`, `back_prop` and `reconstructor_id` are accepted for signature compatibility -
any batch size works (superset of the reference's static-shape graph).
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import numpy as np
import torch

from .. import _native
from .. import weights as _weights
from ..utils.config import load_config, packaged_cfg_path

__all__ = ["DefenseGANBase", "MnistDefenseGAN", "FmnistDefenseDefenseGAN", "CelebADefenseGAN", "dataset_gan_dict"]


class DefenseGANBase(object):
    """Holds the hyper-parameters, binds the generator and exposes reconstruct()."""

    _dataset_default = None
    _image_dim_default = [None, None, None]

    def __init__(self, cfg=None, test_mode=False, verbose=True, **args):
        # defaults of reference models/gan.py:50-69 (those the projection loop reads)
        self.dataset_name = self._dataset_default
        self.batch_size = 32
        self.use_bn = True                 # class default; every shipped cfg sets USE_BN False
        self.test_batch_size = 20
        self.mode = "gp-wgan"
        self.latent_dim = None
        self.net_dim = None
        self.input_transform_type = 0
        self.debug = False
        self.rec_iters = 200
        self.image_dim = list(self._image_dim_default)
        self.rec_rr = 10
        self.rec_lr = 10.0
        self.test_again = False
        self.attribute = "gender"
        self.output_dir = "output"
        # additions of this implementation
        self.precision = "fp32"            # 'fp32' (CUDA-core, reference arithmetic) | 'fp16' (tcgen05 operands)
        self.rec_momentum = 0.7            # tf.train.MomentumOptimizer(momentum=0.7), models/gan.py:389-391
        self.rec_decay_lr = False          # the reference's decay is dead code; True = intended schedule
        self.seed = 11241990               # callers use tf.set_random_seed(11241990) (blackbox.py:464)

        self.test_mode = test_mode
        self.verbose = verbose
        self.is_training = not test_mode
        self.initialized = False
        self._set_attr(cfg, args)
        if self.latent_dim is None:
            self.latent_dim = 128
        if self.net_dim is None:
            self.net_dim = 64
        self._set_checkpoint_dir()
        self._build()
        self._native = None
        self._native_key = None
        self._call_counter = 0
        # TF creates the variables with their random initial values at graph construction;
        # load_generator() later overwrites them from a checkpoint.
        self.weights = _weights.init_generator_weights(self.arch, seed=self.seed, latent_dim=self.latent_dim,
                                                       net_dim=self.net_dim, use_bn=bool(self.use_bn))

    # -- configuration (reference models/base_model.py:120-148) ------------------------------
    def _set_attr(self, cfg, args):
        if cfg is None:
            # no cfg given: the packaged copy of the dataset's yml over default.yml
            ds = args.get("dataset_name", self.dataset_name)
            cfg = load_config(packaged_cfg_path(ds)) if ds is not None else None
        elif isinstance(cfg, str):
            cfg = load_config(cfg)
        self.cfg = cfg
        known = [k for k in self.__dict__.keys() if not k.startswith("_")]
        for attr in known:
            val = None
            if cfg is not None:
                if attr.upper() in cfg:
                    val = cfg[attr.upper()]
                elif attr in cfg:
                    val = cfg[attr]
            if attr in args:
                val = args[attr]
            if val is not None:
                setattr(self, attr, val)
        unknown = [k for k in args if k not in known]
        if unknown:
            raise TypeError("unexpected keyword argument(s): %s" % ", ".join(sorted(unknown)))
        if self.image_dim is not None:
            self.image_dim = [int(v) if v is not None else None for v in self.image_dim]

    def _set_checkpoint_dir(self):
        # reference models/base_model.py:197-232: <output_dir>/gans/<dataset_name>
        self.checkpoint_dir = os.path.join(str(self.output_dir), "gans", str(self.dataset_name))

    def _build(self):
        # reference models/gan.py:101-105
        assert (self.batch_size % self.rec_rr) == 0, 'Batch size should be divisable by random restart'
        self.test_batch_size = self.batch_size
        self.arch = _weights.canonical_arch(self.dataset_name)
        expect = list(_weights.IMAGE_DIMS[self.arch])
        if self.image_dim is None or any(v is None for v in self.image_dim):
            self.image_dim = expect
        if list(self.image_dim) != expect:
            raise ValueError("image_dim %s does not match the %s generator (%s)" % (self.image_dim, self.arch, expect))

    # -- weights -----------------------------------------------------------------------------
    def set_generator_weights(self, weights: Dict[str, np.ndarray]) -> None:
        """Install generator weights (names per the tflib.param registry, see weights.py)."""
        _weights.validate_weights(self.arch, weights, self.latent_dim, self.net_dim, bool(self.use_bn))
        self.weights = {k: np.asarray(v, dtype=np.float32) for k, v in weights.items()}
        self._drop_native()
        self.initialized = True

    def load_generator(self, ckpt_path=None):
        """Restore the generator (reference models/gan.py:80-87 -> base_model.py:294-335).
        `ckpt_path` is a directory holding `generator.npz` or the file itself; default is the
        model's checkpoint dir.  Returns False (and keeps the random-init weights, like the
        reference's failed restore, base_model.py:312-317) when nothing is found."""
        path = ckpt_path if ckpt_path is not None else self.checkpoint_dir
        if os.path.isdir(path):
            path = os.path.join(path, "generator.npz")
        if not os.path.isfile(path):
            if self.verbose:
                print("[-] No generator checkpoint found at {}; keeping random-init weights".format(path))
            return False
        self.set_generator_weights(_weights.load_npz(path))
        if self.verbose:
            print("[*] Generator restored from {}".format(path))
        return True

    def save_generator(self, ckpt_path=None):
        path = ckpt_path if ckpt_path is not None else self.checkpoint_dir
        if not path.endswith(".npz"):
            os.makedirs(path, exist_ok=True)
            path = os.path.join(path, "generator.npz")
        _weights.save_npz(path, self.weights)
        return path

    # -- native handle -------------------------------------------------------------------------
    def _drop_native(self):
        if getattr(self, "_native", None) is not None:
            self._native.close()
        self._native = None
        self._native_key = None

    def _get_native(self, device) -> "_native.NativeGenerator":
        key = (str(device), self.precision, int(self.latent_dim), int(self.net_dim), bool(self.use_bn))
        if self._native is None or self._native_key != key:
            self._drop_native()
            ordered = _weights.validate_weights(self.arch, self.weights, self.latent_dim, self.net_dim, bool(self.use_bn))
            tensors = [torch.as_tensor(np.ascontiguousarray(w), dtype=torch.float32).to(device) for w in ordered]
            self._native = _native.NativeGenerator(self.arch, tensors, latent_dim=self.latent_dim, net_dim=self.net_dim,
                                                   use_bn=bool(self.use_bn), precision=self.precision, device=device)
            self._native_key = key
        return self._native

    # -- the hot path ------------------------------------------------------------------------------
    def input_transform(self, X):
        raise NotImplementedError

    def generator_fn(self, z=None, is_training=False):
        """G(z) (reference models/gan.py:657-665,726-735): z [N, latent] -> images [N,H,W,C]."""
        if z is None:
            raise ValueError("z must be given (sampling inside generator_fn is a training-time feature)")
        z = self._as_cuda(z)
        return self._get_native(z.device).forward(z)

    @staticmethod
    def _as_cuda(t):
        if isinstance(t, np.ndarray):
            t = torch.from_numpy(np.ascontiguousarray(t))
        if not isinstance(t, torch.Tensor):
            raise TypeError("expected a torch.Tensor or numpy array")
        if not t.is_cuda:
            if not torch.cuda.is_available():
                raise RuntimeError("defensegan_b200 needs a CUDA (sm_100) device; there is no CPU fallback")
            t = t.cuda(non_blocking=True)
        return t.to(torch.float32)

    def reconstruct(self, images, batch_size=None, back_prop=True, reconstructor_id=0, z_init_val=None,
                    return_aux=False, out=None):
        """Defense-GAN projection of `images` onto the generator's range (reference
        models/gan.py:333-449): rec_rr restarts x rec_iters momentum-GD steps on
        ||G(z) - x||^2, returns G(z) of the min-loss restart.  Hyper-parameters are read from the
        object at call time.  Fresh z0 ~ N(0, 1/latent_dim) and zero momentum on every call
        (utils/gan_defense.py:119) unless `z_init_val` [B*rec_rr, latent_dim] is given
        (models/gan.py:395-397)."""
        x = self._as_cuda(images)
        if x.dim() != 4 or list(x.shape[1:]) != list(self.image_dim):
            raise ValueError("images must be [B,%d,%d,%d], got %s" % (tuple(self.image_dim) + (tuple(x.shape),)))
        if batch_size is not None and int(batch_size) != x.shape[0]:
            raise ValueError("batch_size (%d) does not match images.shape[0] (%d)" % (int(batch_size), x.shape[0]))
        z0 = self._as_cuda(z_init_val) if z_init_val is not None else None
        native = self._get_native(x.device)
        self._call_counter += 1
        seed = (int(self.seed) * 1000003 + int(reconstructor_id) * 7919 + self._call_counter) & (2 ** 63 - 1)
        res = native.reconstruct(x, int(self.rec_rr), int(self.rec_iters), float(self.rec_lr), z_init_val=z0, seed=seed,
                                 momentum=float(self.rec_momentum), decay_lr=bool(self.rec_decay_lr), out=out,
                                 return_aux=return_aux)
        return res

    def close(self):
        self._drop_native()


class MnistDefenseGAN(DefenseGANBase):
    """reference models/gan.py:649-685"""
    _dataset_default = "mnist"
    _image_dim_default = [28, 28, 1]

    def input_transform(self, X):
        return torch.as_tensor(X).to(torch.float32) / 255.0


class FmnistDefenseDefenseGAN(MnistDefenseGAN):
    """reference models/gan.py:688-698 (same generator as MNIST, different weights/data)"""
    _dataset_default = "f-mnist"


class CelebADefenseGAN(DefenseGANBase):
    """reference models/gan.py:718-765"""
    _dataset_default = "celeba"
    _image_dim_default = [64, 64, 3]

    def input_transform(self, images):
        return 2 * ((torch.as_tensor(images).to(torch.float32) / 255.0) - 0.5)

    def imsave_transform(self, imgs):
        imgs = (imgs + 1.0) / 2
        return imgs.clamp(0.0, 1.0) if isinstance(imgs, torch.Tensor) else np.clip(imgs, 0.0, 1.0)


# reference blackbox.py:56-61 / whitebox.py:49-53
dataset_gan_dict = {
    "mnist": MnistDefenseGAN,
    "f-mnist": FmnistDefenseDefenseGAN,
    "celeba": CelebADefenseGAN,
}
