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
import pickle
import time
from typing import Callable, Dict, Iterable, Optional

import numpy as np
import torch

from .. import _native
from .. import tf_bundle as _tf_bundle
from .. import weights as _weights
from ..utils.config import load_config, packaged_cfg_path

__all__ = ["RecCache", "DefenseGANBase", "MnistDefenseGAN", "FmnistDefenseDefenseGAN", "CelebADefenseGAN", "dataset_gan_dict"]


class RecCache(object):
    """On-disk cache of one split's reconstructions, in the reference's layout (models/gan.py:466-478,503-507) so that
    its consumers (blackbox.py:249-259,294-329) keep working:

        <dir>/feats.pkl                         the whole split, one pickled array (written by save_recs)
        <dir>/pickles/rec_{index:07d}_l{label}.pkl   one pickled [H,W,C] array per image

    `index` = position of the image in the split.  Unreadable or missing entries simply count as misses (the reference
    tolerates broken cache files the same way, gan.py:486-496,526-534)."""

    def __init__(self, directory, reuse=True):
        self.directory = directory
        self.reuse = reuse
        self.pickle_dir = os.path.join(directory, 'pickles')
        os.makedirs(self.pickle_dir, exist_ok=True)

    @property
    def split_path(self):
        return os.path.join(self.directory, 'feats.pkl')

    def image_path(self, index, label):
        return os.path.join(self.pickle_dir, 'rec_{:07d}_l{}.pkl'.format(int(index), label))

    @staticmethod
    def _read(path):
        try:
            with open(path, 'rb') as f:
                return pickle.load(f)
        except Exception:
            return None

    def load_split(self):
        return self._read(self.split_path) if self.reuse else None

    def load_batch(self, first_index, labels):
        """The batch's reconstructions if EVERY image of it is cached, else None."""
        if not self.reuse:
            return None
        out = []
        for i, label in enumerate(labels):
            r = self._read(self.image_path(first_index + i, label))
            if r is None:
                return None
            out.append(r)
        return np.stack(out) if out else None

    def store_batch(self, first_index, labels, recs):
        for i, label in enumerate(labels):
            with open(self.image_path(first_index + i, label), 'wb') as f:
                pickle.dump(recs[i], f, protocol=pickle.HIGHEST_PROTOCOL)

    def store_split(self, recs):
        with open(self.split_path, 'wb') as f:
            pickle.dump(recs, f, pickle.HIGHEST_PROTOCOL)


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
        """Where the model's snapshots live (reference models/base_model.py:197-232, test-mode branch): the directory
        of an experiment's own `cfg.yml`, else `<output_dir>/<cfg path relative to experiments/cfgs, without .yml>`
        (`experiments/cfgs/gans/mnist.yml` -> `output/gans/mnist`).  The packaged cfgs map the same way."""
        cfg_file = str((self.cfg or {}).get('cfg_path', '') or '')
        if os.path.basename(cfg_file) == 'cfg.yml':
            self.checkpoint_dir = os.path.dirname(cfg_file)
            return
        rel = None
        norm = cfg_file.replace(os.sep, '/')
        for marker in ('experiments/cfgs/', '/cfgs/'):
            if marker in norm:
                rel = norm.split(marker)[-1]
                break
        if rel is None:
            rel = os.path.join('gans', str(self.dataset_name))
        if rel.endswith('.yml'):
            rel = rel[:-4]
        self.checkpoint_dir = os.path.join(str(self.output_dir), rel)

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
        `ckpt_path` is the model's checkpoint dir (default), a `generator.npz`, or a TF checkpoint prefix
        (`.../GAN.model-20000`).  A directory is searched for `generator.npz` first, then for the latest
        TensorFlow checkpoint-V2 bundle (`checkpoint` state file / `*.index`), which is read without
        TensorFlow by `defensegan_b200.tf_bundle` - only the `Generator*` variables, as the reference does.
        Returns False (and keeps the random-init weights, like the reference's failed restore,
        base_model.py:312-317) when nothing is found."""
        path = ckpt_path if ckpt_path is not None else self.checkpoint_dir
        for suffix in ('.index', '.meta'):                       # a bundle's file name instead of its prefix
            if path.endswith(suffix):
                path = path[:-len(suffix)]
        if '.data-' in os.path.basename(path):
            path = path[:path.rindex('.data-')]
        npz, prefix = None, None
        if os.path.isdir(path):
            if os.path.isfile(os.path.join(path, "generator.npz")):
                npz = os.path.join(path, "generator.npz")
            else:
                prefix = _tf_bundle.latest_checkpoint(path)
        elif os.path.isfile(path):
            npz = path
        elif os.path.isfile(path + ".index"):
            prefix = path
        if npz is not None:
            self.set_generator_weights(_weights.load_npz(npz))
        elif prefix is not None:
            self.set_generator_weights(_tf_bundle.read_generator_variables(prefix))
        else:
            msg = "[-] No generator checkpoint found at {}; keeping random-init weights".format(path)
            if self.test_mode:
                # the reference's callers ignore the return value (blackbox.py:640-642): projecting onto an untrained
                # generator must not pass silently
                import warnings
                warnings.warn(msg, RuntimeWarning, stacklevel=2)
            elif self.verbose:
                print(msg)
            return False
        if self.verbose:
            print("[*] Generator restored from {}".format(npz or prefix))
        return True

    def save_generator(self, ckpt_path=None, fmt="npz", global_step=0):
        """`fmt="npz"`: `generator.npz`; `fmt="tf"`: a TensorFlow checkpoint-V2 bundle `GAN.model-<step>` plus the
        `checkpoint` state file, the layout of the reference's saver (base_model.py:383-395)."""
        path = ckpt_path if ckpt_path is not None else self.checkpoint_dir
        if fmt == "tf":
            os.makedirs(path, exist_ok=True)
            name = "GAN.model-%d" % int(global_step)
            _tf_bundle.write_bundle(os.path.join(path, name), dict(self.weights))
            with open(os.path.join(path, "checkpoint"), "w") as f:
                f.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (name, name))
            return os.path.join(path, name)
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

    def _next_seed(self, reconstructor_id=0):
        """Philox key of the next call's z0 stream (fresh per call, like re-running the initialiser, gan_defense.py:119)."""
        self._call_counter += 1
        return (int(self.seed) * 1000003 + int(reconstructor_id) * 7919 + self._call_counter) & (2 ** 63 - 1)

    def reconstruct(self, images, batch_size=None, back_prop=True, reconstructor_id=0, z_init_val=None,
                    return_aux=False, out=None, z_row_offset=0):
        """Defense-GAN projection of `images` onto the generator's range (reference
        models/gan.py:333-449): rec_rr restarts x rec_iters momentum-GD steps on
        ||G(z) - x||^2, returns G(z) of the min-loss restart.  Hyper-parameters are read from the
        object at call time.  Fresh z0 ~ N(0, 1/latent_dim) and zero momentum on every call
        (utils/gan_defense.py:119) unless `z_init_val` [B*rec_rr, latent_dim] is given
        (models/gan.py:395-397).  `z_row_offset` (sharded callers only): index of the first latent row of `images`
        in the call's z0 stream, so that a batch split over several GPUs draws what one GPU would."""
        x = self._as_cuda(images)
        if x.dim() != 4 or list(x.shape[1:]) != list(self.image_dim):
            raise ValueError("images must be [B,%d,%d,%d], got %s" % (tuple(self.image_dim) + (tuple(x.shape),)))
        if batch_size is not None and int(batch_size) != x.shape[0]:
            raise ValueError("batch_size (%d) does not match images.shape[0] (%d)" % (int(batch_size), x.shape[0]))
        z0 = self._as_cuda(z_init_val) if z_init_val is not None else None
        native = self._get_native(x.device)
        self.last_seed = seed = self._next_seed(reconstructor_id)
        res = native.reconstruct(x, int(self.rec_rr), int(self.rec_iters), float(self.rec_lr), z_init_val=z0, seed=seed,
                                 momentum=float(self.rec_momentum), decay_lr=bool(self.rec_decay_lr), out=out,
                                 return_aux=return_aux, z_row_offset=int(z_row_offset))
        return res

    # -- bulk offline reconstruction + its on-disk cache (reference models/gan.py:451-587, 604-646) --------------
    def set_dataset_generators(self, train: Optional[Callable] = None, dev: Optional[Callable] = None,
                               test: Optional[Callable] = None) -> None:
        """The reference binds `train_gen_test` / `dev_gen_test` / `test_gen_test` to its dataset readers
        (models/gan.py:666-673; out of scope here).  Each is a zero-argument callable returning an iterable of
        `(images, targets)` batches of RAW images (what `real_data_test_pl` is fed with); `input_transform` is applied."""
        if train is not None:
            self.train_gen_test = train
        if dev is not None:
            self.dev_gen_test = dev
        if test is not None:
            self.test_gen_test = test

    def rec_cache_dir(self, split: str, max_num: int = -1) -> str:
        """`<checkpoint_dir>/recs_rr{R}_lr{lr:.5f}_iters{L}[_num{n}]/<split>[_debug]` - the directory name the
        callers parse back with `recs_rr(.*)_lr(.*)_iters(.*)` (blackbox.py:646-651)."""
        if max_num > 0:
            name = 'recs_rr{:d}_lr{:.5f}_iters{:d}_num{:d}'.format(int(self.rec_rr), float(self.rec_lr),
                                                                   int(self.rec_iters), int(max_num))
        else:
            name = 'recs_rr{:d}_lr{:.5f}_iters{:d}'.format(int(self.rec_rr), float(self.rec_lr), int(self.rec_iters))
        out = os.path.join(self.checkpoint_dir, name, split)
        if self.debug:
            out += '_debug'
        return out

    def _transformed_batch(self, images) -> torch.Tensor:
        x = self.input_transform(np.asarray(images))
        x = x.reshape([-1] + list(self.image_dim))
        return x

    def _dataset_projector(self):
        """(project, is_writer, agree) for reconstruct_dataset.  One process: `self.reconstruct`.  Under
        torch.distributed with several ranks (the bulk offline job on the 8 GPUs of a box): every rank walks the same
        batches, each batch is sharded over the ranks (`parallel.reconstruct_sharded`, one all-gather), rank 0 alone
        writes the cache and its hit / miss decisions are broadcast so that no rank can skip a collective another
        rank enters."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return self.reconstruct, True, (lambda flag: flag)
        from ..parallel import reconstruct_sharded
        dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")

        def project(x):
            return reconstruct_sharded(self, x.to(dev))

        def agree(flag):
            t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=dev)
            dist.broadcast(t, src=0)
            return bool(t.item())

        return project, dist.get_rank() == 0, agree

    def reconstruct_dataset(self, ckpt_path=None, max_num=-1, max_num_load=-1):
        """Projects the train/dev/test splits batch by batch (fresh z0 and zero momentum per batch, reference
        models/gan.py:541) behind the reference's two-level result cache (see `RecCache`).  Returns
        `{split: [all_recs, all_targets, orig_imgs]}` (numpy, images `[-1] + image_dim`), the reference's return value
        (models/gan.py:451-587).  Called on every rank of an initialised torch.distributed group it splits each batch
        over the ranks (`_dataset_projector`); every rank returns the full result."""
        if not self.initialized:
            self.load_generator(ckpt_path=ckpt_path)
        limit = max(max_num, max_num_load)
        shape = [-1] + list(self.image_dim)
        results = {}
        project, is_writer, agree = self._dataset_projector()
        for split in ('train', 'dev', 'test'):
            batches = getattr(self, split + '_gen_test', None)
            if batches is None:
                raise RuntimeError("no '{}_gen_test' generator bound: call set_dataset_generators(...) first "
                                   "(dataset readers are outside this package)".format(split))
            cache = RecCache(self.rec_cache_dir(split, max_num), reuse=not self.test_again)
            whole_split = cache.load_split()
            if not agree(whole_split is not None):          # rank 0's view of the cache decides for everybody
                whole_split = None
            elif whole_split is None:
                raise RuntimeError("rank 0 read {} but this rank cannot (the cache must be on a file system all "
                                   "ranks see)".format(cache.split_path))
            recs, targets, originals = [], [], []
            t_start = time.time()
            first = 0                                   # position of the batch's first image in the split
            for b, (images, labels) in enumerate(batches()):
                n = len(images)
                if (limit > -1 and first > limit) or (self.debug and b > 2):
                    break
                x = self._transformed_batch(images)
                if whole_split is None:
                    r = cache.load_batch(first, labels)
                    if not agree(r is not None):
                        r = project(x).detach().cpu().numpy()
                        if is_writer:
                            cache.store_batch(first, labels, r)
                            if self.verbose:
                                print('[rec] {} batch {:d}: projected in {:.2f} s'.format(split, b, time.time() - t_start))
                    elif r is None:
                        raise RuntimeError("rank 0 found batch {:d} of '{}' in {} but this rank cannot read it".format(
                            b, split, cache.pickle_dir))
                    recs.append(r)
                targets.append(np.asarray(labels))
                originals.append(np.asarray(x.cpu() if isinstance(x, torch.Tensor) else x))
                first += n
            empty = np.zeros([0] + list(self.image_dim), dtype=np.float32)
            if whole_split is not None:
                all_recs = whole_split
            else:
                all_recs = np.concatenate(recs).reshape(shape) if recs else empty
            results[split] = [all_recs,
                              np.concatenate(targets) if targets else np.zeros([0], dtype=np.int64),
                              np.concatenate(originals).reshape(shape) if originals else empty]
        return results

    def save_recs(self, rets: Dict, max_num: int = -1) -> None:
        """Write each split's `feats.pkl` so that the next reconstruct_dataset (and the callers' cache loaders,
        blackbox.py:249-259,294-329) short-cut.  The reference leaves this to train.py --save_recs."""
        for split, (all_recs, _, _) in rets.items():
            RecCache(self.rec_cache_dir(split, max_num)).store_split(all_recs)

    def save_ds(self):
        """Dump the input-transformed dataset: `data/cache/<dataset>_pkl/<split>/feats.pkl` holding two
        consecutive pickles (images, targets) - the layout blackbox.get_cached_gan_data reads (:332-367)."""
        for split in ['train', 'dev', 'test']:
            output_dir = os.path.join('data', 'cache', '{}_pkl'.format(self.dataset_name), split)
            if self.debug:
                output_dir += '_debug'
            os.makedirs(output_dir, exist_ok=True)
            path = os.path.join(output_dir, 'feats.pkl')
            if os.path.exists(path) and not self.test_again:
                print('[#] Dataset is already saved.')
                return
            gen_func = getattr(self, '{}_gen_test'.format(split))
            imgs, tgts = [], []
            for images, targets in gen_func():
                x = self._transformed_batch(images)
                imgs.append(np.asarray(x.cpu() if isinstance(x, torch.Tensor) else x))
                tgts.append(np.asarray(targets))
            with open(path, 'wb') as f:
                pickle.dump(np.concatenate(imgs).reshape([-1] + list(self.image_dim)), f, pickle.HIGHEST_PROTOCOL)
                pickle.dump(np.concatenate(tgts), f, pickle.HIGHEST_PROTOCOL)

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
