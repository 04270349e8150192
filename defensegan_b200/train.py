"""Cache-producing entry point: the `--save_recs` / `--save_ds` half of the reference's `train.py` (:25-56, flags
:66-101).  It fills the two on-disk caches the attack scripts read (`utils/experiment.py`):

    python -m defensegan_b200.train --cfg <gan cfg> --dataset_npz data.npz --save_ds
    python -m defensegan_b200.train --cfg <gan cfg> --dataset_npz data.npz --save_recs [--max_num N] [--init_path CKPT]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 -m defensegan_b200.train ... --save_recs

(the last form shards every batch over the GPUs, `DefenseGANBase.reconstruct_dataset`).  GAN / encoder training
(`--is_train`, `--train_encoder`) and sample plotting (`--test_generator`, `--test_batch`) are outside this package and
are refused by name.  Dataset readers are outside it too: `--dataset_npz` names an `.npz` with RAW images and integer
labels per split - `train_x, train_y, dev_x, dev_y, test_x, test_y` (uint8 `[N,H,W,C]`, what the reference's readers
yield before `input_transform`); programmatic callers bind their own readers with `gan.set_dataset_generators`.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

from .models.gan import dataset_gan_dict
from .utils.config import add_flags, load_config
from .utils.experiment import _flag_bool

_OUT_OF_SCOPE = ("is_train", "train_encoder", "test_generator", "test_decoder", "test_batch", "init_with_enc")


def bind_npz_dataset(gan, path, batch_size=None):
    """`<split>_x`, `<split>_y` arrays of `path` as the model's `<split>_gen_test` batch generators."""
    arrays = np.load(path)
    bs = int(batch_size or gan.batch_size)

    def make(split):
        x, y = arrays[split + "_x"], arrays[split + "_y"]

        def batches():
            for i in range(0, len(x), bs):
                yield x[i:i + bs], y[i:i + bs]
        return batches

    gan.set_dataset_generators(train=make("train"), dev=make("dev"), test=make("test"))
    return gan


def main(cfg, flags):
    refused = [f for f in _OUT_OF_SCOPE if getattr(flags, f, False)]
    if refused:
        raise SystemExit("--%s: GAN / encoder training and sample plotting are not part of defensegan_b200 "
                         "(projection loop only); train the generator with the reference and load its checkpoint"
                         % ", --".join(refused))
    local_rank = os.environ.get("LOCAL_RANK")
    if local_rank is not None:                       # launched by torch.distributed.run: one rank per GPU
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(local_rank))
        if not dist.is_initialized():
            dist.init_process_group("nccl")
    gan = dataset_gan_dict[cfg["DATASET_NAME"]](cfg=cfg, test_mode=True)
    gan.test_again, gan.debug = bool(flags.test_again), bool(flags.debug)
    if flags.dataset_npz:
        bind_npz_dataset(gan, flags.dataset_npz)
    if flags.save_recs:
        rets = gan.reconstruct_dataset(ckpt_path=flags.init_path, max_num=flags.max_num)
        if local_rank in (None, "0"):
            gan.save_recs(rets, max_num=flags.max_num)
    if flags.save_ds:
        if local_rank in (None, "0"):
            gan.save_ds()
    return gan


def _parse(argv):
    first = argparse.ArgumentParser(add_help=False)
    first.add_argument("--cfg", required=True, help="Config file")
    known, _ = first.parse_known_args(argv)
    cfg = load_config(known.cfg)
    parser = argparse.ArgumentParser()
    parser.add_argument("--cfg", required=True, help="Config file")
    for name in ("save_recs", "save_ds", "debug", "test_again") + _OUT_OF_SCOPE:
        parser.add_argument("--" + name, type=_flag_bool, nargs="?", const=True, default=False)
    parser.add_argument("--max_num", type=int, default=-1)
    parser.add_argument("--init_path", type=str, default=None)
    parser.add_argument("--dataset_npz", type=str, default=None)
    add_flags(parser, cfg)
    ns = parser.parse_args(argv)
    for k, v in vars(ns).items():
        if k.upper() in cfg:
            cfg[k.upper()] = v
    return cfg, ns


if __name__ == "__main__":
    if len(sys.argv) == 1:
        print(__doc__)
        sys.exit(1)
    main(*_parse(sys.argv[1:]))
