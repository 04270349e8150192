"""YAML config + command-line flags (mirror of reference utils/config.py:36-94).

`load_config(path)` reads `<path>` (a `.yml` file, or an experiment output dir holding
`cfg.yml`, reference :56-59) on top of the sibling `default.yml` (:61-67) and returns a dict
with the reference's upper-case keys.  Every key also becomes a lower-case `--flag`
(`add_flags`), e.g. `--rec_iters --rec_lr --rec_rr --batch_size --latent_dim --net_dim
--use_bn --image_dim`, with the reference's type dispatch (:26-33,73-88).
"""
from __future__ import annotations

import argparse
import os
from typing import Dict, Optional

import yaml

PACKAGED_CFG_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cfgs", "gans")
_DATASET_CFG = {"mnist": "mnist.yml", "f-mnist": "fmnist.yml", "fmnist": "fmnist.yml", "celeba": "celeba.yml"}


def packaged_cfg_path(dataset_name: str) -> str:
    return os.path.join(PACKAGED_CFG_DIR, _DATASET_CFG[str(dataset_name).lower()])


def _read_yaml(path: str) -> Dict:
    with open(path, "r") as f:
        return yaml.safe_load(f) or {}


def load_config(cfg_path: str, set_flag: bool = False, verbose: bool = False) -> Dict:
    if cfg_path is None or not os.path.exists(cfg_path):
        raise IOError("config path %r does not exist" % (cfg_path,))
    if os.path.isdir(cfg_path):                       # an experiment output dir
        cfg_file = os.path.join(cfg_path, "cfg.yml")
        cfg = _read_yaml(cfg_file)
    else:
        cfg_file = cfg_path
        cfg = {}
        default = os.path.join(os.path.dirname(cfg_path), "default.yml")
        if os.path.exists(default) and os.path.abspath(default) != os.path.abspath(cfg_path):
            cfg.update(_read_yaml(default))
        cfg.update(_read_yaml(cfg_path))
    cfg["cfg_path"] = cfg_file
    if verbose:
        for k in sorted(cfg):
            print("[cfg] {} = {}".format(k, cfg[k]))
    return cfg


def _str2bool(v):
    if isinstance(v, bool):
        return v
    if str(v).lower() in ("true", "t", "1", "yes", "y"):
        return True
    if str(v).lower() in ("false", "f", "0", "no", "n"):
        return False
    raise argparse.ArgumentTypeError("boolean value expected, got %r" % (v,))


def add_flags(parser: argparse.ArgumentParser, cfg: Dict) -> argparse.ArgumentParser:
    """One lower-case flag per config key, typed like the reference's flag dispatch."""
    for key, val in cfg.items():
        flag = "--" + key.lower()
        if any(flag in a.option_strings for a in parser._actions):
            continue
        if isinstance(val, bool):
            parser.add_argument(flag, type=_str2bool, default=val)
        elif isinstance(val, int):
            parser.add_argument(flag, type=int, default=val)
        elif isinstance(val, float):
            parser.add_argument(flag, type=float, default=val)
        elif isinstance(val, (list, tuple)):
            parser.add_argument(flag, type=type(val[0]) if len(val) else str, nargs="*", default=list(val))
        else:
            parser.add_argument(flag, type=str, default=val)
    return parser


def flags_to_cfg(ns: argparse.Namespace, cfg: Optional[Dict] = None) -> Dict:
    """Fold parsed flags back into an upper-case-key cfg dict."""
    out = dict(cfg or {})
    for k, v in vars(ns).items():
        out[k.upper()] = v
    return out
