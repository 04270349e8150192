"""Generator weight tensors: names, shapes, reference-style initialisation, npz I/O.

Names and creation order follow the reference's `tflib.param` registry (tflib/__init__.py:7-33)
as used by mnist_generator / celeba_generator (models/dataset_models.py:36-71,127-165):
`Generator.Input/Generator.Input.W` (in,out) + `.b`, `Generator.N/Generator.N.Filters`
(kh,kw,Cout,Cin) + `.Biases`.  The native library takes them as a list in this order.
"""
from __future__ import annotations

import collections
from typing import Dict, List

import numpy as np

ARCH_ALIASES = {"mnist": "mnist", "f-mnist": "mnist", "fmnist": "mnist", "celeba": "celeba"}
IMAGE_DIMS = {"mnist": (28, 28, 1), "celeba": (64, 64, 3)}


def canonical_arch(name: str) -> str:
    try:
        return ARCH_ALIASES[str(name).lower()]
    except KeyError:
        raise ValueError("unknown dataset / generator architecture %r" % (name,))


def weight_specs(arch: str, latent_dim: int = 128, net_dim: int = 64, use_bn: bool = False):
    """[(name, shape)] in creation order."""
    arch = canonical_arch(arch)
    n_feat = 4 * 4 * 4 * net_dim
    specs = [("Generator.Input/Generator.Input.W", (latent_dim, n_feat)),
             ("Generator.Input/Generator.Input.b", (n_feat,))]
    if use_bn:
        specs += [("Generator.BN1.offset", (1, n_feat)), ("Generator.BN1.scale", (1, n_feat))]
    if arch == "mnist":
        deconvs = [("Generator.2", 4 * net_dim, 2 * net_dim), ("Generator.3", 2 * net_dim, net_dim),
                   ("Generator.5", net_dim, 1)]
    else:
        deconvs = [("Generator.2", 4 * net_dim, 2 * net_dim), ("Generator.3", 2 * net_dim, net_dim),
                   ("Generator.5", net_dim, net_dim), ("Generator.6", net_dim, 3)]
    for i, (name, c_in, c_out) in enumerate(deconvs):
        specs += [("%s/%s.Filters" % (name, name), (5, 5, c_out, c_in)), ("%s/%s.Biases" % (name, name), (c_out,))]
        if use_bn and i < 2:
            specs += [("Generator.BN%d.offset" % (i + 2), (1, 1, 1, c_out)),
                      ("Generator.BN%d.scale" % (i + 2), (1, 1, 1, c_out))]
    return specs


def init_generator_weights(arch: str, seed: int = 11241990, latent_dim: int = 128, net_dim: int = 64,
                           use_bn: bool = False) -> "collections.OrderedDict[str, np.ndarray]":
    """Random-init weights the way the reference draws them at graph-construction time:
    Linear: uniform(+-sqrt(3)*sqrt(2/(in+out))) (tflib/ops/linear.py:41-60), bias 0 (:135-142);
    Deconv2D: uniform(+-sqrt(3)*sqrt(4/(fan_in+fan_out))), fan_in = Cin*25/4, fan_out = Cout*25
    (tflib/ops/deconv2d.py:49-74), bias 0 (:111-117); BN offset 0 / scale 1
    (tflib/ops/batchnorm.py:88-89).  Used when no trained checkpoint is available (synthetic
    benchmarks, tests)."""
    rs = np.random.RandomState(seed)
    out = collections.OrderedDict()
    for name, shape in weight_specs(arch, latent_dim, net_dim, use_bn):
        if name.endswith(".W"):
            stdev = np.sqrt(2.0 / (shape[0] + shape[1]))
            out[name] = rs.uniform(-stdev * np.sqrt(3), stdev * np.sqrt(3), size=shape).astype("float32")
        elif name.endswith(".Filters"):
            k, _, c_out, c_in = shape
            stdev = np.sqrt(4.0 / (c_in * k * k / 4.0 + c_out * k * k))
            out[name] = rs.uniform(-stdev * np.sqrt(3), stdev * np.sqrt(3), size=shape).astype("float32")
        elif name.endswith(".scale"):
            out[name] = np.ones(shape, dtype="float32")
        else:
            out[name] = np.zeros(shape, dtype="float32")
    return out


def validate_weights(arch: str, weights: Dict[str, np.ndarray], latent_dim: int, net_dim: int, use_bn: bool) -> List[np.ndarray]:
    """Check names/shapes and return the tensors as the ordered list the C-ABI expects."""
    ordered = []
    for name, shape in weight_specs(arch, latent_dim, net_dim, use_bn):
        if name not in weights:
            short = name.split("/")[-1]
            if short in weights:
                name_key = short
            else:
                raise KeyError("generator weight %r is missing" % name)
        else:
            name_key = name
        arr = weights[name_key]
        if tuple(arr.shape) != tuple(shape):
            raise ValueError("weight %r has shape %s, expected %s" % (name, tuple(arr.shape), tuple(shape)))
        ordered.append(arr)
    return ordered


def save_npz(path: str, weights: Dict[str, np.ndarray]) -> None:
    np.savez(path, **{k.replace("/", "__"): np.asarray(v) for k, v in weights.items()})


def load_npz(path: str) -> "collections.OrderedDict[str, np.ndarray]":
    out = collections.OrderedDict()
    with np.load(path) as f:
        for k in f.files:
            out[k.replace("__", "/")] = f[k]
    return out
