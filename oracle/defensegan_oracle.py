"""CPU ORACLE (test infrastructure, NOT product code) for Defense-GAN's projection loop.

PARITY UNPINNED: the reference (kabkabm/defensegan, Python-2 / TensorFlow-1.7 graph
mode) cannot be imported or executed in this environment (no tensorflow, no keras,
no Python 2, no network) and it ships no tests, golden vectors or fixtures for this
path.  All arithmetic of the path lives in the
un-vendored third-party dependency TensorFlow 1.7 (README.md:46).  This file is
therefore a restatement *from reading the source*; it is pinned only against

  * a second, loop-level numpy statement of TF's `conv2d_transpose(k=5, stride=2,
    padding='SAME')` derived from TF's documented SAME-padding rule
    (pad_total = max((out-1)*stride + k - in, 0), pad_before = pad_total // 2), and
  * its own fp64 evaluation (used as "truth" to bound fp32 drift).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module, and only as the checker.

Reference lines each function follows (paths relative to /root/reference):

  init_generator_weights  tflib/ops/linear.py:41-60,135-142 (glorot-uniform, b=0);
                          tflib/ops/deconv2d.py:49-74,111-117 (he-uniform, b=0);
                          tflib/ops/batchnorm.py:85-89 (offset 0, scale 1);
                          creation order models/dataset_models.py:36-71,127-165;
                          names tflib/__init__.py:7-33
  tf_deconv_same          tflib/ops/deconv2d.py:100-110 (tf.nn.conv2d_transpose SAME, stride 2)
  batchnorm_batchstat     tflib/ops/batchnorm.py:9 (condition), :80-93 (else branch only)
  mnist_generator         models/dataset_models.py:36-71
  celeba_generator        models/dataset_models.py:127-165
  reconstruct             models/gan.py:333-449 (tile :355-359; z_hat init :370-377,395-397;
                          lr :380-386 + models/base_model.py:153-194 (effectively constant,
                          SURVEY F3); Momentum(0.7) :389-391; body :409-421; loop :430-437;
                          select :438-449)
  model_eval_gan          utils/gan_defense.py:32-179
"""
from __future__ import annotations

import collections
import math
from typing import Callable, Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

WEIGHT_SEED = 11241990          # blackbox.py:464 / whitebox.py:143 use this constant as the TF seed
IMAGE_SEED = 1990
Z0_SEED = 24

ARCH_SPECS = {
    # name: (H, W, C, final activation, crop-after-first-deconv)
    "mnist": dict(image_dim=(28, 28, 1), act="sigmoid", crop7=True,
                  deconvs=[("Generator.2", 4, 2), ("Generator.3", 2, 1), ("Generator.5", 1, 0)]),
    "celeba": dict(image_dim=(64, 64, 3), act="tanh", crop7=False,
                   deconvs=[("Generator.2", 4, 2), ("Generator.3", 2, 1), ("Generator.5", 1, 1),
                            ("Generator.6", 1, -3)]),
}
ARCH_ALIASES = {"mnist": "mnist", "f-mnist": "mnist", "fmnist": "mnist", "celeba": "celeba"}


def canonical_arch(name: str) -> str:
    try:
        return ARCH_ALIASES[name.lower()]
    except KeyError:
        raise ValueError("unknown generator architecture %r" % (name,))


def _deconv_channels(arch: str, net_dim: int):
    """[(name, c_in, c_out)] for the deconv stack; multipliers are of net_dim, negative = absolute."""
    out = []
    for name, m_in, m_out in ARCH_SPECS[arch]["deconvs"]:
        c_in = m_in * net_dim
        c_out = -m_out if m_out < 0 else (m_out * net_dim if m_out > 0 else 1)
        out.append((name, c_in, c_out))
    return out


# --------------------------------------------------------------------------------------
# weights
# --------------------------------------------------------------------------------------
def init_generator_weights(arch: str, seed: int = WEIGHT_SEED, latent_dim: int = 128,
                           net_dim: int = 64, use_bn: bool = False,
                           random_bias: bool = False) -> "collections.OrderedDict[str, np.ndarray]":
    """Synthetic generator weights drawn the way the reference initialises them.

    Draw order = variable creation order of the generator function.  `random_bias`
    (off = reference behaviour, biases 0) draws small non-zero biases / BN params *after*
    all reference-order draws so that tests also exercise the bias and affine paths.
    """
    arch = canonical_arch(arch)
    rs = np.random.RandomState(seed)
    w = collections.OrderedDict()

    def uniform(stdev, size):
        return rs.uniform(low=-stdev * np.sqrt(3), high=stdev * np.sqrt(3), size=size).astype("float32")

    n_feat = 4 * 4 * 4 * net_dim
    # tflib/ops/linear.py:55-60  initialization=None -> glorot
    w["Generator.Input/Generator.Input.W"] = uniform(np.sqrt(2.0 / (latent_dim + n_feat)), (latent_dim, n_feat))
    w["Generator.Input/Generator.Input.b"] = np.zeros((n_feat,), dtype="float32")
    if use_bn:
        w["Generator.BN1.offset"] = np.zeros((1, n_feat), dtype="float32")
        w["Generator.BN1.scale"] = np.ones((1, n_feat), dtype="float32")
    bn_idx = 2
    for li, (name, c_in, c_out) in enumerate(_deconv_channels(arch, net_dim)):
        k, stride = 5, 2
        fan_in = c_in * k ** 2 / (stride ** 2)       # deconv2d.py:49-50 (py2 int division is exact here)
        fan_out = c_out * k ** 2
        stdev = np.sqrt(4.0 / (fan_in + fan_out))    # he_init=True default
        w["%s/%s.Filters" % (name, name)] = uniform(stdev, (k, k, c_out, c_in))
        w["%s/%s.Biases" % (name, name)] = np.zeros((c_out,), dtype="float32")
        if use_bn and li < 2:
            w["Generator.BN%d.offset" % bn_idx] = np.zeros((1, 1, 1, c_out), dtype="float32")
            w["Generator.BN%d.scale" % bn_idx] = np.ones((1, 1, 1, c_out), dtype="float32")
            bn_idx += 1
    if random_bias:
        for key in list(w.keys()):
            if key.endswith(".b") or key.endswith(".Biases") or key.endswith(".offset"):
                w[key] = (0.1 * rs.standard_normal(w[key].shape)).astype("float32")
            elif key.endswith(".scale"):
                w[key] = (1.0 + 0.2 * rs.standard_normal(w[key].shape)).astype("float32")
    return w


def weights_to_torch(weights: Dict[str, np.ndarray], dtype=torch.float32) -> Dict[str, torch.Tensor]:
    return {k: torch.as_tensor(np.asarray(v)).to(dtype) for k, v in weights.items()}


# --------------------------------------------------------------------------------------
# ops
# --------------------------------------------------------------------------------------
def tf_deconv_same(x_nhwc: torch.Tensor, filt: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """tf.nn.conv2d_transpose(x, filt, [N,2H,2W,Cout], strides=[1,2,2,1], 'SAME') + bias.

    filt is (kh, kw, C_out, C_in) (deconv2d.py:67).  TF defines the op as the input-gradient
    of conv2d(SAME, stride 2) whose padding is pad_total = max((H-1)*2 + 5 - 2H, 0) = 3,
    pad_before = 1: forward conv reads in[2*o + k - 1], so the transpose scatters
    out[2*o + k - 1] += in[o] * w[k]  ==  full (VALID) transposed conv cropped [1 : 1+2H].
    """
    n, h, w_, c_in = x_nhwc.shape
    kh, kw, c_out, c_in2 = filt.shape
    assert c_in == c_in2 and kh == 5 and kw == 5
    x = x_nhwc.permute(0, 3, 1, 2)
    wt = filt.permute(3, 2, 0, 1)                      # torch conv_transpose2d weight: (C_in, C_out, kh, kw)
    full = F.conv_transpose2d(x, wt, stride=2)         # [N, C_out, 2H+3, 2W+3]
    out = full[:, :, 1:1 + 2 * h, 1:1 + 2 * w_]
    if bias is not None:
        out = out + bias.view(1, -1, 1, 1)
    return out.permute(0, 2, 3, 1).contiguous()


def tf_deconv_same_definition(x_nhwc: np.ndarray, filt: np.ndarray, bias: Optional[np.ndarray]) -> np.ndarray:
    """Loop-level numpy statement of the same op, for known-answer tests.

    out[n,i,j,co] = b[co] + sum_{o,p,ka,kb,ci : i=2o+ka-1, j=2p+kb-1} in[n,o,p,ci]*F[ka,kb,co,ci]
    """
    n, h, w_, c_in = x_nhwc.shape
    kh, kw, c_out, _ = filt.shape
    out = np.zeros((n, 2 * h, 2 * w_, c_out), dtype=np.float64)
    for o in range(h):
        for p in range(w_):
            for ka in range(kh):
                i = 2 * o + ka - 1
                if i < 0 or i >= 2 * h:
                    continue
                for kb in range(kw):
                    j = 2 * p + kb - 1
                    if j < 0 or j >= 2 * w_:
                        continue
                    # [n, ci] x [co, ci]^T
                    out[:, i, j, :] += x_nhwc[:, o, p, :].astype(np.float64) @ filt[ka, kb].astype(np.float64).T
    if bias is not None:
        out += bias.astype(np.float64).reshape(1, 1, 1, -1)
    return out


def tf_deconv_same_dinput_definition(dout: np.ndarray, filt: np.ndarray) -> np.ndarray:
    """Backward-to-input of the op above, loop-level:
    din[n,o,p,ci] = sum_{ka,kb,co} dout[n,2o+ka-1,2p+kb-1,co] * F[ka,kb,co,ci]."""
    n, h2, w2, c_out = dout.shape
    kh, kw, _, c_in = filt.shape
    h, w_ = h2 // 2, w2 // 2
    din = np.zeros((n, h, w_, c_in), dtype=np.float64)
    for o in range(h):
        for p in range(w_):
            for ka in range(kh):
                i = 2 * o + ka - 1
                if i < 0 or i >= h2:
                    continue
                for kb in range(kw):
                    j = 2 * p + kb - 1
                    if j < 0 or j >= w2:
                        continue
                    din[:, o, p, :] += dout[:, i, j, :].astype(np.float64) @ filt[ka, kb].astype(np.float64)
    return din


def batchnorm_batchstat(x: torch.Tensor, axes, offset: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    """tflib/ops/batchnorm.py:80-93: tf.nn.moments (biased variance) over `axes`, then
    tf.nn.batch_normalization(x, mean, var, offset, scale, 1e-5) = (x-mean)*rsqrt(var+eps)*scale+offset.
    The generator always lands here (axes [0] / [0,1,2] fail the fused-path test, :9), so batch
    statistics are used even at test time."""
    mean = x.mean(dim=axes, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=axes, keepdim=True)
    inv = torch.rsqrt(var + 1e-5) * scale
    return x * inv + (offset - mean * inv)             # TF's batch_normalization evaluates it in this form


def _wget(w, name):
    return w[name]


def generator_forward(arch: str, w: Dict[str, torch.Tensor], z: torch.Tensor, use_bn: bool = False,
                      return_hidden: bool = False):
    """mnist_generator / celeba_generator (models/dataset_models.py:36-71,127-165), NHWC."""
    arch = canonical_arch(arch)
    spec = ARCH_SPECS[arch]
    hidden = []
    out = z @ w["Generator.Input/Generator.Input.W"] + w["Generator.Input/Generator.Input.b"]
    if use_bn:
        out = batchnorm_batchstat(out, [0], w["Generator.BN1.offset"], w["Generator.BN1.scale"])
    out = torch.relu(out)
    n_feat = out.shape[1]
    out = out.reshape(-1, 4, 4, n_feat // 16)          # row-major NHWC reshape
    hidden.append(out)
    names = [d[0] for d in spec["deconvs"]]
    # Generator.2
    out = tf_deconv_same(out, w["%s/%s.Filters" % (names[0], names[0])], w["%s/%s.Biases" % (names[0], names[0])])
    if use_bn:
        out = batchnorm_batchstat(out, [0, 1, 2], w["Generator.BN2.offset"], w["Generator.BN2.scale"])
    out = torch.relu(out)
    if spec["crop7"]:
        out = out[:, :7, :7, :]                        # dataset_models.py:59 (crop AFTER relu)
    hidden.append(out)
    # Generator.3
    out = tf_deconv_same(out, w["%s/%s.Filters" % (names[1], names[1])], w["%s/%s.Biases" % (names[1], names[1])])
    if use_bn:
        out = batchnorm_batchstat(out, [0, 1, 2], w["Generator.BN3.offset"], w["Generator.BN3.scale"])
    out = torch.relu(out)
    hidden.append(out)
    if arch == "mnist":
        out = tf_deconv_same(out, w["Generator.5/Generator.5.Filters"], w["Generator.5/Generator.5.Biases"])
        out = torch.sigmoid(out)
    else:
        # celeba: Generator.5 has NO non-linearity before Generator.6 (dataset_models.py:158-163)
        out = tf_deconv_same(out, w["Generator.5/Generator.5.Filters"], w["Generator.5/Generator.5.Biases"])
        hidden.append(out)
        out = tf_deconv_same(out, w["Generator.6/Generator.6.Filters"], w["Generator.6/Generator.6.Biases"])
        out = torch.tanh(out)
    if return_hidden:
        return out, hidden
    return out


# --------------------------------------------------------------------------------------
# the projection loop
# --------------------------------------------------------------------------------------
def tile_images(images: torch.Tensor, rec_rr: int) -> torch.Tensor:
    """models/gan.py:355-359: reshape [B, HWC] -> tile [1, R] -> reshape [B*R, H, W, C]
    (image-major, restart-minor: rows iR..iR+R-1 are copies of image i)."""
    b = images.shape[0]
    flat = images.reshape(b, -1)
    tiled = flat.repeat(1, rec_rr)
    return tiled.reshape((b * rec_rr,) + tuple(images.shape[1:]))


def effective_learning_rate(rec_lr: float, rec_iters: int, step: int, emulate_dead_decay: bool = True) -> float:
    """models/gan.py:380-386 + base_model.py:185-192.  exponential_decay(rec_lr, global_step=
    rec_iter_const, decay_steps=ceil(0.8 L), 0.1, staircase=True) where rec_iter_const is a
    variable that is initialised to 0 and never assigned => lr == rec_lr at every step.
    emulate_dead_decay=False gives the evidently *intended* schedule (x0.1 from step ceil(0.8 L))."""
    if emulate_dead_decay:
        return float(rec_lr)
    decay_iter = int(np.ceil(rec_iters * 0.8))
    return float(rec_lr) * (0.1 ** (step // decay_iter))


def sample_z0(n_rows: int, latent_dim: int, seed: int = Z0_SEED) -> np.ndarray:
    """z_hat initialiser, models/gan.py:370-377: N(0, 1/latent_dim) i.i.d.  (TF's RNG stream cannot
    be reproduced; parity runs inject this array through `z_init_val`)"""
    rs = np.random.RandomState(seed)
    return (rs.standard_normal((n_rows, latent_dim)) * np.sqrt(1.0 / latent_dim)).astype("float32")


def reconstruct(arch: str, weights: Dict[str, np.ndarray], images: np.ndarray, rec_rr: int, rec_iters: int,
                rec_lr: float = 10.0, z_init_val: Optional[np.ndarray] = None, momentum: float = 0.7,
                use_bn: bool = False, dtype=torch.float32, emulate_dead_decay: bool = True,
                seed: int = Z0_SEED, quantize: Optional[Callable] = None, return_trace: bool = False):
    """DefenseGANBase.reconstruct (models/gan.py:333-449) evaluated eagerly on CPU.

    images: [B,H,W,C] already input-transformed.  Returns dict with
      rec [B,H,W,C], loss_min [B], idx [B] (restart chosen, 0..R-1), loss_all [B*R],
      rec_all [B*R,H,W,C] (pre-update forward of iteration L-1), z_final.
    """
    arch = canonical_arch(arch)
    w = weights_to_torch(weights, dtype)
    if quantize is not None:
        w = {k: (quantize(v) if (k.endswith(".W") or k.endswith(".Filters")) else v) for k, v in w.items()}
    x = torch.as_tensor(np.asarray(images)).to(dtype)
    b = x.shape[0]
    n_rows = b * rec_rr
    latent_dim = w["Generator.Input/Generator.Input.W"].shape[0]
    x_tiled = tile_images(x, rec_rr)
    if z_init_val is None:
        z_init_val = sample_z0(n_rows, latent_dim, seed)
    z = torch.as_tensor(np.asarray(z_init_val)).to(dtype).clone().reshape(n_rows, latent_dim)
    v = torch.zeros_like(z)                                   # Momentum slot, zero per batch (gan_defense.py:119)
    axes = tuple(range(1, x_tiled.dim()))
    trace = []
    y = None
    image_rec_loss = None
    for t in range(rec_iters):
        zt = z.detach().clone().requires_grad_(True)
        if quantize is not None:
            y = _quantized_forward(arch, w, zt, use_bn, quantize)
        else:
            y = generator_forward(arch, w, zt, use_bn=use_bn)
        image_rec_loss = ((y - x_tiled) ** 2).mean(dim=axes)  # gan.py:411-413
        rec_loss = image_rec_loss.sum()                       # gan.py:414
        (g,) = torch.autograd.grad(rec_loss, zt)              # minimize(..., var_list=[z_hat]) gan.py:416-417
        lr = effective_learning_rate(rec_lr, rec_iters, t, emulate_dead_decay)
        v = momentum * v + g                                  # tf.train.MomentumOptimizer (non-Nesterov)
        z = z - lr * v
        if return_trace:
            trace.append(image_rec_loss.detach().clone())
    # loop returns the pre-update forward/loss of the last iteration (gan.py:419-421)
    y = y.detach()
    loss = image_rec_loss.detach()
    loss_r = loss.reshape(b, rec_rr)
    idx = torch.argmin(loss_r, dim=1)                         # lowest index on ties, gan.py:439-444
    rows = torch.arange(b) * rec_rr + idx
    rec = y[rows].reshape(x.shape)
    out = dict(rec=rec.numpy(), loss_min=loss[rows].numpy(), idx=idx.numpy().astype(np.int32),
               loss_all=loss.numpy(), rec_all=y.numpy(), z_final=z.detach().numpy())
    if return_trace:
        out["trace"] = torch.stack(trace).numpy()
    return out


def loss_and_grad(arch: str, weights: Dict[str, np.ndarray], images: np.ndarray, z: np.ndarray, rec_rr: int,
                  use_bn: bool = False, dtype=torch.float32):
    """One evaluation of (G(z), per-row loss, d(sum loss)/dz) - the body of the loop, for layer-level KATs."""
    arch = canonical_arch(arch)
    w = weights_to_torch(weights, dtype)
    x_tiled = tile_images(torch.as_tensor(np.asarray(images)).to(dtype), rec_rr)
    zt = torch.as_tensor(np.asarray(z)).to(dtype).clone().requires_grad_(True)
    y = generator_forward(arch, w, zt, use_bn=use_bn)
    loss = ((y - x_tiled) ** 2).mean(dim=tuple(range(1, y.dim())))
    (g,) = torch.autograd.grad(loss.sum(), zt)
    return y.detach().numpy(), loss.detach().numpy(), g.numpy()


# --------------------------------------------------------------------------------------
# operand-quantised variant (precision study only; not part of the reference semantics)
# --------------------------------------------------------------------------------------
class _QuantSTE(torch.autograd.Function):
    """Round the tensor in forward AND round the incoming gradient in backward: models a
    tensor-core pipeline whose forward activations and backward gradients are both stored in a
    narrow operand format while accumulation stays fp32."""

    @staticmethod
    def forward(ctx, x, fn):
        ctx.fn = fn
        return fn(x)

    @staticmethod
    def backward(ctx, g):
        return ctx.fn(g), None


def _quantized_forward(arch, w, z, use_bn, q):
    spec = ARCH_SPECS[arch]
    qs = lambda t: _QuantSTE.apply(t, q)
    out = qs(z) @ w["Generator.Input/Generator.Input.W"] + w["Generator.Input/Generator.Input.b"]
    out = qs(torch.relu(out)).reshape(-1, 4, 4, out.shape[1] // 16)
    names = [d[0] for d in spec["deconvs"]]
    out = tf_deconv_same(out, w["%s/%s.Filters" % (names[0], names[0])], w["%s/%s.Biases" % (names[0], names[0])])
    out = torch.relu(out)
    if spec["crop7"]:
        out = out[:, :7, :7, :]
    out = qs(out)
    out = tf_deconv_same(out, w["%s/%s.Filters" % (names[1], names[1])], w["%s/%s.Biases" % (names[1], names[1])])
    out = qs(torch.relu(out))
    if arch == "mnist":
        out = tf_deconv_same(out, w["Generator.5/Generator.5.Filters"], w["Generator.5/Generator.5.Biases"])
        return torch.sigmoid(out)
    out = qs(tf_deconv_same(out, w["Generator.5/Generator.5.Filters"], w["Generator.5/Generator.5.Biases"]))
    out = tf_deconv_same(out, w["Generator.6/Generator.6.Filters"], w["Generator.6/Generator.6.Biases"])
    return torch.tanh(out)


def make_quantizer(kind: str) -> Callable:
    if kind == "fp16":
        return lambda t: t.to(torch.float16).to(t.dtype)
    if kind == "bf16":
        return lambda t: t.to(torch.bfloat16).to(t.dtype)
    if kind == "tf32":
        def q(t):
            i = t.detach().to(torch.float32).contiguous().view(torch.int32)
            i = (i + 0x1000) & ~0x1FFF                      # round-half-up to 10 mantissa bits
            return t + (i.view(torch.float32).to(t.dtype) - t).detach()
        return q
    raise ValueError(kind)


# --------------------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------------------
def synthetic_images(arch: str, weights: Dict[str, np.ndarray], batch: int, kind: str = "S1",
                     seed: int = IMAGE_SEED, latent_dim: int = 128) -> np.ndarray:
    """S1: on-manifold + noise, x = clip(G(z*) + 0.1*eps, lo, 1); S2: i.i.d. U[lo, 1]."""
    arch = canonical_arch(arch)
    h, w_, c = ARCH_SPECS[arch]["image_dim"]
    lo = 0.0 if ARCH_SPECS[arch]["act"] == "sigmoid" else -1.0
    rs = np.random.RandomState(seed)
    if kind == "S2":
        return rs.uniform(lo, 1.0, size=(batch, h, w_, c)).astype("float32")
    zstar = (rs.standard_normal((batch, latent_dim)) * np.sqrt(1.0 / latent_dim)).astype("float32")
    eps = rs.standard_normal((batch, h, w_, c)).astype("float32")
    with torch.no_grad():
        g = generator_forward(arch, weights_to_torch(weights), torch.as_tensor(zstar)).numpy()
    return np.clip(g + 0.1 * eps, lo, 1.0).astype("float32")


# --------------------------------------------------------------------------------------
# eval driver
# --------------------------------------------------------------------------------------
def model_eval_gan(reconstruct_fn: Callable, classify_fn: Callable, test_images: np.ndarray,
                   test_labels: np.ndarray, batch_size: int, diff_fn: Optional[Callable] = None):
    """utils/gan_defense.py:113-179 restated with callables in place of graph tensors:
    ceil(n/bs) batches, fresh state per batch (:119), ragged last batch (:126-128), accuracy =
    #correct / n (:165), roc_info = [labels, preds, diffs] (:175)."""
    n = len(test_images)
    nb_batches = int(math.ceil(float(n) / batch_size))
    acc, labels, preds, diffs = 0.0, [], [], []
    for bi in range(nb_batches):
        start, end = bi * batch_size, min(n, (bi + 1) * batch_size)
        xb, yb = test_images[start:end], test_labels[start:end]
        rec = reconstruct_fn(xb, bi)
        logits = classify_fn(rec)
        cur_labels = np.argmax(yb, axis=-1)
        cur_preds = np.argmax(logits, axis=-1)
        acc += float(np.sum(cur_labels == cur_preds))
        labels.append(cur_labels)
        preds.append(cur_preds)
        if diff_fn is not None:
            diffs.append(diff_fn(xb, rec))
    acc /= n
    roc = [np.concatenate(labels), np.concatenate(preds), np.concatenate(diffs) if diff_fn is not None else []]
    return acc, roc
