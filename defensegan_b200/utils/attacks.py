"""The attack / train / eval helpers the reference's callers take from cleverhans, in PyTorch:

* `fgm` - Fast Gradient Method (cleverhans/attacks_tf.py:23-99): L-inf / L1 / L2, the model's own predictions as labels
  when `y` is None (no label leaking), optional clipping, targeted variant;
* `model_loss` (utils_tf.py:20-41: softmax cross-entropy on the logits), `model_train` (utils_tf.py:67-169: Adam,
  shuffled mini-batches, optional adversarial training), `model_eval` and `batch_eval` (utils_tf.py:171-243, 267-323:
  ceil(n / batch) batches, ragged last batch).

Models are callables mapping an NHWC batch to LOGITS (e.g. `utils.network_builder.MLP`).  These are the consumers on
either side of the projection loop: FGSM crafts the input of `gan.reconstruct`, the classifier
reads its output.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def model_loss(y, logits, mean=True):
    """softmax_cross_entropy_with_logits(labels=y, logits) with (possibly soft) one-hot labels."""
    out = -(y * F.log_softmax(logits, dim=-1)).sum(dim=-1)
    return out.mean() if mean else out


def fgm(model, x, y=None, eps=0.3, ord=np.inf, clip_min=None, clip_max=None, targeted=False):
    """x_adv = x + eps * normalised(grad_x loss(model(x), y))   (cleverhans/attacks_tf.py:23-99)."""
    x = x.detach().clone().requires_grad_(True)
    with torch.enable_grad():             # callers evaluate under no_grad (model_eval); the attack needs the graph
        logits = model(x)
        if y is None:
            # model predictions as ground truth to avoid label leaking (attacks_tf.py:52-56)
            y = (logits == logits.max(dim=1, keepdim=True).values).to(logits.dtype).detach()
        y = y / y.sum(dim=1, keepdim=True)
        loss = model_loss(y, logits, mean=False)
        if targeted:
            loss = -loss
        grad, = torch.autograd.grad(loss.sum(), x)
    red = tuple(range(1, x.dim()))
    if ord == np.inf:
        normalized = torch.sign(grad)
    elif ord == 1:
        normalized = grad / grad.abs().sum(dim=red, keepdim=True)
    elif ord == 2:
        normalized = grad / torch.sqrt((grad ** 2).sum(dim=red, keepdim=True))
    else:
        raise NotImplementedError("Only L-inf, L1 and L2 norms are currently implemented.")
    adv = x.detach() + eps * normalized
    if clip_min is not None and clip_max is not None:
        adv = adv.clamp(clip_min, clip_max)
    return adv.detach()


class FastGradientMethod(object):
    """cleverhans.attacks.FastGradientMethod(model).generate(x, **fgsm_params) (blackbox.py:530-534)."""

    def __init__(self, model, back="torch", sess=None):
        self.model = model

    def generate(self, x, **kwargs):
        return fgm(self.model, x, **kwargs)


def model_train(model, X_train, Y_train, args, predictions_adv=None, rng=None, device=None, evaluate=None):
    """Adam on softmax cross-entropy, `nb_epochs` passes over shuffled mini-batches (utils_tf.py:67-169).
    `predictions_adv`: callable x -> adversarial x for adversarial training (loss averaged over clean and adversarial)."""
    nb_epochs, lr, bs = int(args["nb_epochs"]), float(args["learning_rate"]), int(args["batch_size"])
    rng = rng if rng is not None else np.random.RandomState()
    device = device if device is not None else next(model.parameters()).device
    X = torch.as_tensor(np.asarray(X_train), dtype=torch.float32)
    Y = torch.as_tensor(np.asarray(Y_train), dtype=torch.float32)
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    n = len(X)
    nb_batches = int(math.ceil(float(n) / bs))
    for _ in range(nb_epochs):
        model.train()
        order = rng.permutation(n)
        for b in range(nb_batches):
            idx = order[b * bs:(b + 1) * bs]
            xb, yb = X[idx].to(device), Y[idx].to(device)
            loss = model_loss(yb, model(xb))
            if predictions_adv is not None:
                loss = (loss + model_loss(yb, model(predictions_adv(xb)))) / 2
            opt.zero_grad()
            loss.backward()
            opt.step()
        if evaluate is not None:
            evaluate()
    model.eval()
    return True


def batch_eval(fn, X, batch_size, device):
    """fn over X in order, ragged last batch kept (utils_tf.py:267-323); returns the concatenated outputs on the CPU."""
    outs = []
    X = torch.as_tensor(np.asarray(X), dtype=torch.float32)
    for s in range(0, len(X), batch_size):
        outs.append(fn(X[s:s + batch_size].to(device)).detach().cpu())
    return torch.cat(outs) if outs else torch.zeros(0)


def model_eval(model, X_test, Y_test, args, device=None):
    """Accuracy of arg-max(model(x)) against one-hot Y (utils_tf.py:171-243)."""
    assert args.get("batch_size"), "Batch size was not given in args dict"
    if X_test is None or Y_test is None:
        raise ValueError("X_test argument and Y_test argument must be supplied.")
    device = device if device is not None else next(model.parameters()).device
    was_training = getattr(model, "training", False)
    if hasattr(model, "eval"):
        model.eval()
    with torch.no_grad():
        logits = batch_eval(model, X_test, int(args["batch_size"]), device)
    if was_training:
        model.train()
    y = np.asarray(Y_test)
    return float((logits.argmax(dim=-1).numpy() == y.argmax(axis=-1)).mean())
