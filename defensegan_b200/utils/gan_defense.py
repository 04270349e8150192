"""Evaluation driver: mirror of reference utils/gan_defense.py:32-179 (`model_eval_gan`).

Graph-mode -> eager mapping.  The reference takes symbolic tensors that it
evaluates with `sess.run` once per batch; here the same arguments are *callables* evaluated
once per batch on a CUDA tensor holding that batch of `test_images`:

    predictions(x_batch)      -> logits [b, n_classes]   (usually classifier(gan.reconstruct(x)))
    predictions_rec(x_batch)  -> logits [b, n_classes]   (optional)
    diff_op(x_batch)          -> per-example differences [b]
                                 (reference: mean((x - reconstruct(x))**2), blackbox.py:569-572)

`sess`, `images`, `labels` (placeholders) and `feed` are accepted for signature compatibility
and ignored.  As in the reference, every batch starts from fresh latent state - each
`gan.reconstruct` call re-draws z0 and zeroes the momentum (reference :119).  Use
`SharedReconstruction` to let `predictions` and `diff_op` share one projection per batch, as
they share one graph node in the reference.
"""
from __future__ import annotations

import math
import warnings

import numpy as np
import torch


class _ArgsWrapper(object):
    """dict -> attribute shim (cleverhans/utils.py:17-29)."""

    def __init__(self, args):
        for k, v in (args.items() if isinstance(args, dict) else vars(args).items()):
            setattr(self, k, v)

    def __getattr__(self, name):
        return None


class PerBatchMemo(object):
    """fn(x) evaluated once per batch OBJECT: a second call with the same tensor (same object, not modified in place
    since) returns the remembered value.  The batch is kept referenced, so a later tensor cannot be mistaken for it by
    re-using its id()."""

    def __init__(self, fn):
        self.fn = fn
        self._x, self._version, self._val = None, None, None

    def __call__(self, x):
        version = getattr(x, "_version", None)
        if self._x is not x or self._version != version:
            self._val = self.fn(x)
            self._x, self._version = x, version
        return self._val


class SharedReconstruction(object):
    """Memoises gan.reconstruct per input batch object so that several per-batch callables
    (predictions, diff_op) observe the *same* projection, like the shared `reconstructed`
    tensor of reference blackbox.py:565-578."""

    def __init__(self, gan, **kwargs):
        self.gan, self.kwargs = gan, kwargs
        self._memo = PerBatchMemo(lambda x: self.gan.reconstruct(x, **self.kwargs))

    def __call__(self, x):
        return self._memo(x)


def _to_device_batch(arr, device):
    t = torch.as_tensor(np.ascontiguousarray(arr)) if not isinstance(arr, torch.Tensor) else arr
    return t.to(device=device, dtype=torch.float32, non_blocking=True)


def model_eval_gan(sess, images, labels, predictions=None, predictions_rec=None, test_images=None,
                   test_labels=None, feed=None, args=None, model=None, diff_op=None, device=None):
    """Accuracy of `predictions` on (test_images, test_labels) plus the reconstruction
    differences used for attack detection.  Returns `(accuracy, roc_info)` or
    `(accuracy, accuracy_rec, roc_info)` with `roc_info = [labels, preds, diffs]`
    (reference :175-179)."""
    args = _ArgsWrapper(args or {})
    assert args.batch_size, "Batch size was not given in args dict"
    if test_images is None or test_labels is None:
        raise ValueError("X_test argument and Y_test argument must be supplied.")
    if model is None and predictions is None:
        raise ValueError("One of model argument or predictions argument must be supplied.")
    if model is not None:
        warnings.warn("model argument is deprecated. Switch to predictions argument.")
        if predictions is None:
            predictions = model
        else:
            raise ValueError("Exactly one of model argument and predictions argument should be specified.")
    if not callable(predictions):
        raise TypeError("predictions must be a callable evaluated per batch (see module docstring)")
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")

    diffs, all_labels, preds = [], [], []
    accuracy, accuracy_rec = 0.0, 0.0
    n = len(test_images)
    nb_batches = int(math.ceil(float(n) / args.batch_size))
    assert nb_batches * args.batch_size >= n
    end = 0
    for batch in range(nb_batches):
        start = batch * args.batch_size
        end = min(n, start + args.batch_size)           # ragged last batch (reference :126-128)
        x = _to_device_batch(test_images[start:end], device)
        y = np.asarray(test_labels[start:end].cpu() if isinstance(test_labels, torch.Tensor) else test_labels[start:end])
        cur_labels = np.argmax(y, axis=-1)
        logits = predictions(x)
        cur_preds = torch.argmax(logits, dim=-1).cpu().numpy()
        accuracy += float(np.sum(cur_labels == cur_preds))
        if diff_op is not None:
            d = diff_op(x)
            diffs.append(np.atleast_1d(d.detach().cpu().numpy() if isinstance(d, torch.Tensor) else np.asarray(d)))
        if predictions_rec is not None:
            preds_rec = torch.argmax(predictions_rec(x), dim=-1).cpu().numpy()
            accuracy_rec += float(np.sum(cur_labels == preds_rec))
        all_labels.append(cur_labels)
        preds.append(cur_preds)
    assert end >= n
    accuracy /= n
    accuracy_rec /= n
    preds = np.concatenate(preds)
    all_labels = np.concatenate(all_labels)
    if diff_op is not None:
        diffs = np.concatenate(diffs)
    roc_info = [all_labels, preds, diffs]
    if predictions_rec is not None:
        return accuracy, accuracy_rec, roc_info
    return accuracy, roc_info
