"""Black-box (substitute-model) attack against a classifier, with and without Defense-GAN in front of it: the caller
of the projection loop that the reference ships as `blackbox.py` (after arxiv.org/abs/1602.02697 and the cleverhans
MNIST tutorial), in PyTorch around `gan.reconstruct`.

    accuracies = blackbox(gan, data=SplitData(...), defense_type='defense_gan', bb_model='A', sub_model='B', ...)
    python -m defensegan_b200.blackbox --cfg <gan cfg> --defense_type defense_gan --bb_model A --sub_model B ...

Same function names, arguments, defaults and result keys as the reference (`prep_bbox` :65-140, `train_sub` :143-213,
`blackbox` :370-593, result files :596-699, flags :723-759).  Graph-mode -> eager mapping: a "predictions tensor" is a
callable `x -> logits`; `sess` / placeholders disappear; every `gan.reconstruct` call draws fresh z0 and zero momentum,
which is what the reference's `sess.run(tf.local_variables_initializer())` before each evaluation does.

The oracle the adversary queries is `classifier(gan.reconstruct(x))` whenever a GAN is given - for every defense type,
as in the reference (:505-517) - and the transferred adversarial examples are scored through
`model_eval_gan` (accuracy + the per-image reconstruction error used for attack detection).
"""
from __future__ import annotations

import argparse
import os
import sys
from typing import Callable, Optional

import numpy as np
import torch

from .models.gan import dataset_gan_dict
from .utils import attacks
from .utils.config import add_flags, load_config
from .utils.experiment import (Flags, SplitData, add_script_flags, get_cached_gan_data, set_test_time_rec_params,
                               unique_result_path, write_results)
from .utils.gan_defense import PerBatchMemo, SharedReconstruction, model_eval_gan
from .utils.network_builder import model_dict

__all__ = ["prep_bbox", "train_sub", "jacobian_augmentation", "blackbox", "main"]


def _pick_device(device=None):
    if device is not None:
        return torch.device(device)
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")


def _clip_min(gan):
    """Pixel range lower bound: CelebA images live in [-1, 1] (blackbox.py:113-115)."""
    return -1.0 if (gan is not None and "celeba" in str(gan.dataset_name)) else 0.0


class _Projected(torch.nn.Module):
    """classifier(gan.reconstruct(x)) with no gradient into the projection (`tf.stop_gradient`, blackbox.py:93-95)."""

    def __init__(self, classifier, gan, reconstructor_id=0):
        super().__init__()
        self.classifier = classifier
        self._gan = [gan]                  # not a sub-module
        self.reconstructor_id = reconstructor_id

    def forward(self, x):
        rec = self._gan[0].reconstruct(x, batch_size=None, reconstructor_id=self.reconstructor_id)
        return self.classifier(rec.detach().to(x.device))


def prep_bbox(model, images_train, labels_train, images_test, labels_test, nb_epochs, batch_size, learning_rate, rng,
              gan=None, adv_training=False, fgsm_eps_tr=0.15, device=None):
    """Trains the "remote" classifier the adversary will only be able to query (blackbox.py:65-140).
    `gan` given: every training batch is projected first (online training on reconstructions).  `adv_training`:
    FGSM(eps = fgsm_eps_tr) adversarial training.  Returns (model, predictions callable, clean test accuracy)."""
    device = _pick_device(device)
    model.to(device)
    predictions = _Projected(model, gan) if gan is not None else model
    adv_fn = None
    if adv_training:
        lo = _clip_min(gan)
        adv_fn = lambda x: attacks.fgm(model, x, eps=fgsm_eps_tr, ord=np.inf, clip_min=lo, clip_max=1.0)
    train_params = {"nb_epochs": nb_epochs, "batch_size": batch_size, "learning_rate": learning_rate}
    attacks.model_train(predictions, images_train, labels_train, train_params, predictions_adv=adv_fn, rng=rng,
                        device=device)
    accuracy = attacks.model_eval(predictions, images_test, labels_test, {"batch_size": batch_size}, device=device)
    print("Test accuracy of black-box on legitimate test examples: " + str(accuracy))
    return model, predictions, accuracy


def _probabilities(model, x):
    return model.get_probs(x) if hasattr(model, "get_probs") else torch.softmax(model(x), dim=-1)


def jacobian_augmentation(model_sub, X_sub_prev, Y_sub, lmbda, batch_size=128, device=None):
    """Jacobian-based dataset augmentation (cleverhans attacks_tf.py:240-257,551-597): every point x with oracle label y
    gets the companion x + lmbda * sign(d p_y(x) / dx), p = the substitute's class probabilities.  Returns the doubled
    set (old points first), to be labelled by the oracle.  Samples are independent, so they are differentiated a batch
    at a time instead of one `sess.run` per point."""
    X_sub_prev = np.asarray(X_sub_prev, dtype=np.float32)
    Y_sub = np.asarray(Y_sub).astype(np.int64)
    assert len(X_sub_prev) == len(Y_sub)
    device = _pick_device(device)
    was_training = model_sub.training
    model_sub.eval()
    new = np.empty_like(X_sub_prev)
    for s in range(0, len(X_sub_prev), batch_size):
        x = torch.from_numpy(X_sub_prev[s:s + batch_size]).to(device).requires_grad_(True)
        y = torch.from_numpy(Y_sub[s:s + batch_size]).to(device)
        p = _probabilities(model_sub, x)
        assert p.shape[1] >= int(Y_sub.max()) + 1
        grad, = torch.autograd.grad(p.gather(1, y[:, None]).sum(), x)
        new[s:s + batch_size] = (x.detach() + lmbda * torch.sign(grad)).cpu().numpy()
    model_sub.train(was_training)
    return np.vstack([X_sub_prev, new])


def train_sub(bbox_preds: Callable, X_sub, Y_sub, nb_classes, nb_epochs_s, batch_size, learning_rate, data_aug, lmbda,
              rng, substitute_model=None, device=None):
    """The adversary's substitute (blackbox.py:143-213): `data_aug` rounds of {train on the current set; unless last:
    double the set by Jacobian augmentation and let the oracle `bbox_preds` label the new half - arg-max only, the
    adversary sees labels, not probabilities}.  Returns the trained substitute."""
    device = _pick_device(device)
    model_sub = substitute_model.to(device)
    X_sub = np.asarray(X_sub, dtype=np.float32)
    Y_sub = np.asarray(Y_sub).astype(np.int64)
    train_params = {"nb_epochs": nb_epochs_s, "batch_size": batch_size, "learning_rate": learning_rate}
    for rho in range(data_aug):
        print("Substitute training epoch #" + str(rho))
        onehot = np.zeros((len(Y_sub), nb_classes), np.float32)
        onehot[np.arange(len(Y_sub)), Y_sub] = 1.0
        attacks.model_train(model_sub, X_sub, onehot, train_params, rng=rng, device=device)
        if rho == data_aug - 1:
            break
        print("Augmenting substitute training data.")
        n_old = len(X_sub)
        X_sub = jacobian_augmentation(model_sub, X_sub, Y_sub, lmbda, batch_size=batch_size, device=device)
        print("Labeling substitute training data.")
        with torch.no_grad():
            answers = attacks.batch_eval(bbox_preds, X_sub[n_old:], batch_size, device)
        Y_sub = np.concatenate([Y_sub, answers.argmax(dim=1).numpy().astype(np.int64)])
    return model_sub


def blackbox(gan, rec_data_path=None, batch_size=128, learning_rate=0.001, nb_epochs=10, holdout=150, data_aug=6,
             nb_epochs_s=10, lmbda=0.1, online_training=False, train_on_recs=False, test_on_dev=True,
             defense_type="none", data: Optional[SplitData] = None, rec_data: Optional[SplitData] = None,
             flags: Optional[Flags] = None, device=None):
    """The whole experiment (blackbox.py:370-593).  Returns {'bbox': clean accuracy of the black box, 'sub': 0,
    'bbox_on_sub_adv_ex': its accuracy on FGSM examples crafted on the substitute[, 'roc_info': [labels, preds,
    reconstruction errors]]}.

    `data` (original images) / `rec_data` (cached reconstructions, used when `rec_data_path` is set and training is not
    online) may be given directly; otherwise they are read from the dataset caches (`get_cached_gan_data`).
    `flags` carries what the reference reads from FLAGS inside the function: bb_model, sub_model, fgsm_eps, fgsm_eps_tr,
    num_tests, debug."""
    flags = flags if flags is not None else Flags("blackbox")
    device = _pick_device(device)
    accuracies = {}
    defense_type = defense_type or ""
    gan_defense = defense_type == "defense_gan" and gan is not None
    adv_training = "adv_tr" in defense_type

    if data is None:
        data = get_cached_gan_data(gan, test_on_dev, orig_data_flag=True, flags=flags)
    train_images, train_labels = np.asarray(data.train_images), np.asarray(data.train_labels)
    test_images, test_labels = np.asarray(data.test_images), np.asarray(data.test_labels)
    nb_classes = train_labels.shape[1]
    input_shape = [None] + list(train_images.shape[1:])
    bb_model = model_dict[flags.bb_model](input_shape=input_shape, nb_classes=nb_classes)
    sub_model = model_dict[flags.sub_model](input_shape=input_shape, nb_classes=nb_classes)

    if flags.debug:
        train_images, train_labels = train_images[:20 * batch_size], train_labels[:20 * batch_size]

    # the adversary's seed set comes off the front of the test split; the evaluation uses what is left of the first
    # num_tests samples (:427-437)
    images_sub = test_images[:holdout]
    labels_sub = np.argmax(test_labels[:holdout], axis=1)
    if flags.num_tests and flags.num_tests > 0:
        test_images, test_labels = test_images[:flags.num_tests], test_labels[:flags.num_tests]
    test_images, test_labels = test_images[holdout:], test_labels[holdout:]

    rng = np.random.RandomState([11, 24, 1990])
    torch.manual_seed(11241990)

    # what the black box is trained / tested on (:455-477)
    bb_train = (train_images, train_labels, test_images, test_labels)
    train_gan = None
    if "gan" in defense_type:
        if online_training and not train_on_recs:
            train_gan = gan
        elif not online_training and rec_data_path:
            if rec_data is None:
                rec_data = get_cached_gan_data(gan, test_on_dev, orig_data_flag=False, flags=flags)
            bb_train = tuple(np.asarray(a) for a in rec_data)
        else:
            assert not train_on_recs
        if flags.debug:
            bb_train = (bb_train[0][:20 * batch_size], bb_train[1][:20 * batch_size]) + bb_train[2:]
    model, _, accuracies["bbox"] = prep_bbox(bb_model, bb_train[0], bb_train[1], bb_train[2], bb_train[3], nb_epochs,
                                             batch_size, learning_rate, rng=rng, gan=train_gan,
                                             adv_training=adv_training, fgsm_eps_tr=flags.fgsm_eps_tr, device=device)

    print("Training the substitute model.")
    oracle = _Projected(model, gan, reconstructor_id=1) if gan is not None else model
    oracle.eval()
    model_sub = train_sub(oracle, images_sub, labels_sub, nb_classes, nb_epochs_s, batch_size, learning_rate, data_aug,
                          lmbda, rng=rng, substitute_model=sub_model, device=device)
    accuracies["sub"] = 0

    fgsm_par = {"eps": flags.fgsm_eps, "ord": np.inf, "clip_min": _clip_min(gan), "clip_max": 1.0}
    model.eval()
    model_sub.eval()
    craft = lambda x: attacks.fgm(model_sub, x, **fgsm_par)
    eval_params = {"batch_size": batch_size}
    if gan_defense:
        # one projection of the adversarial batch feeds both the classifier and the detection statistic (:565-578)
        adv_batch, rec = PerBatchMemo(craft), SharedReconstruction(gan, reconstructor_id=4)

        def predictions(x):
            with torch.no_grad():
                return model(rec(adv_batch(x)).to(x.device))

        def diff_op(x):
            adv = adv_batch(x)
            return ((adv - rec(adv).to(adv.device)) ** 2).mean(dim=tuple(range(1, adv.dim())))

        acc, roc_info = model_eval_gan(None, None, None, predictions=predictions, test_images=test_images,
                                       test_labels=test_labels, args=eval_params, diff_op=diff_op, device=device)
        accuracies["bbox_on_sub_adv_ex"] = acc
        accuracies["roc_info"] = roc_info
    else:
        acc = attacks.model_eval(lambda x: model(craft(x)), test_images, test_labels, eval_params, device=device)
        accuracies["bbox_on_sub_adv_ex"] = acc
    print("Test accuracy of oracle on adversarial examples generated using the substitute: " + str(acc))
    return accuracies


# ------------------------------------------------------------------------------------------------------------------
# command line (blackbox.py:596-762)
# ------------------------------------------------------------------------------------------------------------------
def _results_dir_filename(gan, flags):
    name = "sub={:d}_eps={:.2f}.txt".format(flags.data_aug, flags.fgsm_eps)
    results_dir = os.path.join("results", "{}_{}".format(flags.defense_type, gan.dataset_name))
    if flags.rec_path and flags.defense_type == "defense_gan":
        results_dir = gan.checkpoint_dir.replace("output", "results")
        name = "teRR={:d}_teLR={:.4f}_teIter={:d}_".format(gan.rec_rr, gan.rec_lr, gan.rec_iters) + name
        if not flags.train_on_recs:
            name = "orig_" + name
    elif flags.defense_type == "adv_tr":
        name = "sub={:d}_trEps={:.2f}_eps={:.2f}.txt".format(flags.data_aug, flags.fgsm_eps_tr, flags.fgsm_eps)
    if flags.num_tests > -1:
        name = "numtest={}_".format(flags.num_tests) + name
    if flags.num_train > -1:
        name = "numtrain={}_".format(flags.num_train) + name
    return results_dir, "bbModel={}_subModel={}_".format(flags.bb_model, flags.sub_model) + name


def main(cfg, argv=None, flags: Optional[Flags] = None, data=None, rec_data=None):
    flags = flags if flags is not None else Flags("blackbox")
    gan = dataset_gan_dict[cfg["DATASET_NAME"] if "DATASET_NAME" in cfg else flags.dataset_name](cfg=cfg, test_mode=True)
    gan.load_generator()
    set_test_time_rec_params(gan, flags, cfg)
    results_dir, file_name = _results_dir_filename(gan, flags)
    path = unique_result_path(os.path.join(results_dir, flags.results_dir or ""), file_name)
    acc = blackbox(gan, rec_data_path=flags.rec_path, batch_size=int(cfg.get("BATCH_SIZE", 128)),
                   learning_rate=flags.learning_rate, nb_epochs=flags.nb_epochs, holdout=flags.holdout,
                   data_aug=flags.data_aug, nb_epochs_s=flags.nb_epochs_s, lmbda=flags.lmbda,
                   online_training=flags.online_training, train_on_recs=flags.train_on_recs,
                   test_on_dev=flags.test_on_dev, defense_type=flags.defense_type, data=data, rec_data=rec_data,
                   flags=flags)
    write_results(path, [acc[k] for k in ("bbox", "sub", "bbox_on_sub_adv_ex")], acc.get("roc_info"))
    return acc


def _parse(argv, script):
    first = argparse.ArgumentParser(add_help=False)
    first.add_argument("--cfg", required=True, help="Config file")
    known, _ = first.parse_known_args(argv)
    cfg = load_config(known.cfg)
    parser = argparse.ArgumentParser()
    parser.add_argument("--cfg", required=True, help="Config file")
    add_script_flags(parser, script)
    add_flags(parser, cfg)                      # every cfg key is a flag too (utils/config.py)
    ns = parser.parse_args(argv)
    for k, v in vars(ns).items():
        if k.upper() in cfg:
            cfg[k.upper()] = v
    return cfg, Flags(script, **vars(ns))


if __name__ == "__main__":
    if len(sys.argv) == 1:
        print("usage: python -m defensegan_b200.blackbox --cfg <path> [--<flag> <value> ...]")
        sys.exit(1)
    _cfg, _flags = _parse(sys.argv[1:], "blackbox")
    main(_cfg, flags=_flags)
