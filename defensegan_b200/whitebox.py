"""White-box attacks (FGSM, RAND+FGSM) on a classifier with no defense, adversarial training or Defense-GAN in front
of it: the caller the reference ships as `whitebox.py`, in PyTorch around `gan.reconstruct`.

    acc_adv, _, roc_info = whitebox(gan, data=SplitData(...), defense_type='defense_gan', attack_type='fgsm', ...)
    python -m defensegan_b200.whitebox --cfg <gan cfg> --defense_type defense_gan --attack_type fgsm --model A ...

Function name, arguments, defaults and return value follow the reference (`whitebox` :56-233; result files :239-342;
flags :362-392).  Three things differ and are stated here rather than hidden:

* **Gradient through the projection.**  With `defense_type='defense_gan'` the reference makes the projection layer 0 of
  the classifier and lets FGSM differentiate "through" it (:185-201).  Its graph has no differentiable path from the
  image to the projection's output (z is a variable updated by an optimiser inside a while-loop) and TF1
  cannot be run here to see what `tf.gradients` made of that, so the behaviour is defined here instead:
  `rec_grad='straight_through'` (default) evaluates the classifier's input-gradient AT the reconstruction and passes
  it to the image unchanged (d rec / d x := I, the usual BPDA treatment of a projection defense);
  `rec_grad='classifier'` attacks the bare classifier (the reference's own `fprop(..., no_rec=True)` view).
  Either way the adversarial examples are then scored through the real projection.
* `attack_type='cw'` (Carlini-Wagner) is outside this package (out of scope) and is rejected by name.
* The reference overwrites the original test images with `get_cached_gan_data(..., orig_data_flag=True)` and evaluates
  clean accuracy on the (possibly reconstructed) first copy (:98-104,126-137); the two roles are explicit arguments here:
  `data` (what the classifier trains on and its clean evaluation set) and `attack_data` (original test images the
  attack starts from; default: `data`'s test split).
"""
from __future__ import annotations

import os
import sys
from typing import Optional

import numpy as np
import torch

from .blackbox import _clip_min, _parse, _pick_device
from .models.gan import dataset_gan_dict
from .utils import attacks
from .utils.experiment import (Flags, SplitData, get_cached_gan_data, set_test_time_rec_params, unique_result_path,
                               write_results)
from .utils.gan_defense import PerBatchMemo, model_eval_gan
from .utils.network_builder import model_dict

__all__ = ["whitebox", "main"]

_WHITEBOX_MODELS = ("A", "B", "C", "D", "E", "F")       # whitebox.py:121


def _through_projection_attack(model, x, rec, rec_grad, **fgsm_par):
    """FGSM on classifier(rec(x)) with the projection's Jacobian defined as in the module docstring."""
    if rec_grad == "classifier":
        return attacks.fgm(lambda u: model.get_logits(u, no_rec=True), x, **fgsm_par)
    if rec_grad != "straight_through":
        raise ValueError("rec_grad must be 'straight_through' or 'classifier'")
    base = rec(x).detach().to(x.device)
    # u = x + (rec(x) - x) held constant: value rec(x), gradient identity
    return attacks.fgm(lambda u: model.get_logits(u + (base - x), no_rec=True), x, **fgsm_par)


def whitebox(gan, rec_data_path=None, batch_size=128, learning_rate=0.001, nb_epochs=10, eps=0.3,
             online_training=False, test_on_dev=True, attack_type="fgsm", defense_type="gan", num_tests=-1,
             num_train=-1, data: Optional[SplitData] = None, attack_data=None, flags: Optional[Flags] = None,
             rec_grad="straight_through", device=None):
    """Trains classifier `flags.model` on `data`, then attacks it.  Returns `(accuracy on adversarial examples, 0,
    roc_info or None)`; with `attack_type=None` `(clean training accuracy, 0, None)` (whitebox.py:56-233).
    `flags` carries what the reference reads from FLAGS inside the function: model, defense_type, attack_type,
    fgsm_eps_tr, same_init, alpha, train_on_recs."""
    flags = flags if flags is not None else Flags("whitebox", defense_type=defense_type, attack_type=attack_type)
    device = _pick_device(device)
    defense = flags.defense_type if flags.defense_type is not None else defense_type
    if defense == "defense_gan":
        assert gan is not None
        if flags.train_on_recs:
            assert rec_data_path is not None or online_training
    if attack_type == "cw" or flags.attack_type == "cw":
        raise ValueError("attack_type 'cw' is not part of this package (FGSM and RAND+FGSM are)")

    from_cache = data is None
    if from_cache:
        data = get_cached_gan_data(gan, test_on_dev, flags=flags)
    train_images, train_labels = np.asarray(data.train_images), np.asarray(data.train_labels)
    clean_images, clean_labels = np.asarray(data.test_images), np.asarray(data.test_labels)
    if attack_data is None:
        if from_cache:                       # the cached ORIGINAL test images, whatever the classifier trained on
            orig = get_cached_gan_data(gan, test_on_dev, orig_data_flag=True, flags=flags)
            attack_data = (orig.test_images, orig.test_labels)
        else:
            attack_data = (clean_images, clean_labels)
    test_images, test_labels = np.asarray(attack_data[0]), np.asarray(attack_data[1])
    if num_tests > 0:
        test_images, test_labels = test_images[:num_tests], test_labels[:num_tests]
        clean_images, clean_labels = clean_images[:num_tests], clean_labels[:num_tests]
    if num_train > 0:
        train_images, train_labels = train_images[:num_train], train_labels[:num_train]

    if flags.model not in _WHITEBOX_MODELS:
        raise KeyError("model must be one of %s" % (_WHITEBOX_MODELS,))
    model = model_dict[flags.model](input_shape=[None] + list(train_images.shape[1:]), nb_classes=train_labels.shape[1])
    model.to(device)
    eval_params = {"batch_size": batch_size}

    def evaluate():                          # after every training epoch (:126-137)
        acc = attacks.model_eval(model, clean_images, clean_labels, eval_params, device=device)
        print("Test accuracy on legitimate examples: %0.4f" % acc)

    rng = np.random.RandomState([11, 24, 1990])
    torch.manual_seed(11241990)
    lo = _clip_min(gan)
    adv_fn = None
    if defense == "adv_tr":
        adv_fn = lambda x: attacks.fgm(model, x, eps=flags.fgsm_eps_tr, clip_min=lo, clip_max=1.0)
    attacks.model_train(model, train_images, train_labels,
                        {"nb_epochs": nb_epochs, "batch_size": batch_size, "learning_rate": learning_rate},
                        predictions_adv=adv_fn, rng=rng, device=device, evaluate=evaluate)
    acc_train = attacks.model_eval(model, train_images, train_labels, eval_params, device=device)
    print("[#] Accuracy on clean examples {}".format(acc_train))
    if attack_type is None:
        return acc_train, 0, None

    kind = flags.attack_type if flags.attack_type not in (None, "none") else attack_type
    if defense == "defense_gan":
        z_init_val = None
        if flags.same_init:                 # one fixed z0 for every batch (:187-190; batches must then be full-size)
            z_init_val = torch.from_numpy(np.random.randn(batch_size * gan.rec_rr, gan.latent_dim).astype(np.float32))
        model.add_rec_model(gan, z_init_val, batch_size)
    if "rand" in kind:                      # RAND+FGSM: a random sign step of size alpha first (:199-203)
        test_images = np.clip(test_images + flags.alpha * np.sign(np.random.randn(*test_images.shape)), lo, 1.0)
        eps -= flags.alpha
    if "fgsm" not in kind:
        raise ValueError("attack_type must be 'fgsm' or 'rand+fgsm', got %r" % (kind,))
    fgsm_par = {"eps": eps, "ord": np.inf, "clip_min": lo, "clip_max": 1.0}
    model.eval()

    if defense == "defense_gan":
        rec_layer = model._rec_layer
        # predictions and diff_op of one batch see the same adversarial examples
        adv_batch = PerBatchMemo(lambda x: _through_projection_attack(model, x, rec_layer.fprop, rec_grad, **fgsm_par))

        def predictions(x):
            with torch.no_grad():
                return model.get_logits(adv_batch(x))          # layer 0 projects the adversarial batch

        def diff_op(x):                     # whitebox.py:219: mean((adv_x - x)^2) per image
            return ((adv_batch(x) - x) ** 2).mean(dim=tuple(range(1, x.dim())))

        acc_adv, roc_info = model_eval_gan(None, None, None, predictions=predictions, test_images=test_images,
                                           test_labels=test_labels, args=eval_params, diff_op=diff_op, device=device)
        print("Test accuracy on adversarial examples: %0.4f\n" % acc_adv)
        return acc_adv, 0, roc_info
    acc_adv = attacks.model_eval(lambda x: model(attacks.fgm(model, x, **fgsm_par)), test_images, test_labels,
                                 eval_params, device=device)
    print("Test accuracy on adversarial examples: %0.4f\n" % acc_adv)
    return acc_adv, 0, None


# ------------------------------------------------------------------------------------------------------------------
# command line (whitebox.py:239-395)
# ------------------------------------------------------------------------------------------------------------------
def _results_dir_filename(gan, flags):
    results_dir = os.path.join("results", "whitebox_{}_{}".format(flags.defense_type, gan.dataset_name))
    if flags.rec_path and flags.defense_type == "defense_gan":
        results_dir = gan.checkpoint_dir.replace("output", "results")
        # the reference's format string prints (rec_rr, rec_lr, rec_iters) under the labels Iter / RR / LR (:318-324);
        # the labels are matched to their values here
        name = "Iter={:d}_RR={:d}_LR={:.4f}_defense=gan".format(gan.rec_iters, gan.rec_rr, gan.rec_lr)
        if not flags.train_on_recs:
            name = "orig_" + name
    elif flags.defense_type == "adv_tr":
        name = "advTrEps={:.2f}".format(flags.fgsm_eps_tr)
    else:
        name = "nodefense_"
    if flags.num_tests > -1:
        name = "numtest={}_".format(flags.num_tests) + name
    if flags.num_train > -1:
        name = "numtrain={}_".format(flags.num_train) + name
    return results_dir, "model={}_".format(flags.model) + name + "attack={}.txt".format(flags.attack_type)


def main(cfg, argv=None, flags: Optional[Flags] = None, data=None, attack_data=None):
    flags = flags if flags is not None else Flags("whitebox")
    gan = dataset_gan_dict[cfg["DATASET_NAME"] if "DATASET_NAME" in cfg else flags.dataset_name](cfg=cfg, test_mode=True)
    gan.load_generator()
    set_test_time_rec_params(gan, flags, cfg)
    results_dir, file_name = _results_dir_filename(gan, flags)
    path = unique_result_path(os.path.join(results_dir, flags.results_dir or ""), file_name)
    acc = whitebox(gan, rec_data_path=flags.rec_path, batch_size=int(cfg.get("BATCH_SIZE", 128)),
                   learning_rate=flags.learning_rate, nb_epochs=flags.nb_epochs, eps=flags.fgsm_eps,
                   online_training=flags.online_training, test_on_dev=flags.test_on_dev,
                   defense_type=flags.defense_type, num_tests=flags.num_tests, attack_type=flags.attack_type,
                   num_train=flags.num_train, data=data, attack_data=attack_data, flags=flags)
    write_results(path, [acc[0], acc[1]], acc[2])
    return acc


if __name__ == "__main__":
    if len(sys.argv) == 1:
        print("usage: python -m defensegan_b200.whitebox --cfg <path> [--alpha a] [--<flag> <value> ...]")
        sys.exit(1)
    _cfg, _flags = _parse(sys.argv[1:], "whitebox")
    main(_cfg, flags=_flags)
