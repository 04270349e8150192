"""f3: the callers of the projection loop - blackbox() / whitebox() experiment drivers, Jacobian
augmentation, the cached-dataset readers and result files (reference blackbox.py, whitebox.py).
CPU tests use a stand-in projector (`TemplateGAN`: nearest class template, a projection onto a 4-point "manifold");
the GPU test at the end drives the real CUDA projection through the same code."""
import os
import pickle

import numpy as np
import pytest
import torch

from defensegan_b200 import blackbox as BB
from defensegan_b200 import whitebox as WB
from defensegan_b200.utils import experiment as E
from defensegan_b200.utils import network_builder as nb

N_CLASSES = 4


def _templates():
    """Four well-separated 28x28 patterns (quadrant blocks)."""
    t = np.zeros((N_CLASSES, 28, 28, 1), np.float32)
    t[0, :14, :14], t[1, :14, 14:], t[2, 14:, :14], t[3, 14:, 14:] = 1.0, 1.0, 1.0, 1.0
    return t


def _dataset(n, seed, noise=0.15):
    rng = np.random.RandomState(seed)
    y = rng.randint(0, N_CLASSES, size=n)
    x = np.clip(_templates()[y] * 0.8 + noise * rng.randn(n, 28, 28, 1), 0.0, 1.0).astype(np.float32)
    return x, E.convert_to_onehot(np.concatenate([y, [N_CLASSES - 1]]))[:n]


class TemplateGAN(object):
    """Stand-in for DefenseGANBase on the CPU: reconstruct(x) = the nearest of four templates (scaled like the data)."""
    dataset_name = "mnist"
    rec_rr, rec_lr, rec_iters, latent_dim = 2, 10.0, 5, 8
    image_dim = [28, 28, 1]
    checkpoint_dir = "output/gans/mnist"

    def __init__(self):
        self.calls = []
        self.t = torch.from_numpy(_templates() * 0.8)

    def reconstruct(self, images, batch_size=None, back_prop=True, reconstructor_id=0, z_init_val=None):
        self.calls.append((int(images.shape[0]), reconstructor_id))
        x = images.detach().cpu()
        d = ((x[:, None] - self.t[None]) ** 2).flatten(2).sum(-1)
        return self.t[d.argmin(dim=1)].to(images.device)


def _data(n_train=256, n_test=160):
    tr = _dataset(n_train, 0)
    te = _dataset(n_test, 1)
    return E.SplitData(tr[0], tr[1], te[0], te[1])


def test_convert_to_onehot_and_flags():
    oh = E.convert_to_onehot([2, 0, 1, 2])
    assert oh.dtype == np.float32 and oh.shape == (4, 3) and oh.argmax(1).tolist() == [2, 0, 1, 2]
    f = E.Flags("blackbox", bb_model="A")
    assert (f.num_tests, f.bb_model, f.sub_model, f.holdout, f.fgsm_eps, f.no_such_flag) == (2000, "A", "E", 150, 0.3, None)
    assert E.Flags("whitebox").num_tests == -1


def test_jacobian_augmentation_matches_per_sample_definition():
    """cleverhans attacks_tf.py:551-597: one gradient evaluation per point, sign, step lmbda, stacked under the old set."""
    torch.manual_seed(0)
    m = nb.model_e(nb_classes=N_CLASSES)
    x, y1 = _dataset(9, 3)
    y = y1.argmax(1)
    got = BB.jacobian_augmentation(m, x, y, lmbda=0.1, batch_size=4, device="cpu")
    assert got.shape == (18, 28, 28, 1) and np.array_equal(got[:9], x)
    m.eval()
    for i in range(9):
        xi = torch.from_numpy(x[i:i + 1]).requires_grad_(True)
        g, = torch.autograd.grad(m.get_probs(xi)[0, y[i]], xi)
        np.testing.assert_allclose(got[9 + i], (xi.detach() + 0.1 * torch.sign(g))[0].numpy(), atol=1e-6)


def test_train_sub_queries_the_oracle_for_labels_only():
    """blackbox.py:143-213: the set doubles data_aug - 1 times; only the NEW half is sent to the oracle each round."""
    torch.manual_seed(1)
    x, y1 = _dataset(20, 4)
    asked = []

    def oracle(xb):
        asked.append(int(xb.shape[0]))
        t = torch.from_numpy(_templates() * 0.8)
        d = ((xb.cpu()[:, None] - t[None]) ** 2).flatten(2).sum(-1)
        return -d                                      # "logits": nearest template wins

    sub = BB.train_sub(oracle, x, y1.argmax(1), N_CLASSES, nb_epochs_s=2, batch_size=16, learning_rate=0.01, data_aug=3,
                       lmbda=0.1, rng=np.random.RandomState(0), substitute_model=nb.model_e(nb_classes=N_CLASSES),
                       device="cpu")
    assert sum(asked) == 20 + 40 and max(asked) <= 16          # rounds 1 and 2 label 20 and 40 new points
    acc = (sub(torch.from_numpy(x)).argmax(1).numpy() == y1.argmax(1)).mean()
    assert acc > 0.9


def test_blackbox_pipeline_without_and_with_the_projection():
    data = _data()
    flags = E.Flags("blackbox", bb_model="E", sub_model="E", num_tests=-1, fgsm_eps=0.3)
    common = dict(batch_size=32, learning_rate=0.005, nb_epochs=4, holdout=40, data_aug=3, nb_epochs_s=4, lmbda=0.1,
                  data=data, flags=flags, device="cpu")
    gan = TemplateGAN()
    plain = BB.blackbox(gan, defense_type="none", **common)
    assert set(plain) == {"bbox", "sub", "bbox_on_sub_adv_ex"} and plain["sub"] == 0
    assert plain["bbox"] > 0.95
    # oracle answers for the substitute go through the projection whenever a GAN is given (blackbox.py:505-517)
    assert {rid for _, rid in gan.calls} == {1}
    gan2 = TemplateGAN()
    defended = BB.blackbox(gan2, defense_type="defense_gan", online_training=True, **common)
    labels, preds, diffs = defended["roc_info"]
    assert len(labels) == len(preds) == len(diffs) == 160 - 40           # holdout removed from the evaluation set
    assert np.all(diffs >= 0) and defended["bbox_on_sub_adv_ex"] == pytest.approx((labels == preds).mean())
    # training was online (reconstructor_id 0), labelling used id 1, the evaluation id 4 - once per evaluation batch
    ids = [rid for _, rid in gan2.calls]
    assert ids.count(4) == int(np.ceil(120 / 32.0)) and 0 in ids and 1 in ids
    # the projection snaps FGSM examples back to their template: the defended classifier must not lose to the plain one
    assert defended["bbox_on_sub_adv_ex"] >= plain["bbox_on_sub_adv_ex"] - 0.05
    assert defended["bbox_on_sub_adv_ex"] > 0.9


def test_blackbox_adversarial_training_and_cached_reconstructions():
    data = _data(128, 96)
    flags = E.Flags("blackbox", bb_model="E", sub_model="E", num_tests=80, fgsm_eps=0.2, fgsm_eps_tr=0.1)
    out = BB.blackbox(None, defense_type="adv_tr", batch_size=32, learning_rate=0.005, nb_epochs=2, holdout=30,
                      data_aug=2, nb_epochs_s=2, data=data, flags=flags, device="cpu")
    assert 0.0 <= out["bbox_on_sub_adv_ex"] <= 1.0 and "roc_info" not in out
    # cached reconstructions: the black box trains on rec_data instead of data when rec_data_path is set (:462-468)
    gan = TemplateGAN()
    rec = E.SplitData(gan.reconstruct(torch.from_numpy(data.train_images)).numpy(), data.train_labels,
                      gan.reconstruct(torch.from_numpy(data.test_images)).numpy(), data.test_labels)
    gan.calls.clear()
    out = BB.blackbox(gan, rec_data_path="output/gans/mnist/recs_rr2_lr10.00000_iters5", defense_type="defense_gan",
                      train_on_recs=True, batch_size=32, learning_rate=0.005, nb_epochs=2, holdout=30, data_aug=2,
                      nb_epochs_s=2, data=data, rec_data=rec, flags=flags, device="cpu")
    assert 0 not in [rid for _, rid in gan.calls]                # no online projection during training
    assert out["bbox"] > 0.9


def test_whitebox_defenses_and_attacks():
    data = _data(256, 96)
    kw = dict(batch_size=32, learning_rate=0.005, nb_epochs=4, data=data, device="cpu")
    f = lambda **v: E.Flags("whitebox", model="E", **v)
    acc_clean, zero, none = WB.whitebox(None, attack_type=None, defense_type="none", flags=f(defense_type="none"), **kw)
    assert acc_clean > 0.95 and zero == 0 and none is None
    acc_plain, _, roc = WB.whitebox(None, eps=0.3, attack_type="fgsm", defense_type="none",
                                    flags=f(defense_type="none", attack_type="fgsm"), **kw)
    assert roc is None and acc_plain < acc_clean
    acc_rand, _, _ = WB.whitebox(None, eps=0.3, attack_type="rand+fgsm", defense_type="adv_tr",
                                 flags=f(defense_type="adv_tr", attack_type="rand+fgsm", fgsm_eps_tr=0.1), **kw)
    assert 0.0 <= acc_rand <= 1.0
    for mode in ("straight_through", "classifier"):
        gan = TemplateGAN()
        acc_gan, _, roc = WB.whitebox(gan, eps=0.3, attack_type="fgsm", defense_type="defense_gan", rec_grad=mode,
                                      flags=f(defense_type="defense_gan", attack_type="fgsm"), **kw)
        labels, preds, diffs = roc
        assert len(diffs) == 96 and acc_gan == pytest.approx((labels == preds).mean())
        # FGSM moves every pixel that is not clipped by eps: the detection statistic is bounded by eps^2
        assert np.all(diffs <= 0.3 ** 2 + 1e-6) and diffs.max() > 0.01
        assert all(rid == 123 for _, rid in gan.calls)           # the projection is layer 0 of the classifier
        assert acc_gan >= acc_plain
    with pytest.raises(ValueError):
        WB.whitebox(None, attack_type="cw", defense_type="none", flags=f(defense_type="none", attack_type="cw"), **kw)


def test_straight_through_attack_gradient_is_taken_at_the_reconstruction():
    torch.manual_seed(3)
    m = nb.model_e(nb_classes=N_CLASSES)
    gan = TemplateGAN()
    m.add_rec_model(gan, None, 8)
    m.eval()
    x = torch.from_numpy(_dataset(8, 7)[0])
    adv = WB._through_projection_attack(m, x, m._rec_layer.fprop, "straight_through", eps=0.1, ord=np.inf,
                                        clip_min=0.0, clip_max=1.0)
    r = gan.reconstruct(x).requires_grad_(True)
    logits = m.get_logits(r, no_rec=True)
    y = torch.nn.functional.one_hot(logits.argmax(1), N_CLASSES).float()
    g, = torch.autograd.grad(-(y * torch.log_softmax(logits, -1)).sum(), r)
    assert torch.allclose(adv, (x + 0.1 * torch.sign(g)).clamp(0, 1))


def test_cached_dataset_readers_and_result_files(tmp_path, monkeypatch):
    """save_ds layout (two pickles in feats.pkl), per-image reconstruction pickles, rec-path parsing, result naming."""
    monkeypatch.chdir(tmp_path)
    x_tr, y_tr = _dataset(12, 0)
    x_te, y_te = _dataset(6, 1)
    for split, (x, y) in (("train", (x_tr, y_tr)), ("test", (x_te, y_te))):
        d = os.path.join(E.orig_data_path("mnist"), split)
        os.makedirs(d)
        with open(os.path.join(d, "feats.pkl"), "wb") as f:
            pickle.dump(x, f, pickle.HIGHEST_PROTOCOL)
            pickle.dump(y.argmax(1), f, pickle.HIGHEST_PROTOCOL)
    gan = TemplateGAN()
    got = E.get_cached_gan_data(gan, True, orig_data_flag=True, flags=E.Flags("blackbox"))
    assert np.array_equal(got.train_images, x_tr) and np.array_equal(got.test_labels.argmax(1), y_te.argmax(1))
    with pytest.raises(IOError):
        E.get_cached_gan_data(gan, False, orig_data_flag=True, flags=E.Flags("blackbox"))      # no 'dev' split cached

    # reconstructions come from the model's reconstruct_dataset when originals are not asked for
    gan.reconstruct_dataset = lambda max_num_load=-1: {"train": [x_tr * 0.5, y_tr.argmax(1), x_tr],
                                                       "test": [x_te * 0.5, y_te.argmax(1), x_te]}
    got = E.get_cached_gan_data(gan, True, flags=E.Flags("blackbox", train_on_recs=True, defense_type="defense_gan"))
    assert np.array_equal(got.train_images, x_tr * 0.5) and got.train_labels.shape == (12, y_tr.argmax(1).max() + 1)

    # per-image pickles, labels parsed from the file names (CelebA branch, blackbox.py:249-259)
    rec_dir = tmp_path / "recs_rr3_lr0.50000_iters7"
    os.makedirs(rec_dir / "train" / "pickles")
    for i in range(5):
        with open(rec_dir / "train" / "pickles" / "rec_{:07d}_l{}.pkl".format(i, i % 2), "wb") as f:
            pickle.dump(x_tr[i], f)
    imgs, labels = E.get_pickle_split(str(rec_dir), "train", [28, 28, 1])
    assert np.array_equal(imgs, x_tr[:5]) and labels.tolist() == [0, 1, 0, 1, 0]

    flags = E.Flags("blackbox", defense_type="defense_gan", rec_path=str(rec_dir) + "/train", data_aug=6, fgsm_eps=0.3,
                    bb_model="A", sub_model="B", num_tests=2000)
    E.set_test_time_rec_params(gan, flags)
    assert (gan.rec_rr, gan.rec_lr, gan.rec_iters) == (3, 0.5, 7)
    rd, name = BB._results_dir_filename(gan, flags)
    assert rd == "results/gans/mnist"
    assert name == "bbModel=A_subModel=B_numtest=2000_orig_teRR=3_teLR=0.5000_teIter=7_sub=6_eps=0.30.txt"
    p0 = E.unique_result_path(rd, name)
    E.write_results(p0, [0.9, 0, 0.5], roc_info=[np.arange(3), np.arange(3), np.zeros(3)])
    assert os.path.basename(p0).startswith("0_") and open(p0).read() == "0.9 0 0.5 \n"
    assert os.path.exists(p0.replace(".txt", "_roc.pkl"))
    assert os.path.basename(E.unique_result_path(rd, name)).startswith("1_")
    wf = E.Flags("whitebox", defense_type="adv_tr", attack_type="fgsm", model="C", fgsm_eps_tr=0.15)
    assert WB._results_dir_filename(gan, wf) == ("results/whitebox_adv_tr_mnist", "model=C_advTrEps=0.15attack=fgsm.txt")


def test_command_line_flags_fold_into_cfg_and_flags():
    """`--cfg` first, then every cfg key and every script flag as `--lower_case` options (reference utils/config.py,
    blackbox.py:723-759, whitebox.py:344-392)."""
    from defensegan_b200.utils.config import packaged_cfg_path
    cfg, fl = BB._parse(["--cfg", packaged_cfg_path("mnist"), "--defense_type", "defense_gan", "--rec_rr", "5",
                         "--bb_model", "A", "--override", "true"], "blackbox")
    assert (cfg["REC_RR"], cfg["DATASET_NAME"], fl.defense_type, fl.bb_model, fl.override, fl.num_tests) == \
        (5, "mnist", "defense_gan", "A", True, 2000)
    gan = TemplateGAN()
    E.set_test_time_rec_params(gan, fl, cfg)                      # --override applies the command-line values
    assert gan.rec_rr == 5 and gan.rec_iters == cfg["REC_ITERS"]
    cfg, fl = BB._parse(["--cfg", packaged_cfg_path("celeba"), "--alpha", "0.1", "--attack_type", "rand+fgsm"], "whitebox")
    assert (fl.alpha, fl.attack_type, fl.num_tests, cfg["DATASET_NAME"]) == (0.1, "rand+fgsm", -1, "celeba")


def test_train_cli_fills_the_caches_the_attack_scripts_read(tmp_path, monkeypatch):
    """`python -m defensegan_b200.train --save_ds --save_recs` (reference train.py:45-56) -> the files
    `get_cached_gan_data` reads back (blackbox.py:272-367).  The projector is replaced (it needs a GPU)."""
    from defensegan_b200 import train as T
    from defensegan_b200.models.gan import DefenseGANBase
    from defensegan_b200.utils.config import packaged_cfg_path
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(DefenseGANBase, "reconstruct", lambda self, x, **kw: x * 0.5)
    monkeypatch.setattr(DefenseGANBase, "load_generator", lambda self, ckpt_path=None: True)
    rs = np.random.RandomState(0)
    arrays = {}
    for sp, n in (("train", 9), ("dev", 4), ("test", 6)):
        arrays[sp + "_x"] = rs.randint(0, 256, size=(n, 28, 28, 1)).astype("uint8")
        arrays[sp + "_y"] = np.arange(n) % 3
    np.savez("raw.npz", **arrays)
    argv = ["--cfg", packaged_cfg_path("mnist"), "--dataset_npz", "raw.npz", "--save_ds", "--save_recs", "--rec_rr", "2",
            "--rec_iters", "3", "--batch_size", "4", "--output_dir", str(tmp_path / "output")]
    gan = T.main(*T._parse(argv))
    assert (gan.rec_rr, gan.rec_iters, gan.batch_size) == (2, 3, 4)
    rec_dir = gan.rec_cache_dir("train")
    assert os.path.isfile(os.path.join(rec_dir, "feats.pkl")) and len(os.listdir(os.path.join(rec_dir, "pickles"))) == 9
    orig = E.get_cached_gan_data(gan, True, orig_data_flag=True, flags=E.Flags("blackbox"))
    np.testing.assert_allclose(orig.train_images, arrays["train_x"] / 255.0, rtol=1e-6)
    np.testing.assert_allclose(orig.test_images, arrays["test_x"] / 255.0, rtol=1e-6)
    assert orig.train_labels.shape == (9, 3) and orig.train_labels.argmax(1).tolist() == list(np.arange(9) % 3)
    recs = E.get_cached_gan_data(gan, True, flags=E.Flags("blackbox", train_on_recs=True, defense_type="defense_gan"))
    np.testing.assert_allclose(recs.train_images, arrays["train_x"] / 255.0 * 0.5, rtol=1e-6)
    np.testing.assert_allclose(recs.test_images, arrays["test_x"] / 255.0 * 0.5, rtol=1e-6)
    imgs, labels = E.get_pickle_split(os.path.dirname(rec_dir), "train", [28, 28, 1])
    np.testing.assert_allclose(imgs, recs.train_images, rtol=1e-6)
    assert labels.tolist() == list(np.arange(9) % 3)
    with pytest.raises(SystemExit):
        T.main(*T._parse(["--cfg", packaged_cfg_path("mnist"), "--is_train"]))


@pytest.mark.gpu
def test_blackbox_and_whitebox_drive_the_cuda_projection():
    """The same drivers with the real projector: MNIST generator (random-init weights, fp16 tensor-core path) in front
    of classifiers trained on the quadrant patterns.  The patterns are far from that generator's range, so nothing is
    claimed about the defended accuracy - the test is that every reconstruct call of the two pipelines (oracle
    labelling, transfer evaluation, straight-through white-box attack, detection statistic) runs on the CUDA path."""
    from defensegan_b200.models.gan import MnistDefenseGAN
    gan = MnistDefenseGAN(test_mode=True, verbose=False, precision="fp16")
    gan.rec_rr, gan.rec_iters = 4, 30
    data = _data(256, 112)
    flags = E.Flags("blackbox", bb_model="E", sub_model="E", num_tests=-1, fgsm_eps=0.2)
    out = BB.blackbox(gan, defense_type="defense_gan", batch_size=32, learning_rate=0.005, nb_epochs=4, holdout=48,
                      data_aug=2, nb_epochs_s=3, data=data, flags=flags)
    labels, preds, diffs = out["roc_info"]
    assert len(diffs) == 64 and np.all(np.isfinite(diffs)) and 0.0 <= out["bbox_on_sub_adv_ex"] <= 1.0
    assert out["bbox"] > 0.9
    acc, _, roc = WB.whitebox(gan, eps=0.2, attack_type="fgsm", defense_type="defense_gan", batch_size=32,
                              learning_rate=0.005, nb_epochs=4, data=data, num_tests=64,
                              flags=E.Flags("whitebox", model="E", defense_type="defense_gan", attack_type="fgsm"))
    assert len(roc[2]) == 64 and np.all(roc[2] <= 0.2 ** 2 + 1e-6) and 0.0 <= acc <= 1.0
    # projections issued: 2 (oracle labels of the 48 augmented points) + 2 (transfer evaluation, shared with the detection
    # statistic) + 2 x 2 (white-box: one for the straight-through attack, one as layer 0 of the evaluated classifier)
    assert gan._call_counter == 8
    gan.close()
