"""CPU tests of the host side: C-ABI library loads and exports every declared symbol, the
reference-mirroring Python surface, config/flags, the eval driver and the shard/gather logic
(gloo, world_size 2)."""
import argparse
import os
import re
import socket
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    from defensegan_b200 import _native
    _native.build_library()
    lib = _native.load_library()
    header = open(os.path.join(ROOT, "include", "defensegan_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(dgan_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations parsed"
    for sym in declared:
        assert hasattr(lib, sym), "library does not export %s" % sym
    assert sorted(declared) == sorted(_native.ABI_SYMBOLS)
    assert lib.dgan_abi_version() == 1
    d = _native.dgan_desc(1, 0, 128, 64, 0, 0)
    import ctypes
    assert lib.dgan_num_weights(ctypes.byref(d)) == 8
    d.arch = 1
    assert lib.dgan_num_weights(ctypes.byref(d)) == 10


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from defensegan_b200.models.gan import MnistDefenseGAN
    gan = MnistDefenseGAN(test_mode=True, verbose=False)
    with pytest.raises(RuntimeError):
        gan.reconstruct(torch.zeros(2, 28, 28, 1))


def test_model_classes_mirror_reference_defaults(tmp_path):
    from defensegan_b200.models import gan as G
    m = G.MnistDefenseGAN(test_mode=True, verbose=False)
    assert (m.rec_iters, m.rec_rr, m.rec_lr) == (200, 10, 10.0)
    assert (m.latent_dim, m.net_dim, m.use_bn, m.batch_size) == (128, 64, False, 50)
    assert m.image_dim == [28, 28, 1] and m.dataset_name == "mnist" and m.test_batch_size == m.batch_size
    assert m.checkpoint_dir == os.path.join("output", "gans", "mnist")
    f = G.FmnistDefenseDefenseGAN(test_mode=True, verbose=False)
    assert f.dataset_name == "f-mnist" and f.arch == "mnist"
    c = G.CelebADefenseGAN(test_mode=True, verbose=False)
    assert c.rec_rr == 2 and c.image_dim == [64, 64, 3]
    assert set(G.dataset_gan_dict) == {"mnist", "f-mnist", "celeba"}
    # callers override hyper-parameters after construction (blackbox.py:649-658)
    m.rec_rr, m.rec_lr, m.rec_iters = 2, 1.0, 10
    # batch_size % rec_rr assertion (models/gan.py:101-104)
    with pytest.raises(AssertionError):
        G.MnistDefenseGAN(test_mode=True, verbose=False, batch_size=32, rec_rr=10)
    with pytest.raises(TypeError):
        G.MnistDefenseGAN(test_mode=True, verbose=False, not_an_attribute=1)
    # input transforms (models/gan.py:684-685,764-765)
    assert float(m.input_transform(np.array([255.0]))[0]) == 1.0
    assert float(c.input_transform(np.array([0.0]))[0]) == -1.0
    # checkpoint round trip
    p = m.save_generator(str(tmp_path))
    m2 = G.MnistDefenseGAN(test_mode=True, verbose=False, seed=7)
    assert not np.array_equal(m2.weights["Generator.Input/Generator.Input.W"], m.weights["Generator.Input/Generator.Input.W"])
    assert m2.load_generator(str(tmp_path)) is True and os.path.isfile(p)
    np.testing.assert_array_equal(m2.weights["Generator.3/Generator.3.Filters"], m.weights["Generator.3/Generator.3.Filters"])
    assert m2.load_generator(str(tmp_path / "missing")) is False


def test_weight_init_matches_oracle():
    from defensegan_b200 import weights as W
    from oracle import defensegan_oracle as O
    for arch in ("mnist", "celeba"):
        a, b = W.init_generator_weights(arch), O.init_generator_weights(arch)
        assert list(a.keys()) == list(b.keys())
        for k in a:
            np.testing.assert_array_equal(a[k], b[k])
    with pytest.raises(ValueError):
        W.validate_weights("mnist", {k: v[..., :1] for k, v in a.items()}, 128, 64, False)


def test_config_and_flags():
    from defensegan_b200.utils import config as C
    cfg = C.load_config(C.packaged_cfg_path("mnist"))
    assert cfg["REC_ITERS"] == 200 and cfg["REC_RR"] == 10 and cfg["REC_LR"] == 10.0
    assert cfg["LATENT_DIM"] == 128 and cfg["USE_BN"] is False and cfg["IMAGE_DIM"] == [28, 28, 1]
    parser = C.add_flags(argparse.ArgumentParser(), cfg)
    ns = parser.parse_args(["--rec_iters", "20", "--rec_lr", "1.5", "--rec_rr", "5", "--use_bn", "False"])
    over = C.flags_to_cfg(ns, cfg)
    assert (over["REC_ITERS"], over["REC_LR"], over["REC_RR"], over["USE_BN"]) == (20, 1.5, 5, False)
    from defensegan_b200.models.gan import MnistDefenseGAN
    m = MnistDefenseGAN(cfg=over, test_mode=True, verbose=False)
    assert (m.rec_iters, m.rec_lr, m.rec_rr) == (20, 1.5, 5)
    with pytest.raises(IOError):
        C.load_config("/nonexistent/x.yml")


def test_model_eval_gan_eager_contract():
    from defensegan_b200.utils.gan_defense import model_eval_gan, SharedReconstruction
    n, bs = 10, 4
    x = np.random.RandomState(0).rand(n, 28, 28, 1).astype("f4")
    labels = np.eye(10, dtype="f4")[np.arange(n) % 10]
    calls = []

    class FakeGan:
        def reconstruct(self, xb, **kw):
            calls.append(int(xb.shape[0]))
            return xb * 0.5

    rec = SharedReconstruction(FakeGan())

    def predictions(xb):
        r = rec(xb)
        out = torch.zeros(xb.shape[0], 10)
        out[torch.arange(xb.shape[0]), (torch.arange(xb.shape[0]) + len(seen)) % 10] = 1.0
        seen.extend(range(xb.shape[0]))
        return out

    seen = []
    acc, roc = model_eval_gan(None, None, None, predictions=predictions, test_images=x, test_labels=labels,
                              args={"batch_size": bs}, diff_op=lambda xb: ((xb - rec(xb)) ** 2).mean(dim=(1, 2, 3)),
                              device=torch.device("cpu"))
    assert calls == [4, 4, 2]                      # one projection per batch, ragged tail kept
    assert acc == 1.0 and len(roc[0]) == n and len(roc[1]) == n and roc[2].shape == (n,)
    np.testing.assert_allclose(roc[2], ((x * 0.5) ** 2).mean(axis=(1, 2, 3)), rtol=1e-5)
    with pytest.raises(AssertionError):
        model_eval_gan(None, None, None, predictions=predictions, test_images=x, test_labels=labels, args={})
    with pytest.raises(ValueError):
        model_eval_gan(None, None, None, predictions=predictions, args={"batch_size": 2})
    acc2, acc_rec, roc2 = model_eval_gan(None, None, None, predictions=lambda xb: torch.ones(len(xb), 10),
                                         predictions_rec=lambda xb: torch.ones(len(xb), 10), test_images=x,
                                         test_labels=labels, args={"batch_size": bs}, device=torch.device("cpu"))
    assert abs(acc2 - 0.1) < 1e-9 and abs(acc_rec - 0.1) < 1e-9


def test_shard_bounds():
    from defensegan_b200.parallel import shard_bounds
    assert shard_bounds(4096, 8) == [512 * i for i in range(9)]
    assert shard_bounds(10, 4) == [0, 3, 6, 8, 10]
    assert shard_bounds(2, 4) == [0, 1, 2, 2, 2]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, n_images, rec_rr, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from defensegan_b200.parallel import sharded_apply
        g = torch.Generator().manual_seed(0)
        images = torch.rand(n_images, 2, 2, 1, generator=g)
        z0 = torch.rand(n_images * rec_rr, 8, generator=g)

        def local_fn(x, z, out):
            # stand-in for the per-rank projection: depends on the image AND its R z0 rows
            out.copy_(x * 2.0 + z.reshape(x.shape[0], rec_rr, -1).sum(dim=(1, 2)).view(-1, 1, 1, 1))

        got = sharded_apply(local_fn, images, rec_rr, z_init_val=z0)
        want = images * 2.0 + z0.reshape(n_images, rec_rr, -1).sum(dim=(1, 2)).view(-1, 1, 1, 1)
        ret[rank] = bool(torch.equal(got, want))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_images", [8, 5])
def test_sharded_gather_gloo_world2(n_images):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, n_images, 3, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret[0] and ret[1]


def test_reconstruct_dataset_cache_format(tmp_path, monkeypatch):
    """f1 (reference models/gan.py:451-587): directory naming, per-image pickles, feats.pkl short-cut and the
    regex the callers use to parse hyper-parameters back out of the path (blackbox.py:646-651)."""
    import pickle
    import re
    from defensegan_b200.models.gan import MnistDefenseGAN
    gan = MnistDefenseGAN(test_mode=True, verbose=False, output_dir=str(tmp_path))
    gan.initialized = True
    gan.rec_rr, gan.rec_lr, gan.rec_iters = 2, 0.5, 3
    calls = []

    def fake_reconstruct(x, **kw):                 # the projector needs a GPU; the cache logic does not
        calls.append(int(x.shape[0]))
        return x * 0.5

    monkeypatch.setattr(gan, "reconstruct", fake_reconstruct)
    rs = np.random.RandomState(0)
    data = {sp: (rs.randint(0, 256, size=(n, 28, 28, 1)).astype("uint8"), np.arange(n) % 10)
            for sp, n in (("train", 5), ("dev", 3), ("test", 4))}

    def gen(sp, bs=2):
        def g():
            x, y = data[sp]
            for i in range(0, len(x), bs):
                yield x[i:i + bs], y[i:i + bs]
        return g

    gan.set_dataset_generators(train=gen("train"), dev=gen("dev"), test=gen("test"))
    rets = gan.reconstruct_dataset()
    assert calls == [2, 2, 1, 2, 1, 2, 2]
    d = gan.rec_cache_dir("test")
    assert d == os.path.join(str(tmp_path), "gans", "mnist", "recs_rr2_lr0.50000_iters3", "test")
    assert re.compile("recs_rr(.*)_lr(.*)_iters(.*)").findall(os.path.dirname(d))[0] == ("2", "0.50000", "3")
    recs, targets, orig = rets["test"]
    assert recs.shape == (4, 28, 28, 1) and orig.shape == (4, 28, 28, 1) and list(targets) == [0, 1, 2, 3]
    np.testing.assert_allclose(orig, data["test"][0] / 255.0, rtol=1e-6)
    np.testing.assert_allclose(recs, orig * 0.5, rtol=1e-6)
    with open(os.path.join(d, "pickles", "rec_0000003_l3.pkl"), "rb") as f:
        np.testing.assert_allclose(pickle.load(f), recs[3])
    # second call: every batch comes from the per-image cache, no projection runs
    calls.clear()
    rets2 = gan.reconstruct_dataset()
    assert calls == []
    np.testing.assert_array_equal(rets2["train"][0], rets["train"][0])
    # feats.pkl short-cut
    gan.save_recs(rets)
    assert os.path.isfile(os.path.join(d, "feats.pkl"))
    rets3 = gan.reconstruct_dataset()
    np.testing.assert_array_equal(rets3["dev"][0], rets["dev"][0])
    # test_again forces recomputation; max_num changes the directory name
    gan.test_again = True
    gan.reconstruct_dataset()
    assert calls == [2, 2, 1, 2, 1, 2, 2]
    assert gan.rec_cache_dir("dev", max_num=100).endswith(os.path.join("recs_rr2_lr0.50000_iters3_num100", "dev"))
