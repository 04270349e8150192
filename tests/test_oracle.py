"""CPU tests of the oracle itself: known-answer tests of each TF semantic it pins
and agreement with the committed golden vectors."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import defensegan_oracle as O


def _digest(w):
    h = hashlib.sha256()
    for k, v in w.items():
        h.update(k.encode())
        h.update(np.ascontiguousarray(v).tobytes())
    return h.hexdigest()


def test_deconv_matches_loop_definition():
    rs = np.random.RandomState(1)
    for (h, w_, ci, co) in [(4, 4, 6, 3), (7, 7, 4, 2), (3, 5, 2, 1)]:
        x = rs.randn(2, h, w_, ci).astype("f4")
        f = rs.randn(5, 5, co, ci).astype("f4")
        b = rs.randn(co).astype("f4")
        got = O.tf_deconv_same(torch.tensor(x, dtype=torch.float64), torch.tensor(f, dtype=torch.float64),
                               torch.tensor(b, dtype=torch.float64)).numpy()
        want = O.tf_deconv_same_definition(x, f, b)
        assert got.shape == (2, 2 * h, 2 * w_, co)
        np.testing.assert_allclose(got, want, atol=1e-12)


def test_deconv_single_pixel_known_answer():
    # one input pixel at (o,p)=(1,2) with value 1, one channel: out[2o+ka-1, 2p+kb-1] = F[ka,kb]
    x = np.zeros((1, 3, 4, 1), "f4"); x[0, 1, 2, 0] = 1.0
    f = np.arange(25, dtype="f4").reshape(5, 5, 1, 1)
    out = O.tf_deconv_same(torch.tensor(x), torch.tensor(f), None).numpy()[0, :, :, 0]
    want = np.zeros((6, 8), "f4")
    for ka in range(5):
        for kb in range(5):
            i, j = 2 * 1 + ka - 1, 2 * 2 + kb - 1
            if 0 <= i < 6 and 0 <= j < 8:
                want[i, j] = f[ka, kb, 0, 0]
    np.testing.assert_array_equal(out, want)
    # the PyTorch padding=2/output_padding=1 variant is NOT the TF op
    alt = torch.nn.functional.conv_transpose2d(torch.tensor(x).permute(0, 3, 1, 2), torch.tensor(f).permute(3, 2, 0, 1),
                                               stride=2, padding=2, output_padding=1).numpy()[0, 0]
    assert np.abs(alt - want).max() > 1.0


def test_deconv_backward_to_input_definition():
    rs = np.random.RandomState(2)
    x = torch.tensor(rs.randn(2, 4, 3, 5), dtype=torch.float64, requires_grad=True)
    f = rs.randn(5, 5, 2, 5)
    dout = rs.randn(2, 8, 6, 2)
    y = O.tf_deconv_same(x, torch.tensor(f), None)
    (gx,) = torch.autograd.grad((y * torch.tensor(dout)).sum(), x)
    want = O.tf_deconv_same_dinput_definition(dout, f)
    np.testing.assert_allclose(gx.numpy(), want, atol=1e-11)


def test_tiling_is_image_major_restart_minor():
    x = torch.arange(3 * 2 * 2 * 1, dtype=torch.float32).reshape(3, 2, 2, 1)
    t = O.tile_images(x, 4)
    assert t.shape == (12, 2, 2, 1)
    for i in range(3):
        for r in range(4):
            assert torch.equal(t[i * 4 + r], x[i])


def test_learning_rate_is_constant_dead_decay():
    assert all(O.effective_learning_rate(10.0, 200, t) == 10.0 for t in range(200))
    assert O.effective_learning_rate(10.0, 200, 159, emulate_dead_decay=False) == 10.0
    assert abs(O.effective_learning_rate(10.0, 200, 160, emulate_dead_decay=False) - 1.0) < 1e-12


def test_batchnorm_batch_statistics():
    rs = np.random.RandomState(3)
    x = torch.tensor(rs.randn(6, 3, 3, 4))
    off, sc = torch.tensor(rs.randn(1, 1, 1, 4)), torch.tensor(rs.rand(1, 1, 1, 4) + 0.5)
    y = O.batchnorm_batchstat(x, [0, 1, 2], off, sc)
    m = x.mean(dim=(0, 1, 2), keepdim=True)
    v = x.var(dim=(0, 1, 2), unbiased=False, keepdim=True)
    np.testing.assert_allclose(y.numpy(), ((x - m) / torch.sqrt(v + 1e-5) * sc + off).numpy(), atol=1e-12)


def test_generator_shapes_and_layouts():
    for arch, shape in [("mnist", (28, 28, 1)), ("celeba", (64, 64, 3))]:
        w = O.weights_to_torch(O.init_generator_weights(arch))
        z = torch.tensor(O.sample_z0(3, 128))
        y, hidden = O.generator_forward(arch, w, z, return_hidden=True)
        assert tuple(y.shape) == (3,) + shape
        assert tuple(hidden[0].shape) == (3, 4, 4, 256)
        # Linear column f = (h*4+w)*256 + c
        lin = torch.relu(z @ w["Generator.Input/Generator.Input.W"] + w["Generator.Input/Generator.Input.b"])
        assert torch.equal(hidden[0][1, 2, 3, 17], lin[1, (2 * 4 + 3) * 256 + 17])
        if arch == "mnist":
            assert tuple(hidden[1].shape) == (3, 7, 7, 128) and float(y.min()) >= 0.0 and float(y.max()) <= 1.0
        else:
            assert tuple(hidden[1].shape) == (3, 8, 8, 128) and float(y.min()) >= -1.0


def test_momentum_loop_semantics_small():
    """Hand-rolled check of F4/F5/F6 on a 1-image problem: output is the pre-update forward of
    the last iteration, velocity accumulates lr-free, arg-min picks the lowest index on ties."""
    arch = "mnist"
    w = O.init_generator_weights(arch)
    imgs = O.synthetic_images(arch, w, 1)
    z0 = O.sample_z0(2, 128)
    z0[1] = z0[0]                                   # identical restarts -> exact tie
    r = O.reconstruct(arch, w, imgs, 2, 3, rec_lr=10.0, z_init_val=z0, dtype=torch.float64, return_trace=True)
    assert r["idx"][0] == 0 and r["loss_all"][0] == r["loss_all"][1]
    # replay by hand
    wt = O.weights_to_torch(w, torch.float64)
    x = torch.tensor(imgs, dtype=torch.float64)
    z = torch.tensor(z0[:1], dtype=torch.float64); v = torch.zeros_like(z)
    for t in range(3):
        zt = z.clone().requires_grad_(True)
        y = O.generator_forward(arch, wt, zt)
        loss = ((y - x) ** 2).mean()
        (g,) = torch.autograd.grad(loss, zt)
        if t == 2:
            break
        v = 0.7 * v + g
        z = z - 10.0 * v
    np.testing.assert_allclose(r["rec"][0], y.detach().numpy()[0], atol=1e-12)
    np.testing.assert_allclose(r["loss_min"][0], float(loss), atol=1e-14)
    assert r["trace"].shape == (3, 2)


@pytest.mark.parametrize("case", ["mnist_c1", "mnist_ragged_bias", "celeba_small", "mnist_bn", "celeba_bn"])
def test_golden_vectors_reproduce(golden_dir, case):
    g = np.load(os.path.join(golden_dir, case + ".npz"))
    arch = str(g["arch"])
    use_bn = bool(int(g["use_bn"])) if "use_bn" in g.files else False
    w = O.init_generator_weights(arch, random_bias=bool(int(g["random_bias"])), use_bn=use_bn)
    assert _digest(w) == str(g["weights_sha256"])
    r = O.reconstruct(arch, w, g["images"], int(g["R"]), int(g["L"]), rec_lr=float(g["lr"]), z_init_val=g["z0"],
                      use_bn=use_bn)
    np.testing.assert_allclose(r["rec"], g["rec32"], atol=2e-5 if use_bn else 2e-6)
    np.testing.assert_allclose(r["loss_min"], g["loss_min32"], atol=1e-7)
    np.testing.assert_array_equal(r["idx"], g["idx32"])
    # fp32 run stays close to the fp64 truth at these horizons
    assert np.abs(g["rec32"] - g["rec64"]).max() < 2e-5
    np.testing.assert_array_equal(g["idx32"], g["idx64"])


def test_model_eval_gan_batching():
    seen = []

    def rec_fn(xb, bi):
        seen.append(len(xb))
        return xb

    x = np.arange(10 * 4, dtype="f4").reshape(10, 4)
    labels = np.eye(4, dtype="f4")[np.arange(10) % 4]
    acc, roc = O.model_eval_gan(rec_fn, lambda r: labels[: len(r)] if False else np.eye(4)[np.argmax(r, -1)],
                                x, np.eye(4, dtype="f4")[np.full(10, 3)], batch_size=4,
                                diff_fn=lambda a, b: np.zeros(len(a)))
    assert seen == [4, 4, 2] and acc == 1.0 and len(roc[0]) == 10 and len(roc[2]) == 10
