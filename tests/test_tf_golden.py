"""The TensorFlow pin.  tools/make_tf_golden.py (run on a Python-2.7 / TF-1.x machine, which this container is not)
drives the UNMODIFIED reference models/gan.py:333-449 on the fixtures of tests/golden/ and writes tests/golden/tf/.
When those files exist these tests tie the oracle (and, with -m gpu, the CUDA path) and the checkpoint reader to
TensorFlow's own outputs; until then they are skipped and parity stays "unpinned" (DESIGN.md section 5).

The part that CAN be checked here is: the Python-2 script's weight restatement reproduces the fixtures' weights.
"""
import importlib.util
import os

import numpy as np
import pytest

from oracle import defensegan_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TF_DIR = os.path.join(ROOT, "tests", "golden", "tf")
CASES = ["mnist_c1", "mnist_ragged_bias", "celeba_small"]


def _tool():
    spec = importlib.util.spec_from_file_location("make_tf_golden", os.path.join(ROOT, "tools", "make_tf_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("case", CASES)
def test_tf_script_restates_the_fixture_weights(golden_dir, case):
    tool = _tool()
    g = np.load(os.path.join(golden_dir, case + ".npz"))
    w = tool.seeded_weights(str(g["arch"]), bool(int(g["random_bias"])))
    assert tool.digest(w) == str(g["weights_sha256"])
    ours = O.init_generator_weights(str(g["arch"]), random_bias=bool(int(g["random_bias"])))
    assert list(w.keys()) == list(ours.keys())
    assert all(np.array_equal(w[k], ours[k]) for k in w)


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_tensorflow_dump(golden_dir, case):
    path = os.path.join(TF_DIR, case + ".npz")
    if not os.path.exists(path):
        pytest.skip("no TensorFlow dump (run tools/make_tf_golden.py on a TF1 machine)")
    g, t = np.load(os.path.join(golden_dir, case + ".npz")), np.load(path)
    # fp32 TF (Eigen/MKL summation order) vs fp32 oracle: short horizons (L <= 10) keep rounding-order effects ~1e-5
    assert np.abs(t["rec_tf"] - g["rec32"]).max() <= 1e-4
    assert np.abs(t["loss_min_tf"] - g["loss_min32"]).max() <= 1e-5


@pytest.mark.parametrize("case", CASES)
def test_bundle_reader_reads_tensorflow_checkpoint(golden_dir, case):
    ckpt = os.path.join(TF_DIR, case + "_ckpt")
    if not os.path.isdir(ckpt):
        pytest.skip("no TensorFlow checkpoint (run tools/make_tf_golden.py on a TF1 machine)")
    from defensegan_b200 import tf_bundle
    g = np.load(os.path.join(golden_dir, case + ".npz"))
    prefix = tf_bundle.latest_checkpoint(ckpt)
    got = tf_bundle.read_generator_variables(prefix)
    want = O.init_generator_weights(str(g["arch"]), random_bias=bool(int(g["random_bias"])))
    assert sorted(got) == sorted(want)
    assert all(np.array_equal(np.asarray(got[k]).reshape(want[k].shape), want[k]) for k in want)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_cuda_path_matches_tensorflow_dump(golden_dir, case):
    path = os.path.join(TF_DIR, case + ".npz")
    if not os.path.exists(path):
        pytest.skip("no TensorFlow dump (run tools/make_tf_golden.py on a TF1 machine)")
    import torch
    from defensegan_b200 import _native
    g, t = np.load(os.path.join(golden_dir, case + ".npz")), np.load(path)
    w = O.init_generator_weights(str(g["arch"]), random_bias=bool(int(g["random_bias"])))
    for precision, tol in (("fp32", 1e-4), ("fp16", 2e-2)):
        gen = _native.NativeGenerator(str(g["arch"]), [torch.as_tensor(v).cuda() for v in w.values()], precision=precision)
        rec = gen.reconstruct(torch.tensor(g["images"]).cuda(), int(g["R"]), int(g["L"]), float(g["lr"]),
                              z_init_val=torch.tensor(g["z0"]).cuda())
        assert np.abs(rec.cpu().numpy() - t["rec_tf"]).max() <= tol
        gen.close()
