"""bench.py's contract with the driver, as far as it can be exercised without a GPU: the reference arm prints exactly
ONE JSON line with the contract's keys, and both arms describe the workload with the same `config` object."""
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench_module():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_reference_arm_prints_one_contract_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1",
                          "--warmup", "0", "--ref_sample", "1", "--rec_iters", "2"], stdout=subprocess.PIPE,
                         stderr=subprocess.PIPE, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout
    line = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["steps"] == 1 and line["n_gpus"] == 1 and line["gpu_launches"] == 0
    assert line["e2e"] == {"value": line["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == line["value"] and 1 <= cb["cores"] and cb["cores_present"] >= 1
    assert "workload" in line["config"] and "model" not in line["config"]
    # the same object the GPU arm prints for this command line
    assert line["config"] == _bench_module().arm_config("mnist", 256, 10, 2, 1, "fp16")


def test_config_names_the_baseline_configuration():
    b = _bench_module()
    assert b.arm_config("mnist", 256, 10, 200, 1, "fp16")["baseline_config"] == "configs[1]"
    assert b.arm_config("f-mnist", 256, 10, 200, 1, "fp16")["baseline_config"] == "configs[2]"
    assert b.arm_config("celeba", 128, 10, 200, 1, "fp16")["baseline_config"] == "configs[3]"
    c5 = b.arm_config("mnist", 512, 10, 200, 8, "fp16")
    assert c5["baseline_config"] == "configs[4]" and c5["global_batch"] == 4096 and c5["per_gpu_batch"] == 512
    assert b.arm_config("mnist", 256, 10, 20, 1, "fp16")["baseline_config"] == "custom"
    assert "192 MiB" in c5["l2"]                       # the timing rule: say how L2 is flushed, in `config`
