"""bench.py's JSON contract, checked on CPU through the reference arm (the only arm that runs without a GPU)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "local_laplacian Mpixels/s" and d["unit"] == "Mpixels/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["warmup"] >= 3 and d["steps"] == 1
    assert d["value"] > 0 and abs(d["e2e"]["value"] - d["value"]) < 1e-9
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    # the north-star configuration is the default workload; both arms print the same `config` object
    assert d["config"]["workload"] == "local_laplacian_16k" and d["config"]["frame"] == [16384, 16384, 3]
    assert d["scaling"] == "strong" and d["data"] == "synthetic" and d["gpu_launches"] == 0
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.make_config("local_laplacian_16k", 1)


def test_product_arm_fails_loudly_without_cuda():
    """No CPU fallback: without a GPU the product arm must exit with an error, not print a number."""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True,
                         text=True, timeout=600, cwd=ROOT)
    assert out.returncode != 0
    assert not any(l.startswith("{") and '"value"' in l for l in out.stdout.splitlines())
