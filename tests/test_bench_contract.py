"""bench.py's reference arm (the CPU restatement of the reference, the one place outside tests/ and smoke() that may
execute oracle/) prints the contract's JSON line; the product arm refuses to run without a GPU instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _run(args, timeout=300):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    res = _run(["--impl", "reference", "--steps", "1", "--warmup", "1"])
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("keypoint-voting frames/s") and line["value"] > 0
    assert line["steps"] == 1 and line["n_gpus"] == 1 and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["gpu_launches"] == 0 and "workload" in line["config"]


def test_product_arm_needs_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    res = _run(["--steps", "1", "--warmup", "1", "--no-cpu-baseline", "--no-e2e"], timeout=120)
    assert res.returncode != 0  # no CPU or PyTorch fallback of the product path
    assert not any(ln.startswith("{") for ln in res.stdout.splitlines())
