"""bench.py's reference arm runs without a GPU: the CPU form of the hot path on the host cores.
Checks the one-line JSON contract (keys, units, the cpu_baseline and e2e objects) on a tiny budget,
and that the GPU arm refuses to run without a device instead of falling back to anything."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def run_bench(*args, timeout=300):
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True,
                          timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_the_contract_line(oracle):
    out = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-budget", "3")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "Mrays/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("Mrays/s @1080p 64spp") and line["n_gpus"] == 1 and line["steps"] == 1
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["gpu_launches"] == 0
    base = line["cpu_baseline"]
    assert base["kind"] in ("reference", "port") and base["cores"] >= 1 and base["value"] == line["value"]
    assert "spp per step at 1920x1080" in base["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]


def test_reference_arm_on_other_ranks_prints_nothing(oracle, monkeypatch):
    monkeypatch.setenv("RANK", "1")
    monkeypatch.setenv("WORLD_SIZE", "2")
    out = run_bench("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--cpu-budget", "1")
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_gpu_arm_refuses_to_run_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    for extra in ([], ["--workload", "mesh"]):
        out = run_bench(*extra, "--steps", "1", "--no-cpu")
        assert out.returncode != 0
        assert "no CUDA device" in (out.stderr + out.stdout)
