"""bench.py's output contract where it can be checked without a GPU: the reference arm (`--impl reference` times the CPU
port of the path on the host cores) prints ONE JSON line on stdout with the keys the driver reads, and the default arm
refuses to run without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_one_json_line():
    out = _run("--impl", "reference", "--steps", "2", "--warmup", "1", "--series", "4", "--points", "50000", "--cpu-seconds", "1")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [line for line in out.stdout.splitlines() if line.strip()]
    assert len(lines) == 1, out.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["steps"] == 2 and line["warmup"] == 1 and line["n_gpus"] == 1
    assert line["metric"] == "compress+grid+aggregate data points/s" and line["unit"] == "points/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["vs_baseline"] is None
    assert "workload" in line["config"] and "model" not in line["config"]
    baseline = line["cpu_baseline"]
    assert baseline["kind"] == "port" and baseline["cores"] >= 1 and baseline["sample"] and baseline["value"] == line["value"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["unit"] == line["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert line["gpu_launches"] == 0


def test_default_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present: the default arm would run the benchmark")
    out = _run("--steps", "1", "--warmup", "0", "--series", "2", "--points", "1000", "--no-e2e", "--no-cpu-baseline", timeout=300)
    assert out.returncode != 0  # no CPU fallback behind the product arm
    assert not [line for line in out.stdout.splitlines() if line.strip().startswith("{")]
