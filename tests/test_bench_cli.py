"""bench.py contract on the CPU: the reference arm (the oracle on the host cores) prints exactly one JSON line with the
keys the driver reads; our arm refuses to run without a GPU instead of falling back."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, **env):
    e = dict(os.environ, BENCH_POINTS="20000", **env)
    return subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), *args], capture_output=True, text=True, env=e, timeout=600)


def test_reference_arm_prints_one_json_line():
    r = run_bench("--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mpoints/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 2 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    r = run_bench("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = run_bench("--steps", "1", "--warmup", "3")
    assert r.returncode != 0
    assert r.stdout.strip() == ""          # no JSON line that could be mistaken for a measurement
    assert "no CUDA device" in (r.stderr + r.stdout) or "CUDA" in r.stderr
