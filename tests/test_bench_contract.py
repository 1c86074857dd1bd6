"""CPU: the reference arm of bench.py (`--impl reference`) runs here without a GPU and prints one JSON line with the contract's
keys; our own arm must refuse to run without CUDA instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, capture_output=True, text=True, timeout=600)


def test_reference_arm_prints_the_contract_line():
    p = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-envs", "8")   # (default: calibrated E in 64..512)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["scaling"] == "weak" and line["vs_baseline"] is None
    assert line["metric"].startswith("env_steps_per_sec") and line["unit"] == "env*steps/s" and line["data"] == "synthetic"
    assert line["steps"] == 1 and line["warmup"] == 0 and line["n_gpus"] == 1 and line["value"] > 0 and line["ms_per_step"] > 0
    assert "workload" in line["config"] and "model" not in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_our_arm_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        return                                                        # on a GPU box the real bench is the driver's job
    p = _run("--steps", "1", "--warmup", "0")
    assert p.returncode != 0 and "CUDA" in (p.stderr + p.stdout)
