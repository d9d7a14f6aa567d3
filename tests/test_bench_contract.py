"""CPU-side checks of bench.py's contract: the reference arm (the one leg that runs without a GPU) prints ONE JSON line
with the keys the driver reads, and the product arm refuses to run without CUDA instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout=300):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, capture_output=True, text=True,
                          timeout=timeout)


def test_reference_arm_prints_the_contract_line():
    p = _run(["--impl", "reference", "--workload", "tiny", "--steps", "1", "--warmup", "1"])
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["metric"] == "EI candidates/sec" and d["unit"] == "candidates/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] in ("port", "reference")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_product_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the product arm would simply run")
    p = _run(["--workload", "tiny", "--steps", "1", "--warmup", "1", "--no-cpu", "--no-next"], timeout=120)
    assert p.returncode != 0                     # fails loudly ...
    assert not [l for l in p.stdout.splitlines() if l.startswith("{") and '"value"' in l]   # ... and reports no number
    assert "CUDA" in (p.stderr + p.stdout)
