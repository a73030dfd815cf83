"""bench.py's contract on a box without a GPU: the reference arm (the CPU restatement, `--impl reference`) runs and prints ONE JSON line
with the keys the driver reads; the product arm refuses to run without CUDA instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600, cwd=ROOT)


def test_reference_arm_json_line():
    r = run_bench("--impl", "reference", "--reads", "2000", "--ref-bases", "150000", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "aligned reads/sec" and d["unit"] == "reads/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("this box has a GPU: the product arm runs (tests -m gpu, bench.py)")
    r = run_bench("--reads", "2000", "--ref-bases", "150000", "--steps", "1", "--warmup", "1", "--no-cpu-baseline")
    assert r.returncode != 0, "bench.py must not produce a product number without CUDA"
    assert not any(ln.strip().startswith("{") for ln in r.stdout.splitlines()), r.stdout
