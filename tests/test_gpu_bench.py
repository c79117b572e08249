"""bench.py's own arm on a small workload: one JSON line carrying every key of the measurement contract."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_line_contract():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--sites", "4096", "--steps", "2", "--warmup", "3",
                          "--no-cpu-baseline"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-800:]
    lines = out.stdout.splitlines()
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "gpu_launches", "e2e", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["metric"] == "candidate_sites_per_sec" and d["unit"] == "sites/s" and d["scaling"] == "weak"
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["vs_baseline"] is None
    assert d["value"] > 0 and d["gpu_launches"] >= 2 * 6                       # six kernels per chunk in the fused mode
    assert d["config"]["workload"].startswith("illumina_30x") and "model" not in d["config"]
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 4096 * 20 * 900 and e["d2h_bytes_per_step"] > 0
    assert e["value"] != d["value"]
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and 0 < r["frac"] < 1 / 3 + 1e-6
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
